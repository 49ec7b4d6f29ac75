// hostio.cu -- host <-> device copies for CALLER-OWNED buffers that may be pageable.
//
// The reference's callers hand `transform` two memory maps of the challenge / response files
// (powersoftau/src/bin/compute_constrained.rs:83-132: `MmapOptions::new().map(&reader)`, `map_mut(&writer)`), i.e. pageable,
// file-backed memory.  cudaMemcpyAsync from such memory is staged by the driver on the calling thread (one thread,
// synchronous), which serialises the three-stream pipelines of this library.  Here:
//   * a buffer that is already page-locked (cudaHostAlloc / cudaHostRegister) is copied directly, asynchronously;
//   * a pageable source is copied by a small pool of host threads into a ring of pinned slots, each slot then goes to the
//     device with cudaMemcpyAsync on the caller's stream (the next slot is being filled while the previous one is in flight);
//   * a pageable destination receives its bytes from a ring of pinned slots that a drain thread empties as the D2H events
//     complete, so the thread that queues GPU work never waits for a device -> pageable copy.
// cudaHostRegister on the caller's range is not used: page-locking costs about as much per byte as the copy itself, it is
// paid before the first byte moves, and it is refused for some file-backed mappings.
#include <algorithm>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>
#include "p2b_internal.h"

namespace p2b {

struct HostIO {
    static constexpr size_t SLOT = (size_t)8 << 20;
    static constexpr int NIN = 6, NOUT = 12;
    struct Slot { char *buf = nullptr; cudaEvent_t ev = nullptr; };
    Slot in[NIN], out[NOUT];
    int next_in = 0, next_out = 0;
    int device = 0;
    // ---- pool for parallel memcpy (pageable -> pinned)
    int nthreads = 1;
    std::vector<std::thread> pool;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    char *j_dst = nullptr;
    const char *j_src = nullptr;
    size_t j_len = 0;
    uint64_t gen = 0;
    int remaining = 0;
    bool stop = false;
    // ---- drain thread (pinned -> pageable)
    struct Item { int slot; char *dst; size_t len; };
    std::thread drain;
    std::deque<Item> q;
    std::mutex qmu;
    std::condition_variable qcv, qfree;
    int out_busy[NOUT] = {};
    size_t inflight = 0;
    uint64_t staged_in = 0, staged_out = 0;     // bytes that went through the rings (p2b_io_stats)
    bool ok = true;
};

static void piece(const HostIO *io, int t, size_t &lo, size_t &hi) {
    const size_t per = ((io->j_len + io->nthreads - 1) / io->nthreads + 4095) & ~(size_t)4095;
    lo = std::min(io->j_len, per * (size_t)t);
    hi = std::min(io->j_len, lo + per);
}
static void pool_main(HostIO *io, int t) {
    uint64_t seen = 0;
    for (;;) {
        std::unique_lock<std::mutex> lk(io->mu);
        io->cv_work.wait(lk, [&] { return io->stop || io->gen != seen; });
        if (io->stop) return;
        seen = io->gen;
        size_t lo, hi;
        piece(io, t, lo, hi);
        char *d = io->j_dst;
        const char *s = io->j_src;
        lk.unlock();
        if (hi > lo) memcpy(d + lo, s + lo, hi - lo);
        lk.lock();
        if (--io->remaining == 0) io->cv_done.notify_one();
    }
}
static void par_memcpy(HostIO *io, char *dst, const char *src, size_t len) {
    if (io->nthreads <= 1 || len < ((size_t)1 << 20)) { memcpy(dst, src, len); return; }
    {
        std::lock_guard<std::mutex> lk(io->mu);
        io->j_dst = dst; io->j_src = src; io->j_len = len;
        io->remaining = io->nthreads - 1;
        io->gen++;
    }
    io->cv_work.notify_all();
    size_t lo, hi;
    piece(io, 0, lo, hi);
    if (hi > lo) memcpy(dst + lo, src + lo, hi - lo);
    std::unique_lock<std::mutex> lk(io->mu);
    io->cv_done.wait(lk, [&] { return io->remaining == 0; });
}
static void drain_main(HostIO *io) {
    cudaSetDevice(io->device);
    for (;;) {
        HostIO::Item it;
        {
            std::unique_lock<std::mutex> lk(io->qmu);
            io->qcv.wait(lk, [&] { return io->stop || !io->q.empty(); });
            if (io->q.empty()) return;          // stop requested and nothing left
            it = io->q.front();
            io->q.pop_front();
        }
        if (cudaEventSynchronize(io->out[it.slot].ev) != cudaSuccess) io->ok = false;
        else memcpy(it.dst, io->out[it.slot].buf, it.len);
        {
            std::lock_guard<std::mutex> lk(io->qmu);
            io->out_busy[it.slot] = 0;
            io->inflight--;
        }
        io->qfree.notify_all();
    }
}

static HostIO *io_get(Ctx *c) {
    if (c->io) return c->io;
    HostIO *io = new HostIO();
    io->device = c->device;
    bool ok = true;
    for (auto &s : io->in) ok = ok && cudaHostAlloc((void **)&s.buf, HostIO::SLOT, cudaHostAllocDefault) == cudaSuccess &&
                                cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming) == cudaSuccess;
    for (auto &s : io->out) ok = ok && cudaHostAlloc((void **)&s.buf, HostIO::SLOT, cudaHostAllocDefault) == cudaSuccess &&
                                 cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { delete io; return nullptr; }        // (slots leak on this path; the ctx is unusable anyway)
    int nt = 6;
    if (const char *e = getenv("P2B_COPY_THREADS")) nt = atoi(e);
    const int hw = (int)std::thread::hardware_concurrency();
    if (hw > 0 && nt > hw) nt = hw;
    if (nt < 1) nt = 1;
    io->nthreads = nt;
    for (int t = 1; t < nt; t++) io->pool.emplace_back(pool_main, io, t);
    io->drain = std::thread(drain_main, io);
    c->io = io;
    return io;
}
void io_destroy(Ctx *c) {
    HostIO *io = c->io;
    if (!io) return;
    { std::lock_guard<std::mutex> lk(io->mu); io->stop = true; }
    { std::lock_guard<std::mutex> lk(io->qmu); io->stop = true; }
    io->cv_work.notify_all();
    io->qcv.notify_all();
    for (auto &t : io->pool) t.join();
    if (io->drain.joinable()) io->drain.join();
    for (auto &s : io->in) { if (s.buf) cudaFreeHost(s.buf); if (s.ev) cudaEventDestroy(s.ev); }
    for (auto &s : io->out) { if (s.buf) cudaFreeHost(s.buf); if (s.ev) cudaEventDestroy(s.ev); }
    delete io;
    c->io = nullptr;
}

// page-locked (cudaHostAlloc / cudaHostRegister / managed) host memory can be handed to the copy engine directly
static bool host_is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}
static bool force_staging() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("P2B_FORCE_STAGING"); v = e && atoi(e) ? 1 : 0; }   // test hook: treat every buffer as pageable
    return v == 1;
}

int io_h2d(Ctx *c, void *d_dst, const void *h_src, size_t bytes, cudaStream_t s) {
    if (!bytes) return P2B_OK;
    if (bytes <= 65536 || (!force_staging() && host_is_pinned(h_src))) {
        P2B_CUDA(c, cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, s));
        return P2B_OK;
    }
    HostIO *io = io_get(c);
    if (!io) return ctx_fail(c, P2B_ECUDA, "pinned staging buffers could not be allocated");
    for (size_t off = 0; off < bytes; off += HostIO::SLOT) {
        const size_t len = std::min(HostIO::SLOT, bytes - off);
        HostIO::Slot &sl = io->in[io->next_in];
        io->next_in = (io->next_in + 1) % HostIO::NIN;
        P2B_CUDA(c, cudaEventSynchronize(sl.ev));              // the copy that last read this slot has finished
        par_memcpy(io, sl.buf, (const char *)h_src + off, len);
        P2B_CUDA(c, cudaMemcpyAsync((char *)d_dst + off, sl.buf, len, cudaMemcpyHostToDevice, s));
        P2B_CUDA(c, cudaEventRecord(sl.ev, s));
    }
    io->staged_in += bytes;
    return P2B_OK;
}

int io_d2h(Ctx *c, void *h_dst, const void *d_src, size_t bytes, cudaStream_t s) {
    if (!bytes) return P2B_OK;
    if (bytes <= 65536 || (!force_staging() && host_is_pinned(h_dst))) {
        P2B_CUDA(c, cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, s));
        return P2B_OK;
    }
    HostIO *io = io_get(c);
    if (!io) return ctx_fail(c, P2B_ECUDA, "pinned staging buffers could not be allocated");
    for (size_t off = 0; off < bytes; off += HostIO::SLOT) {
        const size_t len = std::min(HostIO::SLOT, bytes - off);
        const int slot = io->next_out;
        io->next_out = (io->next_out + 1) % HostIO::NOUT;
        {
            std::unique_lock<std::mutex> lk(io->qmu);
            io->qfree.wait(lk, [&] { return !io->out_busy[slot]; });
            io->out_busy[slot] = 1;
            io->inflight++;
        }
        cudaError_t e = cudaMemcpyAsync(io->out[slot].buf, (const char *)d_src + off, len, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaEventRecord(io->out[slot].ev, s);
        if (e != cudaSuccess) {
            std::lock_guard<std::mutex> lk(io->qmu);
            io->out_busy[slot] = 0;
            io->inflight--;
            return ctx_cuda(c, e, "staged D2H");
        }
        {
            std::lock_guard<std::mutex> lk(io->qmu);
            io->q.push_back(HostIO::Item{slot, (char *)h_dst + off, len});
        }
        io->qcv.notify_one();
    }
    io->staged_out += bytes;
    return P2B_OK;
}

// every staged D2H has landed in the caller's buffer (called after the streams have drained)
int io_flush(Ctx *c) {
    HostIO *io = c->io;
    if (!io) return P2B_OK;
    std::unique_lock<std::mutex> lk(io->qmu);
    io->qfree.wait(lk, [&] { return io->inflight == 0; });
    if (!io->ok) { io->ok = true; return ctx_fail(c, P2B_ECUDA, "a staged device-to-host copy failed"); }
    return P2B_OK;
}
void io_stats(Ctx *c, uint64_t *staged_in, uint64_t *staged_out) {
    if (staged_in) *staged_in = c->io ? c->io->staged_in : 0;
    if (staged_out) *staged_out = c->io ? c->io->staged_out : 0;
}

}  // namespace p2b
