// hostio.cu -- host <-> device copies for CALLER-OWNED buffers that may be pageable.
//
// The reference's callers hand `transform` two memory maps of the challenge / response files
// (powersoftau/src/bin/compute_constrained.rs:83-132: `MmapOptions::new().map(&reader)`, `map_mut(&writer)`), i.e. pageable,
// file-backed memory.  cudaMemcpyAsync from such memory is staged by the driver on the calling thread (one thread,
// synchronous), which serialises the three-stream pipelines of this library.  Here:
//   * a buffer that is already page-locked (cudaHostAlloc / cudaHostRegister) is copied directly, asynchronously;
//   * a pageable source is cut into 8 MiB pieces that a pool of host threads copies into their own pinned slots (two per
//     thread: one being filled while the other is in flight); each thread issues its piece's cudaMemcpyAsync on the caller's
//     stream itself, and the call returns once every piece has been issued;
//   * a pageable destination receives its bytes from a ring of pinned slots that the same threads empty as the D2H events
//     complete (several at a time: the first touch of a fresh output mapping is page-fault bound), so the thread that queues
//     GPU work never waits for a device -> pageable copy.
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
#include <nvtx3/nvToolsExt.h>
#include "p2b_internal.h"

namespace p2b {

NvtxRange::NvtxRange(const char *name) { nvtxRangePushA(name); }
NvtxRange::~NvtxRange() { nvtxRangePop(); }

struct HostIO {
    static constexpr size_t SLOT = (size_t)8 << 20;
    static constexpr int MAXT = 16, NOUT = 16;
    struct Slot { char *buf = nullptr; cudaEvent_t ev = nullptr; };
    // H2D: every worker owns two pinned slots (fill one while the other is in flight); D2H: a shared ring the queueing thread
    // hands out and the workers empty
    Slot in[MAXT][2];
    Slot out[NOUT];
    int next_out = 0;
    int device = 0;
    int nthreads = 1;
    std::vector<std::thread> pool;
    struct Task {
        int kind;                 // 0: stage `len` bytes from src into a slot and copy them to d_dst on `stream`; 1: drain out[slot] into dst
        const char *src; char *d_dst; cudaStream_t stream;
        int slot; char *dst;
        size_t len;
    };
    std::deque<Task> q_in, q_out;   // separate queues and threads per direction: a drain task blocks on a GPU event that may
    int n_in = 1;                   // be a whole chunk of compute away, and must never sit in front of the next chunk's uploads
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    size_t h2d_pending = 0;       // staged H2D tasks queued or running (the queueing thread waits for 0: all copies ISSUED)
    size_t d2h_inflight = 0;
    int out_busy[NOUT] = {};
    bool stop = false;
    bool ok = true;
    uint64_t staged_in = 0, staged_out = 0;     // bytes that went through the rings (p2b_io_stats)
};

static void worker_main(HostIO *io, int t) {
    cudaSetDevice(io->device);
    int flip = 0;
    std::deque<HostIO::Task> &q = t < io->n_in ? io->q_in : io->q_out;
    for (;;) {
        HostIO::Task k;
        {
            std::unique_lock<std::mutex> lk(io->mu);
            io->cv_work.wait(lk, [&] { return io->stop || !q.empty(); });
            if (q.empty()) return;              // stop requested and nothing left
            k = q.front();
            q.pop_front();
        }
        bool good = true;
        if (k.kind == 0) {
            HostIO::Slot &sl = io->in[t][flip];
            flip ^= 1;
            good = cudaEventSynchronize(sl.ev) == cudaSuccess;          // this worker's previous copy out of the slot has finished
            if (good) {
                memcpy(sl.buf, k.src, k.len);
                good = cudaMemcpyAsync(k.d_dst, sl.buf, k.len, cudaMemcpyHostToDevice, k.stream) == cudaSuccess &&
                       cudaEventRecord(sl.ev, k.stream) == cudaSuccess;
            }
        } else {
            good = cudaEventSynchronize(io->out[k.slot].ev) == cudaSuccess;
            if (good) memcpy(k.dst, io->out[k.slot].buf, k.len);
        }
        {
            std::lock_guard<std::mutex> lk(io->mu);
            if (!good) io->ok = false;
            if (k.kind == 0) io->h2d_pending--;
            else { io->out_busy[k.slot] = 0; io->d2h_inflight--; }
        }
        io->cv_done.notify_all();
    }
}

static HostIO *io_get(Ctx *c) {
    if (c->io) return c->io;
    HostIO *io = new HostIO();
    io->device = c->device;
    int nt = 6;
    if (const char *e = getenv("P2B_COPY_THREADS")) nt = atoi(e);
    const int hw = (int)std::thread::hardware_concurrency();
    if (hw > 0 && nt > hw) nt = hw;
    if (nt < 1) nt = 1;
    if (nt > HostIO::MAXT) nt = HostIO::MAXT;
    io->nthreads = nt;
    io->n_in = nt > 2 ? nt - (nt >= 6 ? 2 : 1) : 1;      // 6 threads: 4 upload + 2 drain
    if (nt == 1) { nt = 2; io->nthreads = 2; }           // always at least one thread per direction
    bool ok = true;
    auto mk = [&](HostIO::Slot &s) {
        ok = ok && cudaHostAlloc((void **)&s.buf, HostIO::SLOT, cudaHostAllocDefault) == cudaSuccess &&
             cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming) == cudaSuccess;
    };
    for (int t = 0; t < nt; t++) { mk(io->in[t][0]); mk(io->in[t][1]); }
    for (auto &s : io->out) mk(s);
    if (!ok) { delete io; return nullptr; }        // (slots leak on this path; the ctx is unusable anyway)
    for (int t = 0; t < nt; t++) io->pool.emplace_back(worker_main, io, t);
    c->io = io;
    return io;
}
void io_destroy(Ctx *c) {
    HostIO *io = c->io;
    if (!io) return;
    { std::lock_guard<std::mutex> lk(io->mu); io->stop = true; }
    io->cv_work.notify_all();
    for (auto &t : io->pool) t.join();
    for (int t = 0; t < HostIO::MAXT; t++)
        for (auto &s : io->in[t]) { if (s.buf) cudaFreeHost(s.buf); if (s.ev) cudaEventDestroy(s.ev); }
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
    {
        std::lock_guard<std::mutex> lk(io->mu);
        for (size_t off = 0; off < bytes; off += HostIO::SLOT) {
            const size_t len = std::min(HostIO::SLOT, bytes - off);
            io->q_in.push_back(HostIO::Task{0, (const char *)h_src + off, (char *)d_dst + off, s, 0, nullptr, len});
            io->h2d_pending++;
        }
    }
    io->cv_work.notify_all();
    // every slot copy has been ISSUED on `s` when this returns, so whatever the caller queues on `s` next is ordered after them
    std::unique_lock<std::mutex> lk(io->mu);
    io->cv_done.wait(lk, [&] { return io->h2d_pending == 0; });
    io->staged_in += bytes;
    if (!io->ok) { io->ok = true; return ctx_fail(c, P2B_ECUDA, "a staged host-to-device copy failed"); }
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
            std::unique_lock<std::mutex> lk(io->mu);
            io->cv_done.wait(lk, [&] { return !io->out_busy[slot]; });
            io->out_busy[slot] = 1;
            io->d2h_inflight++;
        }
        cudaError_t e = cudaMemcpyAsync(io->out[slot].buf, (const char *)d_src + off, len, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaEventRecord(io->out[slot].ev, s);
        if (e != cudaSuccess) {
            std::lock_guard<std::mutex> lk(io->mu);
            io->out_busy[slot] = 0;
            io->d2h_inflight--;
            return ctx_cuda(c, e, "staged D2H");
        }
        {
            std::lock_guard<std::mutex> lk(io->mu);
            io->q_out.push_back(HostIO::Task{1, nullptr, nullptr, nullptr, slot, (char *)h_dst + off, len});
        }
        io->cv_work.notify_all();
    }
    io->staged_out += bytes;
    return P2B_OK;
}

// every staged D2H has landed in the caller's buffer (called after the streams have drained)
int io_flush(Ctx *c) {
    HostIO *io = c->io;
    if (!io) return P2B_OK;
    std::unique_lock<std::mutex> lk(io->mu);
    io->cv_done.wait(lk, [&] { return io->d2h_inflight == 0; });
    if (!io->ok) { io->ok = true; return ctx_fail(c, P2B_ECUDA, "a staged device-to-host copy failed"); }
    return P2B_OK;
}
void io_stats(Ctx *c, uint64_t *staged_in, uint64_t *staged_out) {
    if (staged_in) *staged_in = c->io ? c->io->staged_in : 0;
    if (staged_out) *staged_out = c->io ? c->io->staged_out : 0;
}

}  // namespace p2b
