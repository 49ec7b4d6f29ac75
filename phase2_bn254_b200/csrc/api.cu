// api.cu -- context management and the host side of the C ABI (include/p2b.h).
//
// Host-side mirror, in C++, of the two reference entry points this core sits behind:
//   BatchedAccumulator::transform   powersoftau/src/batched_accumulator.rs:1119-1292  -> p2b_pot_transform
//   MPCParameters::contribute       phase2/src/parameters.rs:414-522                  -> p2b_phase2_contribute
// plus the level-1 batch_exp replacements.  Host buffers are streamed through the GPU in chunks with a
// 3-stream pipeline (H2D | compute | D2H, double buffered) -- the GPU analogue of the reference's out-of-core
// chunk loop over two mmaps (batched_accumulator.rs:1187,1242).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "blake2b.h"
#include "ec.cuh"
#include "p2b_internal.h"

namespace p2b {

int ctx_fail(Ctx *c, int code, const std::string &msg) {
    c->last_error = msg;
    return code;
}
int ctx_cuda(Ctx *c, cudaError_t e, const char *what) {
    c->last_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
    return P2B_ECUDA;
}
int dev_reserve(Ctx *c, DevBuf &b, size_t bytes) {
    if (bytes <= b.cap) return P2B_OK;
    if (b.p) {
        P2B_CUDA(c, cudaStreamSynchronize(c->stream));
        P2B_CUDA(c, cudaFree(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    size_t cap = (bytes + (1u << 20) - 1) & ~(size_t)((1u << 20) - 1);
    P2B_CUDA(c, cudaMalloc(&b.p, cap));
    b.cap = cap;
    return P2B_OK;
}
void prof_begin(Ctx *c, int slot, cudaStream_t stream) {
    if (!c->prof) return;
    Ctx::ProfSlot &s = c->prof_slot[slot];
    if (s.used == s.ev.size()) {
        cudaEvent_t a, b;
        if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
        s.ev.emplace_back(a, b);
    }
    cudaEventRecord(s.ev[s.used].first, stream ? stream : c->stream);
}
void prof_end(Ctx *c, int slot, int kernels, cudaStream_t stream) {
    if (!c->prof) return;
    Ctx::ProfSlot &s = c->prof_slot[slot];
    if (s.used >= s.ev.size()) return;
    cudaEventRecord(s.ev[s.used].second, stream ? stream : c->stream);
    s.used++;
    s.kernels += (uint64_t)kernels;
}
static int err_reset(Ctx *c) {
    P2B_CUDA(c, cudaMemsetAsync(c->d_err, 0xff, sizeof(unsigned long long), c->stream));
    return P2B_OK;
}
int ctx_collect_error(Ctx *c) {
    P2B_CUDA(c, cudaMemcpyAsync(c->h_err, c->d_err, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    P2B_CUDA(c, cudaStreamSynchronize(c->stream));
    P2B_CUDA(c, cudaStreamSynchronize(c->copy_out));
    {
        int rc = io_flush(c);
        if (rc) return rc;
    }
    unsigned long long e = *c->h_err;
    if (e == ERR_NONE) return P2B_OK;
    cudaMemsetAsync(c->d_err, 0xff, sizeof(unsigned long long), c->stream);   // a reported error never leaks into later calls
    int kind = (int)((e >> 4) & 0xf), sub = (int)(e & 0xf);
    c->err_index = e >> 8;
    c->err_sub = sub;
    char msg[160];
    static const char *names[] = {"ok", "scalar not canonical", "point decoding error", "point at infinity in the input",
                                  "your contribution happened to produce a point at infinity, please re-run", "cuda"};
    snprintf(msg, sizeof msg, "%s (element %llu, sub-code %d)", names[kind < 6 ? kind : 5],
             (unsigned long long)c->err_index, sub);
    c->last_error = msg;
    return kind;
}

// ------------------------------------------------------------------------------------------------- chunked host pipeline
static constexpr size_t CHUNK_POINTS = (size_t)1 << 21;

struct HostJob {
    int g2;
    const uint8_t *in;
    uint8_t *out;
    size_t n;
    ScalarSpec sc;                 // mode 0: sc.d_scalars unused, host_scalars used instead
    const uint8_t *host_scalars;   // n x 32 (mode 0)
    int in_enc, out_enc, flags;
};

// queues the whole job; does not collect the error word (callers batch several jobs, then collect once)
static int run_host_job(Ctx *c, const HostJob &j) {
    if (j.n == 0) return P2B_OK;
    const size_t isz = enc_size(j.g2, j.in_enc), osz = enc_size(j.g2, j.out_enc);
    const size_t chunk = j.n < CHUNK_POINTS ? j.n : CHUNK_POINTS;
    const size_t in_bytes = chunk * isz + (j.sc.mode == 0 ? chunk * 32 : 0);
    int rc;
    for (int b = 0; b < 2; b++) {
        if ((rc = dev_reserve(c, c->stage_in[b], in_bytes))) return rc;
        if ((rc = dev_reserve(c, c->stage_out[b], chunk * osz))) return rc;
    }
    cudaEvent_t *ev_in = c->ev, *ev_cdone = c->ev + 2, *ev_odone = c->ev + 4;
    size_t nchunks = (j.n + chunk - 1) / chunk;
    for (size_t ci = 0; ci < nchunks; ci++) {
        const int b = (int)(ci & 1);
        const size_t off = ci * chunk, m = (off + chunk <= j.n) ? chunk : j.n - off;
        // H2D (the staging buffer is free once the compute of chunk ci-2 is done)
        if (ci >= 2) P2B_CUDA(c, cudaStreamWaitEvent(c->copy_in, ev_cdone[b], 0));
        else P2B_CUDA(c, cudaStreamWaitEvent(c->copy_in, c->ev[6], 0));   // after earlier work on the compute stream
        char *d_in = (char *)c->stage_in[b].p;
        if ((rc = io_h2d(c, d_in, j.in + off * isz, m * isz, c->copy_in))) return rc;
        ScalarSpec sc = j.sc;
        if (sc.mode == 0) {
            char *d_sc = d_in + chunk * isz;
            if ((rc = io_h2d(c, d_sc, j.host_scalars + off * 32, m * 32, c->copy_in))) return rc;
            sc.d_scalars = d_sc;
        } else if (sc.mode == 2) sc.start = j.sc.start + off;
        P2B_CUDA(c, cudaEventRecord(ev_in[b], c->copy_in));
        // compute (the output staging buffer is free once the D2H of chunk ci-2 is done)
        P2B_CUDA(c, cudaStreamWaitEvent(c->stream, ev_in[b], 0));
        if (ci >= 2) P2B_CUDA(c, cudaStreamWaitEvent(c->stream, ev_odone[b], 0));
        if ((rc = launch_batch_mul(c, j.g2, d_in, c->stage_out[b].p, m, sc, j.in_enc, j.out_enc, j.flags, off))) return rc;
        P2B_CUDA(c, cudaEventRecord(ev_cdone[b], c->stream));
        // D2H
        P2B_CUDA(c, cudaStreamWaitEvent(c->copy_out, ev_cdone[b], 0));
        if (j.out && (rc = io_d2h(c, j.out + off * osz, c->stage_out[b].p, m * osz, c->copy_out))) return rc;   // out == NULL: validation only
        P2B_CUDA(c, cudaEventRecord(ev_odone[b], c->copy_out));
    }
    // the next job may reuse both staging pairs: make the compute stream wait for the last two D2H copies,
    // and publish a marker the next job's first H2D waits on
    P2B_CUDA(c, cudaStreamWaitEvent(c->stream, ev_odone[0], 0));
    if (nchunks > 1) P2B_CUDA(c, cudaStreamWaitEvent(c->stream, ev_odone[1], 0));
    P2B_CUDA(c, cudaEventRecord(c->ev[6], c->stream));
    return P2B_OK;
}

static int begin_call(Ctx *c) {
    P2B_CUDA(c, cudaSetDevice(c->device));
    c->last_error.clear();
    c->err_index = 0;
    c->err_sub = 0;
    int rc = err_reset(c);
    if (rc) return rc;
    P2B_CUDA(c, cudaEventRecord(c->ev[6], c->stream));
    return P2B_OK;
}

static int host_batch(Ctx *c, int g2, const uint8_t *in, uint8_t *out, size_t n, const uint8_t *scalars, size_t n_scalars,
                      int in_enc, int out_enc, int flags) {
    if (!in || !out) return ctx_fail(c, P2B_EARG, "null buffer");
    if (n_scalars != 1 && n_scalars != n) return ctx_fail(c, P2B_EARG, "n_scalars must be 1 or n");
    if (!scalars) return ctx_fail(c, P2B_EARG, "null scalars");
    HostJob j;
    memset(&j, 0, sizeof j);
    j.g2 = g2; j.in = in; j.out = out; j.n = n; j.in_enc = in_enc; j.out_enc = out_enc; j.flags = flags;
    if (n_scalars == 1) {
        j.sc.mode = 1;
        if (!read_scalar_be(scalars, j.sc.k)) return ctx_fail(c, P2B_EARG, "scalar not canonical");
    } else {
        j.sc.mode = 0;
        j.host_scalars = scalars;
    }
    int rc = begin_call(c);
    if (rc) return rc;
    if ((rc = run_host_job(c, j))) return rc;
    return ctx_collect_error(c);
}

static int host_recode(Ctx *c, int g2, const uint8_t *in, uint8_t *out, size_t n, int in_enc, int out_enc, int flags) {
    if (!in) return ctx_fail(c, P2B_EARG, "null buffer");      // out may be NULL: decode + checks only, nothing is copied back
    HostJob j;
    memset(&j, 0, sizeof j);
    j.g2 = g2; j.in = in; j.out = out; j.n = n; j.in_enc = in_enc; j.out_enc = out_enc; j.flags = flags;
    j.sc.mode = 3;
    int rc = begin_call(c);
    if (rc) return rc;
    if ((rc = run_host_job(c, j))) return rc;
    return ctx_collect_error(c);
}

static int powers_spec(Ctx *c, ScalarSpec &sc, const uint8_t tau_be[32], const uint8_t *coeff_be, uint64_t start) {
    memset(&sc, 0, sizeof sc);
    sc.mode = 2;
    sc.start = start;
    if (!tau_be || !read_scalar_be(tau_be, sc.tau)) return ctx_fail(c, P2B_EARG, "tau not canonical");
    if (coeff_be) {
        if (!read_scalar_be(coeff_be, sc.coeff)) return ctx_fail(c, P2B_EARG, "coeff not canonical");
    } else {
        memset(sc.coeff, 0, 32);
        sc.coeff[0] = 1;
    }
    return P2B_OK;
}

static int host_batch_powers(Ctx *c, int g2, const uint8_t *in, uint8_t *out, size_t n, const uint8_t tau_be[32],
                             const uint8_t *coeff_be, uint64_t start, int in_enc, int out_enc, int flags) {
    if (!in || !out) return ctx_fail(c, P2B_EARG, "null buffer");
    HostJob j;
    memset(&j, 0, sizeof j);
    j.g2 = g2; j.in = in; j.out = out; j.n = n; j.in_enc = in_enc; j.out_enc = out_enc; j.flags = flags;
    int rc = powers_spec(c, j.sc, tau_be, coeff_be, start);
    if (rc) return rc;
    if ((rc = begin_call(c))) return rc;
    if ((rc = run_host_job(c, j))) return rc;
    return ctx_collect_error(c);
}

static int dev_batch(Ctx *c, int g2, const void *d_in, void *d_out, size_t n, const uint8_t *scalars, size_t n_scalars,
                     int in_enc, int out_enc, int flags) {
    if (!d_in || !d_out || !scalars) return ctx_fail(c, P2B_EARG, "null buffer");
    if (n_scalars != 1 && n_scalars != n) return ctx_fail(c, P2B_EARG, "n_scalars must be 1 or n");
    P2B_CUDA(c, cudaSetDevice(c->device));
    ScalarSpec sc;
    memset(&sc, 0, sizeof sc);
    int rc;
    if (n_scalars == 1) {
        sc.mode = 1;
        if (!read_scalar_be(scalars, sc.k)) return ctx_fail(c, P2B_EARG, "scalar not canonical");
    } else {
        sc.mode = 0;
        if ((rc = dev_reserve(c, c->scal, n * 32))) return rc;
        P2B_CUDA(c, cudaMemcpyAsync(c->scal.p, scalars, n * 32, cudaMemcpyHostToDevice, c->stream));
        sc.d_scalars = c->scal.p;
    }
    return launch_batch_mul(c, g2, d_in, d_out, n, sc, in_enc, out_enc, flags, 0);
}
static int dev_batch_powers(Ctx *c, int g2, const void *d_in, void *d_out, size_t n, const uint8_t tau_be[32],
                            const uint8_t *coeff_be, uint64_t start, int in_enc, int out_enc, int flags) {
    if (!d_in || !d_out) return ctx_fail(c, P2B_EARG, "null buffer");
    P2B_CUDA(c, cudaSetDevice(c->device));
    ScalarSpec sc;
    int rc = powers_spec(c, sc, tau_be, coeff_be, start);
    if (rc) return rc;
    return launch_batch_mul(c, g2, d_in, d_out, n, sc, in_enc, out_enc, flags, 0);
}

// ------------------------------------------------------------------------------------------------- phase 1
static uint64_t acc_size(uint32_t size_log2, int compressed) {
    uint64_t p = 1ull << size_log2, pg1 = 2 * p - 1, s1 = compressed ? 32 : 64, s2 = compressed ? 64 : 128;
    return pg1 * s1 + p * s2 + 2 * p * s1 + s2 + 64;
}
struct Section {
    int g2;
    uint64_t count;
    int coeff;   // 0 none, 1 alpha, 2 beta
};
static void shard_range(uint64_t count, uint32_t idx, uint32_t cnt, uint64_t &lo, uint64_t &hi) {
    uint64_t per = count / cnt, rem = count % cnt;
    lo = per * idx + (idx < rem ? idx : rem);
    hi = lo + per + (idx < rem ? 1 : 0);
}

// recode: no scalar multiplication, the five sections are only re-encoded (BatchedAccumulator::decompress)
static int pot_transform(Ctx *c, const uint8_t *challenge, uint64_t challenge_len, uint8_t *response, uint64_t response_len,
                         uint32_t size_log2, uint32_t batch_size, int in_c, int out_c, int check, const uint8_t *tau,
                         const uint8_t *alpha, const uint8_t *beta, uint32_t shard_index, uint32_t shard_count,
                         bool recode = false) {
    if (!challenge || !response || (!recode && (!tau || !alpha || !beta))) return ctx_fail(c, P2B_EARG, "null argument");
    if (size_log2 == 0 || size_log2 > 28) return ctx_fail(c, P2B_EARG, "size_log2 out of range");
    if (batch_size == 0) return ctx_fail(c, P2B_EARG, "batch_size must be positive");
    if (shard_count == 0 || shard_index >= shard_count) return ctx_fail(c, P2B_EARG, "bad shard");
    if (challenge_len < acc_size(size_log2, in_c))
        return ctx_fail(c, P2B_EARG, "The size of challenge file should be accumulator_size");
    if (response_len < acc_size(size_log2, out_c)) return ctx_fail(c, P2B_EARG, "response buffer too small");
    const uint64_t powers = 1ull << size_log2, powers_g1 = 2 * powers - 1;
    const uint64_t g1i = in_c ? 32 : 64, g2i = in_c ? 64 : 128, g1o = out_c ? 32 : 64, g2o = out_c ? 64 : 128;
    // section order in both files: TauG1, TauG2, AlphaG1, BetaG1, BetaG2 (batched_accumulator.rs:87-94,96-178)
    const Section sections[4] = {{0, powers_g1, 0}, {1, powers, 0}, {0, powers, 1}, {0, powers, 2}};
    const int flags = ((check & 1) ? P2B_CHECK_INPUT : 0) | P2B_REJECT_INFINITY | (check & (P2B_G2_SUBGROUP | P2B_G2_EXACT));
    int rc = begin_call(c);
    if (rc) return rc;
    uint64_t ioff = 64, ooff = 64;
    for (int s = 0; s < 4; s++) {
        const Section &sec = sections[s];
        const uint64_t isz = sec.g2 ? g2i : g1i, osz = sec.g2 ? g2o : g1o;
        uint64_t lo, hi;
        shard_range(sec.count, shard_index, shard_count, lo, hi);
        HostJob j;
        memset(&j, 0, sizeof j);
        j.g2 = sec.g2;
        j.in = challenge + ioff + lo * isz;
        j.out = response + ooff + lo * osz;
        j.n = hi - lo;
        j.in_enc = in_c ? P2B_ENC_COMPRESSED : P2B_ENC_UNCOMPRESSED;
        j.out_enc = out_c ? P2B_ENC_COMPRESSED : P2B_ENC_UNCOMPRESSED;
        j.flags = flags;
        if (recode) j.sc.mode = 3;
        else if ((rc = powers_spec(c, j.sc, tau, sec.coeff == 1 ? alpha : sec.coeff == 2 ? beta : nullptr, lo))) return rc;
        if ((rc = run_host_job(c, j))) return rc;
        if ((rc = ctx_collect_error(c))) return rc;   // per section, so the index is section-relative like the reference's
        ioff += sec.count * isz;
        ooff += sec.count * osz;
    }
    if (shard_index == 0) {   // beta_g2 = beta_g2.mul(beta) (batched_accumulator.rs:1230-1234)
        HostJob j;
        memset(&j, 0, sizeof j);
        j.g2 = 1; j.in = challenge + ioff; j.out = response + ooff; j.n = 1;
        j.in_enc = in_c ? P2B_ENC_COMPRESSED : P2B_ENC_UNCOMPRESSED;
        j.out_enc = out_c ? P2B_ENC_COMPRESSED : P2B_ENC_UNCOMPRESSED;
        j.flags = flags;
        j.sc.mode = recode ? 3 : 1;
        if (!recode && !read_scalar_be(beta, j.sc.k)) return ctx_fail(c, P2B_EARG, "beta not canonical");
        if ((rc = run_host_job(c, j))) return rc;
        if ((rc = ctx_collect_error(c))) return rc;
    }
    return P2B_OK;
}

// ------------------------------------------------------------------------------------------------- phase 2
struct ParamsLayout {
    uint64_t delta_g1, delta_g2, h_off, h_n, l_off, l_n, cs_hash, contrib_count_off, contrib_off, contrib_n, total;
};
static uint32_t rd_u32be(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
static void wr_u32be(uint8_t *p, uint32_t v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; }
// Parameters::write / MPCParameters::write order: bellman/src/groth16/mod.rs:141-158,252-285; phase2/src/parameters.rs:663-677
static int params_layout(Ctx *c, const uint8_t *buf, uint64_t len, ParamsLayout &L) {
    uint64_t off = 0;
    auto need = [&](uint64_t n) { return off + n <= len; };
    if (!need(64 + 64 + 128 + 128 + 64 + 128)) return ctx_fail(c, P2B_EARG, "params truncated (vk)");
    off = 64 + 64 + 128 + 128;
    L.delta_g1 = off; off += 64;
    L.delta_g2 = off; off += 128;
    const uint64_t sizes[6] = {64, 64, 64, 64, 64, 128};   // ic, h, l, a, b_g1, b_g2
    for (int v = 0; v < 6; v++) {
        if (!need(4)) return ctx_fail(c, P2B_EARG, "params truncated (vector length)");
        uint64_t n = rd_u32be(buf + off);
        off += 4;
        if (!need(n * sizes[v])) return ctx_fail(c, P2B_EARG, "params truncated (vector body)");
        if (v == 1) { L.h_off = off; L.h_n = n; }
        if (v == 2) { L.l_off = off; L.l_n = n; }
        off += n * sizes[v];
    }
    if (!need(64 + 4)) return ctx_fail(c, P2B_EARG, "params truncated (cs_hash)");
    L.cs_hash = off; off += 64;
    L.contrib_count_off = off;
    L.contrib_n = rd_u32be(buf + off);
    off += 4;
    L.contrib_off = off;
    if (!need(L.contrib_n * 384)) return ctx_fail(c, P2B_EARG, "params truncated (contributions)");
    off += L.contrib_n * 384;
    if (off != len) return ctx_fail(c, P2B_EARG, "trailing bytes after MPCParameters");
    L.total = off;
    return P2B_OK;
}

// one point times one scalar through the GPU path (host in, host out)
static int single_mul(Ctx *c, int g2, const uint8_t *in, const uint8_t *k_be, uint8_t *out) {
    HostJob j;
    memset(&j, 0, sizeof j);
    j.g2 = g2; j.in = in; j.out = out; j.n = 1; j.in_enc = P2B_ENC_UNCOMPRESSED; j.out_enc = P2B_ENC_UNCOMPRESSED;
    j.flags = P2B_CHECK_INPUT;
    j.sc.mode = 1;
    if (!read_scalar_be(k_be, j.sc.k)) return ctx_fail(c, P2B_EARG, "scalar not canonical");
    int rc = run_host_job(c, j);
    if (rc) return rc;
    return ctx_collect_error(c);
}

// H(cs_hash | previous pubkeys | s | s_delta)  (phase2/src/parameters.rs:872-885)
static int phase2_transcript(Ctx *c, const uint8_t *params, uint64_t len, const uint8_t *delta, const uint8_t *s,
                             uint8_t *s_delta_out, uint8_t transcript[64]) {
    ParamsLayout L;
    int rc = params_layout(c, params, len, L);
    if (rc) return rc;
    if ((rc = begin_call(c))) return rc;
    uint8_t s_delta[64];
    if ((rc = single_mul(c, 0, s, delta, s_delta))) return rc;
    Blake2b h;
    h.update(params + L.cs_hash, 64);
    h.update(params + L.contrib_off, L.contrib_n * 384);
    h.update(s, 64);
    h.update(s_delta, 64);
    h.finish(transcript);
    if (s_delta_out) memcpy(s_delta_out, s_delta, 64);
    return P2B_OK;
}

// shard_index / shard_count: the H and L vectors are split into contiguous ranges (one per rank / GPU, no collective: every rank
// writes a disjoint byte range of the shared output); shard 0 also writes everything else (header, the other vectors, the
// new delta_g1 / delta_g2, the appended public key) and is the one whose hash_out is meaningful for the file -- the hash is
// computed on every shard from the same inputs, so all ranks return the same 64 bytes.
static int phase2_contribute(Ctx *c, const uint8_t *params, uint64_t len, uint8_t *out, uint64_t out_len, const uint8_t *delta,
                             const uint8_t *s, const uint8_t *r_g2, uint8_t hash_out[64], uint32_t shard_index = 0,
                             uint32_t shard_count = 1) {
    if (!params || !out || !delta || !s || !r_g2 || !hash_out) return ctx_fail(c, P2B_EARG, "null argument");
    if (shard_count == 0 || shard_index >= shard_count) return ctx_fail(c, P2B_EARG, "bad shard");
    ParamsLayout L;
    int rc = params_layout(c, params, len, L);
    if (rc) return rc;
    if (out_len < len + 384) return ctx_fail(c, P2B_EARG, "params_out must hold params_len + 384 bytes");
    uint32_t dk[8];
    if (!read_scalar_be(delta, dk)) return ctx_fail(c, P2B_EARG, "delta not canonical");
    // delta^-1 (parameters.rs:498); expect("nonzero")
    Fr d;
    for (int i = 0; i < 8; i++) d.l[i] = dk[i];
    if (is_zero(d)) return ctx_fail(c, P2B_EARG, "delta must be nonzero");
    Fr dinv = from_mont(inv(to_mont(d)));
    uint8_t dinv_be[32];
    for (int i = 0; i < 8; i++) {
        uint32_t w = dinv.l[7 - i];
        dinv_be[4 * i] = w >> 24; dinv_be[4 * i + 1] = w >> 16; dinv_be[4 * i + 2] = w >> 8; dinv_be[4 * i + 3] = w;
    }
    // keypair (parameters.rs:860-908) with the RNG-derived s and r supplied by the caller
    uint8_t pubkey[384];
    uint8_t *pk_delta_after = pubkey, *pk_s = pubkey + 64, *pk_s_delta = pubkey + 128, *pk_r_delta = pubkey + 192,
            *pk_transcript = pubkey + 320;
    memcpy(pk_s, s, 64);
    if ((rc = begin_call(c))) return rc;
    // l and h scaled by delta^-1 (parameters.rs:499-505); infinity tolerated (no assert in the phase-2 batch_exp).  All GPU
    // work of the call is queued first and collected once: the two big jobs, then the four single multiplications by delta
    // (s, delta_g1 in G1; r, delta_g2 in G2) as two 2-point jobs.
    const struct { uint64_t off, n; } vecs[2] = {{L.l_off, L.l_n}, {L.h_off, L.h_n}};
    for (int v = 0; v < 2; v++) {
        HostJob j;
        memset(&j, 0, sizeof j);
        uint64_t lo, hi;
        shard_range(vecs[v].n, shard_index, shard_count, lo, hi);
        j.g2 = 0; j.in = params + vecs[v].off + lo * 64; j.out = out + vecs[v].off + lo * 64; j.n = hi - lo;
        j.in_enc = P2B_ENC_UNCOMPRESSED; j.out_enc = P2B_ENC_UNCOMPRESSED; j.flags = 0;
        j.sc.mode = 1;
        memcpy(j.sc.k, dinv.l, 32);
        if ((rc = run_host_job(c, j))) return rc;
    }
    uint8_t g1_in[128], g1_out[128], g2_in[256], g2_out[256];
    memcpy(g1_in, s, 64); memcpy(g1_in + 64, params + L.delta_g1, 64);
    memcpy(g2_in, r_g2, 128); memcpy(g2_in + 128, params + L.delta_g2, 128);
    for (int g2 = 0; g2 < 2; g2++) {
        HostJob j;
        memset(&j, 0, sizeof j);
        j.g2 = g2; j.in = g2 ? g2_in : g1_in; j.out = g2 ? g2_out : g1_out; j.n = 2;
        j.in_enc = P2B_ENC_UNCOMPRESSED; j.out_enc = P2B_ENC_UNCOMPRESSED; j.flags = P2B_CHECK_INPUT;
        j.sc.mode = 1;
        memcpy(j.sc.k, dk, 32);
        if ((rc = run_host_job(c, j))) return rc;
    }
    // everything that does not change is copied through while the GPU works (h and l are rewritten by the jobs above)
    if (out != params && shard_index == 0) {
        const uint64_t h_end = L.h_off + L.h_n * 64, l_end = L.l_off + L.l_n * 64;
        memcpy(out, params, L.h_off);
        memcpy(out + h_end, params + h_end, L.l_off - h_end);
        memcpy(out + l_end, params + l_end, len - l_end);
    }
    if ((rc = ctx_collect_error(c))) return rc;
    // keypair (parameters.rs:860-908): s_delta, the transcript hash H(cs_hash | contributions | s | s_delta), r_delta, delta_after
    memcpy(pk_s_delta, g1_out, 64);
    memcpy(pk_delta_after, g1_out + 64, 64);
    memcpy(pk_r_delta, g2_out, 128);
    {
        Blake2b h;
        h.update(params + L.cs_hash, 64);
        h.update(params + L.contrib_off, L.contrib_n * 384);
        h.update(s, 64);
        h.update(pk_s_delta, 64);
        h.finish(pk_transcript);
    }
    if (shard_index == 0) {
        // vk.delta_g1 / vk.delta_g2 *= delta (parameters.rs:507-508)
        memcpy(out + L.delta_g1, pk_delta_after, 64);
        memcpy(out + L.delta_g2, g2_out + 128, 128);
        // contributions.push(pubkey)
        wr_u32be(out + L.contrib_count_off, (uint32_t)(L.contrib_n + 1));
        memcpy(out + len, pubkey, 384);
    }
    Blake2b::hash(pubkey, 384, hash_out);
    return P2B_OK;
}

}  // namespace p2b

using namespace p2b;

extern "C" {

const char *p2b_version(void) { return "p2b 0.1 (sm_100a)"; }

int p2b_init(int device, p2b_ctx **out) { P2B_RANGE("p2b_init");
    if (!out) return P2B_EARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return P2B_ECUDA;
    if (cudaSetDevice(device) != cudaSuccess) return P2B_ECUDA;
    if (const char *e = getenv("P2B_L2_FETCH")) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(e));   // tuning hook
    p2b_ctx *h = new p2b_ctx();
    Ctx *c = &h->c;
    c->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
    bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&c->copy_in, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&c->copy_out, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithPriority(&c->sort_stream, cudaStreamNonBlocking, -1) == cudaSuccess &&
              cudaMalloc(&c->d_err, sizeof(unsigned long long)) == cudaSuccess &&
              cudaMallocHost(&c->h_err, sizeof(unsigned long long)) == cudaSuccess;
    for (int i = 0; ok && i < 8; i++) ok = cudaEventCreateWithFlags(&c->ev[i], cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; ok && i < 4; i++) ok = cudaEventCreateWithFlags(&c->msm_ev[i], cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; ok && i < 2; i++) ok = cudaEventCreate(&c->h2d_ev[i]) == cudaSuccess;
    if (ok) ok = cudaMemsetAsync(c->d_err, 0xff, sizeof(unsigned long long), c->stream) == cudaSuccess;
    if (!ok) { p2b_destroy(h); return P2B_ECUDA; }
    *out = h;
    return P2B_OK;
}

void p2b_destroy(p2b_ctx *h) { P2B_RANGE("p2b_destroy");
    if (!h) return;
    Ctx *c = &h->c;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    io_destroy(c);
    DevBuf *bufs[] = {&c->jac, &c->prefix, &c->stage_in[0], &c->stage_in[1], &c->stage_out[0], &c->stage_out[1], &c->scal,
                      &c->tables, &c->misc, &c->msm_a, &c->msm_b, &c->msm_c, &c->msm_d, &c->msm_e, &c->msm_f, &c->fft_tw_dir[0], &c->fft_tw_dir[1], &c->gtable, &c->gfft, &c->probe};
    for (DevBuf *b : bufs) if (b->p) cudaFree(b->p);
    for (auto &sl : c->prof_slot)
        for (auto &e : sl.ev) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    if (c->d_err) cudaFree(c->d_err);
    if (c->h_err) cudaFreeHost(c->h_err);
    for (int i = 0; i < 8; i++) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    for (int i = 0; i < 4; i++) if (c->msm_ev[i]) cudaEventDestroy(c->msm_ev[i]);
    for (int i = 0; i < 2; i++) if (c->h2d_ev[i]) cudaEventDestroy(c->h2d_ev[i]);
    if (c->sort_stream) cudaStreamDestroy(c->sort_stream);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->copy_in) cudaStreamDestroy(c->copy_in);
    if (c->copy_out) cudaStreamDestroy(c->copy_out);
    delete h;
}

const char *p2b_last_error(p2b_ctx *h) { return h ? h->c.last_error.c_str() : "null ctx"; }
void p2b_error_detail(p2b_ctx *h, uint64_t *index, int *sub) {
    if (!h) return;
    if (index) *index = h->c.err_index;
    if (sub) *sub = h->c.err_sub;
}
void *p2b_stream(p2b_ctx *h) { return h ? (void *)h->c.stream : nullptr; }
uint64_t p2b_launch_count(p2b_ctx *h) { return h ? h->c.launches : 0; }
int p2b_g2_probe_stats(p2b_ctx *h, uint64_t *probes, int *last_verdict) {
    if (!h) return P2B_EARG;
    Ctx *c = &h->c;
    if (probes) *probes = c->probes;
    if (last_verdict) {
        *last_verdict = -1;
        if (c->probes && c->probe.p) {
            uint32_t v = 0;
            P2B_CUDA(c, cudaSetDevice(c->device));
            P2B_CUDA(c, cudaStreamSynchronize(c->stream));
            P2B_CUDA(c, cudaMemcpy(&v, c->probe.p, 4, cudaMemcpyDeviceToHost));
            *last_verdict = v ? 1 : 0;
        }
    }
    return P2B_OK;
}

void p2b_io_stats(p2b_ctx *h, uint64_t *staged_in, uint64_t *staged_out) {
    if (h) io_stats(&h->c, staged_in, staged_out);
}

int p2b_profile_enable(p2b_ctx *h, int on) {
    if (!h) return P2B_EARG;
    h->c.prof = on != 0;
    for (auto &sl : h->c.prof_slot) { sl.used = 0; sl.kernels = 0; }
    return P2B_OK;
}
int p2b_profile_read(p2b_ctx *h, int slot, double *total_ms, uint64_t *kernels) {
    if (!h || slot < 0 || slot >= P2B_PROF_SLOTS) return P2B_EARG;
    Ctx *c = &h->c;
    P2B_CUDA(c, cudaSetDevice(c->device));
    P2B_CUDA(c, cudaStreamSynchronize(c->stream));
    double t = 0;
    Ctx::ProfSlot &s = c->prof_slot[slot];
    for (size_t i = 0; i < s.used; i++) {
        float ms = 0;
        P2B_CUDA(c, cudaEventElapsedTime(&ms, s.ev[i].first, s.ev[i].second));
        t += ms;
    }
    if (total_ms) *total_ms = t;
    if (kernels) *kernels = s.kernels;
    return P2B_OK;
}

int p2b_g1_batch_mul(p2b_ctx *h, const uint8_t *in, uint8_t *out, size_t n, const uint8_t *s, size_t ns, int ie, int oe, int fl) { P2B_RANGE("p2b_g1_batch_mul");
    return h ? host_batch(&h->c, 0, in, out, n, s, ns, ie, oe, fl) : P2B_EARG;
}
int p2b_g2_batch_mul(p2b_ctx *h, const uint8_t *in, uint8_t *out, size_t n, const uint8_t *s, size_t ns, int ie, int oe, int fl) { P2B_RANGE("p2b_g2_batch_mul");
    return h ? host_batch(&h->c, 1, in, out, n, s, ns, ie, oe, fl) : P2B_EARG;
}
int p2b_g1_batch_mul_powers(p2b_ctx *h, const uint8_t *in, uint8_t *out, size_t n, const uint8_t tau[32], const uint8_t *coeff,
                            uint64_t start, int ie, int oe, int fl) { P2B_RANGE("p2b_g1_batch_mul_powers");
    return h ? host_batch_powers(&h->c, 0, in, out, n, tau, coeff, start, ie, oe, fl) : P2B_EARG;
}
int p2b_g2_batch_mul_powers(p2b_ctx *h, const uint8_t *in, uint8_t *out, size_t n, const uint8_t tau[32], const uint8_t *coeff,
                            uint64_t start, int ie, int oe, int fl) { P2B_RANGE("p2b_g2_batch_mul_powers");
    return h ? host_batch_powers(&h->c, 1, in, out, n, tau, coeff, start, ie, oe, fl) : P2B_EARG;
}
int p2b_g1_batch_mul_dev(p2b_ctx *h, const void *in, void *out, size_t n, const uint8_t *s, size_t ns, int ie, int oe, int fl) { P2B_RANGE("p2b_g1_batch_mul_dev");
    return h ? dev_batch(&h->c, 0, in, out, n, s, ns, ie, oe, fl) : P2B_EARG;
}
int p2b_g2_batch_mul_dev(p2b_ctx *h, const void *in, void *out, size_t n, const uint8_t *s, size_t ns, int ie, int oe, int fl) { P2B_RANGE("p2b_g2_batch_mul_dev");
    return h ? dev_batch(&h->c, 1, in, out, n, s, ns, ie, oe, fl) : P2B_EARG;
}
int p2b_g1_batch_mul_powers_dev(p2b_ctx *h, const void *in, void *out, size_t n, const uint8_t tau[32], const uint8_t *coeff,
                                uint64_t start, int ie, int oe, int fl) { P2B_RANGE("p2b_g1_batch_mul_powers_dev");
    return h ? dev_batch_powers(&h->c, 0, in, out, n, tau, coeff, start, ie, oe, fl) : P2B_EARG;
}
int p2b_g2_batch_mul_powers_dev(p2b_ctx *h, const void *in, void *out, size_t n, const uint8_t tau[32], const uint8_t *coeff,
                                uint64_t start, int ie, int oe, int fl) { P2B_RANGE("p2b_g2_batch_mul_powers_dev");
    return h ? dev_batch_powers(&h->c, 1, in, out, n, tau, coeff, start, ie, oe, fl) : P2B_EARG;
}
int p2b_sync(p2b_ctx *h) { P2B_RANGE("p2b_sync");
    if (!h) return P2B_EARG;
    int rc = ctx_collect_error(&h->c);
    cudaMemsetAsync(h->c.d_err, 0xff, sizeof(unsigned long long), h->c.stream);
    return rc;
}

uint64_t p2b_pot_accumulator_size(uint32_t size_log2, int compressed) { return acc_size(size_log2, compressed); }

int p2b_pot_transform(p2b_ctx *h, const uint8_t *challenge, uint64_t challenge_len, uint8_t *response, uint64_t response_len,
                      uint32_t size_log2, uint32_t batch_size, int in_compressed, int out_compressed, int check_input,
                      const uint8_t tau[32], const uint8_t alpha[32], const uint8_t beta[32], uint32_t shard_index,
                      uint32_t shard_count) { P2B_RANGE("p2b_pot_transform");
    return h ? pot_transform(&h->c, challenge, challenge_len, response, response_len, size_log2, batch_size, in_compressed,
                             out_compressed, check_input, tau, alpha, beta, shard_index, shard_count)
             : P2B_EARG;
}

int p2b_pot_decompress(p2b_ctx *h, const uint8_t *response, uint64_t response_len, uint8_t *challenge, uint64_t challenge_len,
                       uint32_t size_log2, int check_input, uint32_t shard_index, uint32_t shard_count) { P2B_RANGE("p2b_pot_decompress");
    return h ? pot_transform(&h->c, response, response_len, challenge, challenge_len, size_log2, 1, 1, 0, check_input, nullptr,
                             nullptr, nullptr, shard_index, shard_count, true)
             : P2B_EARG;
}
int p2b_g1_recode(p2b_ctx *h, const uint8_t *in, uint8_t *out, size_t n, int ie, int oe, int fl) { P2B_RANGE("p2b_g1_recode");
    return h ? host_recode(&h->c, 0, in, out, n, ie, oe, fl) : P2B_EARG;
}
int p2b_g2_recode(p2b_ctx *h, const uint8_t *in, uint8_t *out, size_t n, int ie, int oe, int fl) { P2B_RANGE("p2b_g2_recode");
    return h ? host_recode(&h->c, 1, in, out, n, ie, oe, fl) : P2B_EARG;
}

int p2b_phase2_transcript(p2b_ctx *h, const uint8_t *params, uint64_t params_len, const uint8_t delta[32], const uint8_t s[64],
                          uint8_t transcript_out[64]) { P2B_RANGE("p2b_phase2_transcript");
    if (!h || !params || !delta || !s || !transcript_out) return P2B_EARG;
    return phase2_transcript(&h->c, params, params_len, delta, s, nullptr, transcript_out);
}
int p2b_phase2_contribute(p2b_ctx *h, const uint8_t *params, uint64_t params_len, uint8_t *params_out, uint64_t params_out_len,
                          const uint8_t delta[32], const uint8_t s[64], const uint8_t r[128], uint8_t hash_out[64]) { P2B_RANGE("p2b_phase2_contribute");
    return h ? phase2_contribute(&h->c, params, params_len, params_out, params_out_len, delta, s, r, hash_out) : P2B_EARG;
}
int p2b_phase2_contribute_sharded(p2b_ctx *h, const uint8_t *params, uint64_t params_len, uint8_t *params_out, uint64_t params_out_len,
                                  const uint8_t delta[32], const uint8_t s[64], const uint8_t r[128], uint8_t hash_out[64],
                                  uint32_t shard_index, uint32_t shard_count) { P2B_RANGE("p2b_phase2_contribute_sharded");
    return h ? phase2_contribute(&h->c, params, params_len, params_out, params_out_len, delta, s, r, hash_out, shard_index, shard_count)
             : P2B_EARG;
}

}  // extern "C"
