// msm_g1.cu -- G1 instantiation of the Pippenger MSM kernels and the MSM entry points of the C ABI.
#include <vector>
#include "msm_impl.cuh"

namespace p2b {
int msm_typed_g2(Ctx *c, const MsmJob &j);
static int msm_typed_any(Ctx *c, int g2, const MsmJob &j) { return g2 ? msm_typed_g2(c, j) : msm_typed<Fq>(c, j); }

// One call of the MSM family.  Buffers are HOST buffers unless `dev`.
struct MsmCall {
    int g2 = 0;
    bool dev = false;
    const uint8_t *points = nullptr, *points_b = nullptr;   // points_b: MSM_PAIR_SEPARATE
    const uint8_t *scalars = nullptr;                       // n x 32 B big-endian; nullptr: generated on the device from `seed`
    size_t n = 0;                                           // TERMS (a shifted pair reads n + 1 points)
    int pair = MSM_SINGLE;
    int in_enc = P2B_ENC_UNCOMPRESSED;                      // wire encoding of the points (uncompressed / compressed)
    int flags = 0;                                          // P2B_CHECK_INPUT, P2B_REJECT_INFINITY
    const uint8_t *seed = nullptr;
    uint32_t scalar_bits = 0;
    uint8_t *out_a = nullptr, *out_b = nullptr;
};

static int msm_begin(Ctx *c) {
    P2B_CUDA(c, cudaSetDevice(c->device));
    c->last_error.clear();
    P2B_CUDA(c, cudaMemsetAsync(c->d_err, 0xff, sizeof(unsigned long long), c->stream));
    return P2B_OK;
}
static ChaKey cha_key(const uint8_t *seed) {
    ChaKey k;
    for (int i = 0; i < 8; i++) k.k[i] = (uint32_t)seed[4 * i] | ((uint32_t)seed[4 * i + 1] << 8) | ((uint32_t)seed[4 * i + 2] << 16) | ((uint32_t)seed[4 * i + 3] << 24);
    return k;
}
static int random_scalars_dev(Ctx *c, uint32_t *d_scalars, size_t n, uint64_t first, const uint8_t *seed, uint32_t bits, cudaStream_t s) {
    if (!n) return P2B_OK;
    int grid = (int)((n / 2 + 256) / 256);
    if (grid > c->sm_count * 16) grid = c->sm_count * 16;
    k_msm_random_scalars<<<grid, 256, 0, s>>>(d_scalars, n, first, cha_key(seed), bits);
    c->launches++;
    P2B_CUDA(c, cudaGetLastError());
    return P2B_OK;
}

// Host buffers larger than STREAM_MIN terms are streamed: chunks of up to STREAM_CHUNK terms are copied on the H2D stream into
// a double-buffered staging area while the previous chunk is sorted and accumulated into the SAME buckets; the bucket
// reduction runs once at the end.  The sum is unchanged (bucket contents are sums of the same terms).
// Chunk size: 2^22 terms.  Measured at 2^26 terms (tools/e2e_probe.py): maximum chunk 2^25 189.5 ms, 2^24 183.5, 2^23 181.1, 2^22 177.1,
// 2^21 178.1, 2^20 186.5, 2^19 222 -- small chunks keep the points a chunk's accumulation gathers (2^22 x 64 B = 268 MB) largely in
// the 126 MB L2 across the 14 windows, which outweighs the per-chunk sort and the reload of the bucket accumulators; the streamed
// MSM then runs as fast as the device-resident one, and device-resident inputs are tiled the same way (P2B_MSM_TILE).
// G2 does not gain from it (its gathers are whole 128-byte lines and a term costs 3.5x the arithmetic; measured at 2^25: untiled
// 301 ms, tiles of 2^20 / 2^21 / 2^22 terms 311 / 311 / 308 ms): device-resident G2 inputs are not tiled, host buffers stream in
// 2^24-term chunks.
static constexpr size_t MSM_STREAM_MIN = (size_t)1 << 23, MSM_STREAM_CHUNK = (size_t)1 << 22, MSM_STREAM_CHUNK_G2 = (size_t)1 << 24;
static size_t msm_tile() {               // P2B_MSM_TILE=<terms> (0 = no tiling of device-resident inputs): tuning / test hook
    const char *e = getenv("P2B_MSM_TILE");
    return e ? (size_t)atol(e) : MSM_STREAM_CHUNK;
}
// growth of the chunk sizes, in eighths (P2B_MSM_STREAM_GROWTH, tuning): a chunk's copy (1.75 ns per G1 term over PCIe 5) hides
// behind the work on the terms already there (2.8 ns per term) only while the cumulative size grows by <= 1.6x per chunk;
// plain doubling makes the GPU wait for the copy of every chunk from the fourth on
static size_t msm_stream_growth8() {
    static const size_t g = [] { const char *e = getenv("P2B_MSM_STREAM_GROWTH"); long v = e ? atol(e) : 0; return v >= 9 && v <= 32 ? (size_t)v : (size_t)13; }();
    return g;
}
static size_t msm_stream_chunk() {       // P2B_MSM_STREAM_CHUNK=<terms>: test hook to exercise the streamed path at small sizes
    const char *e = getenv("P2B_MSM_STREAM_CHUNK");
    long v = e ? atol(e) : 0;
    return v > 0 ? (size_t)v : 0;
}

static int msm_call(Ctx *c, const MsmCall &a) {
    const bool pair = a.pair != MSM_SINGLE;
    if (!a.out_a || (pair && !a.out_b)) return ctx_fail(c, P2B_EARG, "null output");
    if (a.n && (!a.points || (a.pair == MSM_PAIR_SEPARATE && !a.points_b))) return ctx_fail(c, P2B_EARG, "null buffer");
    if (!a.scalars && a.n) {
        if (!a.seed) return ctx_fail(c, P2B_EARG, "need scalars or a seed for device-generated ones");
        if (a.scalar_bits < 64 || a.scalar_bits > 253) return ctx_fail(c, P2B_EARG, "scalar_bits must be in [64, 253] for device-generated scalars");
    }
    if (a.scalar_bits > 253) return ctx_fail(c, P2B_EARG, "scalar_bits must be <= 253 (0 = any canonical scalar)");
    if (a.in_enc != P2B_ENC_UNCOMPRESSED && a.in_enc != P2B_ENC_COMPRESSED) return ctx_fail(c, P2B_EARG, "bad encoding");
    if (a.dev && a.in_enc != P2B_ENC_UNCOMPRESSED) return ctx_fail(c, P2B_EARG, "device entry points take uncompressed points");
    int rc = msm_begin(c);
    if (rc) return rc;
    const size_t psz = a.g2 ? 128 : 64, isz = enc_size(a.g2, a.in_enc);
    const size_t extra = a.pair == MSM_PAIR_SHIFTED ? 1 : 0;
    const size_t ov = msm_stream_chunk();
    const bool streamed = !a.dev && a.n > (ov ? ov : MSM_STREAM_MIN);
    // device-resident input: same chunks, no copies (above 2^23 terms; above the tile size when the hook overrides it)
    const bool tile_hook = getenv("P2B_MSM_TILE") != nullptr;
    const bool tiled = a.dev && msm_tile() && (tile_hook || !a.g2) && a.n > (tile_hook ? msm_tile() : MSM_STREAM_MIN);
    const bool chunked = streamed || tiled;
    const size_t chunk = streamed ? (ov ? ov : (a.g2 ? MSM_STREAM_CHUNK_G2 : MSM_STREAM_CHUNK)) : tiled ? msm_tile() : (a.n ? a.n : 1);
    if ((rc = dev_reserve(c, c->misc, 4096))) return rc;
    uint32_t *d_out = (uint32_t *)c->misc.p;
    // staging per buffer: points A (chunk + 1), points B (chunk), scalars (chunk)
    const size_t off_b = (chunk + 1) * isz, off_s = off_b + (a.pair == MSM_PAIR_SEPARATE ? chunk * isz : 0);
    const size_t stage_bytes = off_s + chunk * 32;
    const int nbuf = streamed ? 2 : 1;
    if (!a.dev || !a.scalars)
        for (int b = 0; b < nbuf; b++)
            if ((rc = dev_reserve(c, c->stage_in[b], stage_bytes))) return rc;
    const bool compressed = a.in_enc == P2B_ENC_COMPRESSED;
    if (compressed && (rc = dev_reserve(c, c->msm_f, (chunk + 1) * psz * (a.pair == MSM_PAIR_SEPARATE ? 2 : 1)))) return rc;
    cudaEvent_t *ev_in = c->ev, *ev_done = c->ev + 2;
    P2B_CUDA(c, cudaEventRecord(c->ev[6], c->stream));
    cudaStream_t CP = streamed ? c->copy_in : c->stream;
    // Chunk plan.  Sizes grow geometrically from 1/16 of the maximum: the GPU starts after a short copy, and every later copy
    // is hidden behind the work on the terms already on the device.  When the host link is the slower side (several ranks
    // pulling from one host: 23 GB/s per rank measured at 8 GPUs against 55 GB/s alone), the run ends with the COMPUTE of the
    // last chunk, which nothing overlaps -- so the plan then ends with small chunks (1/4, 1/8, 1/8 of the maximum) and uses a
    // smaller maximum.  The link speed is the one measured on this context's previous streamed call (c->h2d_gbps).
    std::vector<size_t> plan;
    if (tiled) {
        for (size_t rem = a.n; rem;) { const size_t m = rem < chunk ? rem : chunk; plan.push_back(m); rem -= m; }
    } else if (!streamed) plan.push_back(a.n);
    else {
        const double ns_copy = c->h2d_gbps > 0 ? (double)(psz * (a.pair == MSM_PAIR_SEPARATE ? 2 : 1) + (a.scalars ? 32 : 0)) / c->h2d_gbps : 0;
        const double ns_compute = (a.g2 ? 10.0 : 2.8) * (pair ? 2.0 : 1.0);      // per term, measured (bench.py, one B200)
        bool copy_bound = ns_copy > 0.85 * ns_compute;
        if (const char *e = getenv("P2B_MSM_TAIL")) copy_bound = atoi(e) != 0;   // tuning / test hook
        const size_t cap = copy_bound && !ov ? chunk / 2 : chunk;
        const size_t tail[3] = {cap / 4, cap / 8, cap / 8};
        const size_t tail_total = copy_bound ? tail[0] + tail[1] + tail[2] : 0;
        size_t rem = a.n, next = chunk / 16 ? chunk / 16 : 1;
        while (rem > tail_total) {
            size_t m = next < rem - tail_total ? next : rem - tail_total;
            plan.push_back(m);
            rem -= m;
            if (plan.size() >= 2 && next < cap) {
                const size_t grown = next * msm_stream_growth8() / 8 + 1;
                next = grown < cap ? grown : cap;
            }
        }
        for (int t = 0; t < 3 && rem; t++) {
            const size_t m = t == 2 || tail[t] > rem ? rem : tail[t];
            plan.push_back(m);
            rem -= m;
        }
    }
    // link-speed probe of the previous streamed call
    if (c->h2d_probe_bytes && cudaEventQuery(c->h2d_ev[1]) == cudaSuccess) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, c->h2d_ev[0], c->h2d_ev[1]) == cudaSuccess && ms > 0) c->h2d_gbps = (double)c->h2d_probe_bytes / (ms * 1e6);
        c->h2d_probe_bytes = 0;
    }
    size_t probe_ci = plan.size();
    if (streamed && !a.dev) {            // time the copies of the largest chunk
        probe_ci = 0;
        for (size_t i = 1; i < plan.size(); i++) if (plan[i] > plan[probe_ci]) probe_ci = i;
    }
    size_t off = 0;
    const int check = ((a.flags & P2B_CHECK_INPUT) ? 1 : 0) | ((a.flags & P2B_REJECT_INFINITY) ? 2 : 0);
    for (size_t ci = 0; ci < plan.size(); ci++) {
        const int b = (int)(ci & 1) % nbuf;
        const size_t m = plan[ci];
        const void *d_pts = nullptr, *d_pts_b = nullptr, *d_sc = nullptr;
        if (a.dev) {
            d_pts = a.points + off * isz;
            d_pts_b = a.points_b ? a.points_b + off * isz : nullptr;
            d_sc = a.scalars ? a.scalars + off * 32 : nullptr;
        } else {
            if (streamed) P2B_CUDA(c, cudaStreamWaitEvent(CP, ci >= 2 ? ev_done[b] : c->ev[6], 0));
            if (ci == probe_ci) P2B_CUDA(c, cudaEventRecord(c->h2d_ev[0], CP));
            char *st = (char *)c->stage_in[b].p;
            if ((m + extra) && (rc = io_h2d(c, st, a.points + off * isz, (m + extra) * isz, CP))) return rc;
            d_pts = st;
            if (a.pair == MSM_PAIR_SEPARATE) {
                if (m && (rc = io_h2d(c, st + off_b, a.points_b + off * isz, m * isz, CP))) return rc;
                d_pts_b = st + off_b;
            }
            if (a.scalars) {
                if (m && (rc = io_h2d(c, st + off_s, a.scalars + off * 32, m * 32, CP))) return rc;
                d_sc = st + off_s;
            }
        }
        if (ci == probe_ci && !a.dev) {
            P2B_CUDA(c, cudaEventRecord(c->h2d_ev[1], CP));
            c->h2d_probe_bytes = (m + extra) * isz + (a.pair == MSM_PAIR_SEPARATE ? m * isz : 0) + (a.scalars ? m * 32 : 0);
        }
        if (!a.scalars) {
            char *st = (char *)c->stage_in[b].p;
            if ((rc = random_scalars_dev(c, (uint32_t *)(st + off_s), m, off, a.seed, a.scalar_bits, CP))) return rc;
            d_sc = st + off_s;
        }
        if (streamed) {
            P2B_CUDA(c, cudaEventRecord(ev_in[b], CP));
            P2B_CUDA(c, cudaStreamWaitEvent(c->stream, ev_in[b], 0));
        }
        MsmJob j;
        j.n = m; j.d_scalars = d_sc; j.d_out_wire = d_out; j.geom_n = chunk; j.err_base = off; j.total_n = chunked ? a.n : 0;
        j.phase = (ci == 0 ? MSM_FIRST : 0) | (off + m == a.n ? MSM_LAST : 0);
        j.pair = a.pair; j.scalar_bits = a.scalar_bits;
        if (compressed) {     // decompress (square root per point) into raw Montgomery form; the MSM's prepare kernel then only copies
            ScalarSpec sc;
            memset(&sc, 0, sizeof sc);
            sc.mode = 3;
            char *raw = (char *)c->msm_f.p;
            if ((m + extra) && (rc = launch_batch_mul(c, a.g2, d_pts, raw, m + extra, sc, P2B_ENC_COMPRESSED, P2B_ENC_RAW_MONT_LE,
                                                      a.flags & P2B_REJECT_INFINITY, off))) return rc;
            j.d_points = raw;
            if (a.pair == MSM_PAIR_SEPARATE) {
                char *raw_b = raw + (chunk + 1) * psz;
                if (m && (rc = launch_batch_mul(c, a.g2, d_pts_b, raw_b, m, sc, P2B_ENC_COMPRESSED, P2B_ENC_RAW_MONT_LE,
                                                a.flags & P2B_REJECT_INFINITY, off))) return rc;
                j.d_points_b = raw_b;
            }
            j.in_enc = ENC_RAW_MONT_LE;
            j.check = 0;
        } else {
            j.d_points = d_pts; j.d_points_b = d_pts_b; j.in_enc = ENC_UNCOMPRESSED; j.check = check;
        }
        if ((rc = msm_typed_any(c, a.g2, j))) return rc;
        if (streamed) P2B_CUDA(c, cudaEventRecord(ev_done[b], c->stream));
        off += m;
    }
    P2B_CUDA(c, cudaMemcpyAsync(a.out_a, d_out, psz, cudaMemcpyDeviceToHost, c->stream));
    if (pair) P2B_CUDA(c, cudaMemcpyAsync(a.out_b, d_out + psz / 4, psz, cudaMemcpyDeviceToHost, c->stream));
    return ctx_collect_error(c);
}

static int sum_points(Ctx *c, int g2, const uint8_t *points, size_t count, uint8_t *out) {
    if (!out || (count && !points)) return ctx_fail(c, P2B_EARG, "null buffer");
    if (count > 65536) return ctx_fail(c, P2B_EARG, "sum_points: too many points");
    int rc = msm_begin(c);
    if (rc) return rc;
    const size_t psz = g2 ? 128 : 64;
    if ((rc = dev_reserve(c, c->misc, 4096 + count * psz))) return rc;
    uint32_t *d_out = (uint32_t *)c->misc.p, *d_in = d_out + 1024;
    if (count) P2B_CUDA(c, cudaMemcpyAsync(d_in, points, count * psz, cudaMemcpyHostToDevice, c->stream));
    if (g2) msm_launch_sum_points_g2(c, d_in, (uint32_t)count, d_out);
    else msm_launch_sum_points_g1(c, d_in, (uint32_t)count, d_out);
    c->launches++;
    P2B_CUDA(c, cudaMemcpyAsync(out, d_out, psz, cudaMemcpyDeviceToHost, c->stream));
    return ctx_collect_error(c);
}

static int random_scalars_host(Ctx *c, const uint8_t *seed, uint64_t first, size_t n, uint32_t bits, uint8_t *out) {
    if (!seed || (n && !out)) return ctx_fail(c, P2B_EARG, "null buffer");
    if (bits < 1 || bits > 253) return ctx_fail(c, P2B_EARG, "scalar_bits must be in [1, 253]");
    int rc = msm_begin(c);
    if (rc) return rc;
    if ((rc = dev_reserve(c, c->stage_in[0], (n ? n : 1) * 32))) return rc;
    if ((rc = random_scalars_dev(c, (uint32_t *)c->stage_in[0].p, n, first, seed, bits, c->stream))) return rc;
    if ((rc = io_d2h(c, out, c->stage_in[0].p, n * 32, c->stream))) return rc;
    return ctx_collect_error(c);
}

}  // namespace p2b

using namespace p2b;
static int single(p2b_ctx *h, int g2, bool dev, const void *points, const void *scalars, size_t n, uint8_t *out) {
    if (!h) return P2B_EARG;
    if (n && !scalars) return ctx_fail(&h->c, P2B_EARG, "null buffer");
    MsmCall a;
    a.g2 = g2; a.dev = dev; a.points = (const uint8_t *)points; a.scalars = (const uint8_t *)scalars; a.n = n; a.out_a = out;
    return msm_call(&h->c, a);
}
static int pair_call(p2b_ctx *h, int g2, const uint8_t *pa, const uint8_t *pb, const uint8_t *scalars, size_t n, const uint8_t *seed,
                     uint32_t bits, int in_enc, int flags, uint8_t *out_a, uint8_t *out_b, bool shifted) {
    if (!h) return P2B_EARG;
    MsmCall a;
    a.g2 = g2; a.points = pa; a.points_b = pb; a.scalars = scalars; a.n = n; a.seed = seed; a.scalar_bits = bits;
    a.in_enc = in_enc; a.flags = flags; a.out_a = out_a; a.out_b = out_b;
    a.pair = shifted ? MSM_PAIR_SHIFTED : MSM_PAIR_SEPARATE;
    return msm_call(&h->c, a);
}
extern "C" {
int p2b_g1_msm(p2b_ctx *h, const uint8_t *points, const uint8_t *scalars, size_t n, uint8_t *out) { P2B_RANGE("p2b_g1_msm"); return single(h, 0, false, points, scalars, n, out); }
int p2b_g2_msm(p2b_ctx *h, const uint8_t *points, const uint8_t *scalars, size_t n, uint8_t *out) { P2B_RANGE("p2b_g2_msm"); return single(h, 1, false, points, scalars, n, out); }
int p2b_g1_msm_dev(p2b_ctx *h, const void *d_points, const void *d_scalars, size_t n, uint8_t *out) { P2B_RANGE("p2b_g1_msm_dev"); return single(h, 0, true, d_points, d_scalars, n, out); }
int p2b_g2_msm_dev(p2b_ctx *h, const void *d_points, const void *d_scalars, size_t n, uint8_t *out) { P2B_RANGE("p2b_g2_msm_dev"); return single(h, 1, true, d_points, d_scalars, n, out); }
int p2b_g1_msm_pair(p2b_ctx *h, const uint8_t *points_a, const uint8_t *points_b, const uint8_t *scalars, size_t n, const uint8_t seed[32],
                    uint32_t scalar_bits, int in_enc, int flags, uint8_t out_a[64], uint8_t out_b[64]) { P2B_RANGE("p2b_g1_msm_pair");
    return pair_call(h, 0, points_a, points_b, scalars, n, seed, scalar_bits, in_enc, flags, out_a, out_b, false);
}
int p2b_g2_msm_pair(p2b_ctx *h, const uint8_t *points_a, const uint8_t *points_b, const uint8_t *scalars, size_t n, const uint8_t seed[32],
                    uint32_t scalar_bits, int in_enc, int flags, uint8_t out_a[128], uint8_t out_b[128]) { P2B_RANGE("p2b_g2_msm_pair");
    return pair_call(h, 1, points_a, points_b, scalars, n, seed, scalar_bits, in_enc, flags, out_a, out_b, false);
}
int p2b_g1_power_pairs(p2b_ctx *h, const uint8_t *points, size_t n_points, const uint8_t *scalars, const uint8_t seed[32], uint32_t scalar_bits,
                       int in_enc, int flags, uint8_t out_a[64], uint8_t out_b[64]) { P2B_RANGE("p2b_g1_power_pairs");
    if (h && n_points < 1) return ctx_fail(&h->c, P2B_EARG, "power_pairs needs at least one point");
    return pair_call(h, 0, points, nullptr, scalars, n_points - 1, seed, scalar_bits, in_enc, flags, out_a, out_b, true);
}
int p2b_g2_power_pairs(p2b_ctx *h, const uint8_t *points, size_t n_points, const uint8_t *scalars, const uint8_t seed[32], uint32_t scalar_bits,
                       int in_enc, int flags, uint8_t out_a[128], uint8_t out_b[128]) { P2B_RANGE("p2b_g2_power_pairs");
    if (h && n_points < 1) return ctx_fail(&h->c, P2B_EARG, "power_pairs needs at least one point");
    return pair_call(h, 1, points, nullptr, scalars, n_points - 1, seed, scalar_bits, in_enc, flags, out_a, out_b, true);
}
int p2b_random_scalars(p2b_ctx *h, const uint8_t seed[32], uint64_t first_index, size_t n, uint32_t scalar_bits, uint8_t *out) { P2B_RANGE("p2b_random_scalars");
    return h ? random_scalars_host(&h->c, seed, first_index, n, scalar_bits, out) : P2B_EARG;
}
int p2b_g1_sum_points(p2b_ctx *h, const uint8_t *points, size_t count, uint8_t out[64]) { P2B_RANGE("p2b_g1_sum_points");
    return h ? sum_points(&h->c, 0, points, count, out) : P2B_EARG;
}
int p2b_g2_sum_points(p2b_ctx *h, const uint8_t *points, size_t count, uint8_t out[128]) { P2B_RANGE("p2b_g2_sum_points");
    return h ? sum_points(&h->c, 1, points, count, out) : P2B_EARG;
}
}
