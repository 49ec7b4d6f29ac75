// msm_g1.cu -- G1 instantiation of the Pippenger MSM kernels and the MSM entry points of the C ABI.
#include "msm_impl.cuh"

namespace p2b {
int msm_typed_g2(Ctx *c, const void *d_points, const void *d_scalars, size_t n, uint32_t *d_out_wire, size_t geom_n,
                 int phase, uint64_t err_base, size_t total_n);

// result: uncompressed wire bytes in device memory c->misc (first 128 bytes)
static int msm_run(Ctx *c, int g2, const void *d_points, const void *d_scalars, size_t n, uint8_t *out_host) {
    int rc;
    if ((rc = dev_reserve(c, c->misc, 4096))) return rc;
    uint32_t *d_out = (uint32_t *)c->misc.p;
    rc = g2 ? msm_typed_g2(c, d_points, d_scalars, n, d_out, n, MSM_FIRST | MSM_LAST, 0, 0)
            : msm_typed<Fq>(c, d_points, d_scalars, n, d_out, n, MSM_FIRST | MSM_LAST, 0);
    if (rc) return rc;
    P2B_CUDA(c, cudaMemcpyAsync(out_host, d_out, g2 ? 128 : 64, cudaMemcpyDeviceToHost, c->stream));
    return ctx_collect_error(c);
}

static int msm_begin(Ctx *c) {
    P2B_CUDA(c, cudaSetDevice(c->device));
    c->last_error.clear();
    P2B_CUDA(c, cudaMemsetAsync(c->d_err, 0xff, sizeof(unsigned long long), c->stream));
    return P2B_OK;
}

// Host buffers larger than STREAM_MIN terms are streamed: chunks of up to STREAM_CHUNK terms are copied on the H2D stream into
// a double-buffered staging area while the previous chunk is sorted and accumulated into the SAME buckets; the bucket
// reduction runs once at the end.  The sum is unchanged (bucket contents are sums of the same terms).
static constexpr size_t MSM_STREAM_MIN = (size_t)1 << 23, MSM_STREAM_CHUNK = (size_t)1 << 24;
static size_t msm_stream_chunk() {       // P2B_MSM_STREAM_CHUNK=<terms>: test hook to exercise the streamed path at small sizes
    const char *e = getenv("P2B_MSM_STREAM_CHUNK");
    long v = e ? atol(e) : 0;
    return v > 0 ? (size_t)v : 0;
}
static int msm_host_streamed(Ctx *c, int g2, const uint8_t *points, const uint8_t *scalars, size_t n, uint8_t *out) {
    const size_t ov = msm_stream_chunk();
    const size_t psz = g2 ? 128 : 64, chunk = ov ? ov : MSM_STREAM_CHUNK;
    int rc;
    for (int b = 0; b < 2; b++)
        if ((rc = dev_reserve(c, c->stage_in[b], chunk * (psz + 32)))) return rc;
    if ((rc = dev_reserve(c, c->misc, 4096))) return rc;
    uint32_t *d_out = (uint32_t *)c->misc.p;
    cudaEvent_t *ev_in = c->ev, *ev_done = c->ev + 2;
    P2B_CUDA(c, cudaEventRecord(c->ev[6], c->stream));
    // chunk sizes double from 1/8 of the maximum (1/8, 1/8, 1/4, 1/2, 1, 1, ..): the GPU starts after a short copy, and
    // every later copy is hidden behind the work on the terms already on the device
    size_t off = 0, next = chunk / 8 ? chunk / 8 : 1;
    for (size_t ci = 0; off < n; ci++) {
        const int b = (int)(ci & 1);
        size_t m = next;
        if (ci >= 1 && next < chunk) next = next * 2 < chunk ? next * 2 : chunk;
        if (m > n - off) m = n - off;
        P2B_CUDA(c, cudaStreamWaitEvent(c->copy_in, ci >= 2 ? ev_done[b] : c->ev[6], 0));
        char *d_pts = (char *)c->stage_in[b].p, *d_sc = d_pts + chunk * psz;
        if ((rc = io_h2d(c, d_pts, points + off * psz, m * psz, c->copy_in))) return rc;
        if ((rc = io_h2d(c, d_sc, scalars + off * 32, m * 32, c->copy_in))) return rc;
        P2B_CUDA(c, cudaEventRecord(ev_in[b], c->copy_in));
        P2B_CUDA(c, cudaStreamWaitEvent(c->stream, ev_in[b], 0));
        const int phase = (ci == 0 ? MSM_FIRST : 0) | (off + m == n ? MSM_LAST : 0);
        rc = g2 ? msm_typed_g2(c, d_pts, d_sc, m, d_out, chunk, phase, off, n) : msm_typed<Fq>(c, d_pts, d_sc, m, d_out, chunk, phase, off, n);
        if (rc) return rc;
        P2B_CUDA(c, cudaEventRecord(ev_done[b], c->stream));
        off += m;
    }
    P2B_CUDA(c, cudaMemcpyAsync(out, d_out, psz, cudaMemcpyDeviceToHost, c->stream));
    return ctx_collect_error(c);
}

static int msm_host(Ctx *c, int g2, const uint8_t *points, const uint8_t *scalars, size_t n, uint8_t *out) {
    if (!out || (n && (!points || !scalars))) return ctx_fail(c, P2B_EARG, "null buffer");
    int rc = msm_begin(c);
    if (rc) return rc;
    if (n > (msm_stream_chunk() ? msm_stream_chunk() : MSM_STREAM_MIN)) return msm_host_streamed(c, g2, points, scalars, n, out);
    const size_t psz = g2 ? 128 : 64;
    if ((rc = dev_reserve(c, c->stage_in[0], (n ? n : 1) * psz))) return rc;
    if ((rc = dev_reserve(c, c->stage_in[1], (n ? n : 1) * 32))) return rc;
    if (n) {
        if ((rc = io_h2d(c, c->stage_in[0].p, points, n * psz, c->stream))) return rc;
        if ((rc = io_h2d(c, c->stage_in[1].p, scalars, n * 32, c->stream))) return rc;
    }
    return msm_run(c, g2, c->stage_in[0].p, c->stage_in[1].p, n, out);
}
static int msm_dev(Ctx *c, int g2, const void *d_points, const void *d_scalars, size_t n, uint8_t *out) {
    if (!out || (n && (!d_points || !d_scalars))) return ctx_fail(c, P2B_EARG, "null buffer");
    int rc = msm_begin(c);
    if (rc) return rc;
    return msm_run(c, g2, d_points, d_scalars, n, out);
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

}  // namespace p2b

using namespace p2b;
extern "C" {
int p2b_g1_msm(p2b_ctx *h, const uint8_t *points, const uint8_t *scalars, size_t n, uint8_t *out) {
    return h ? msm_host(&h->c, 0, points, scalars, n, out) : P2B_EARG;
}
int p2b_g2_msm(p2b_ctx *h, const uint8_t *points, const uint8_t *scalars, size_t n, uint8_t *out) {
    return h ? msm_host(&h->c, 1, points, scalars, n, out) : P2B_EARG;
}
int p2b_g1_msm_dev(p2b_ctx *h, const void *d_points, const void *d_scalars, size_t n, uint8_t *out) {
    return h ? msm_dev(&h->c, 0, d_points, d_scalars, n, out) : P2B_EARG;
}
int p2b_g2_msm_dev(p2b_ctx *h, const void *d_points, const void *d_scalars, size_t n, uint8_t *out) {
    return h ? msm_dev(&h->c, 1, d_points, d_scalars, n, out) : P2B_EARG;
}
int p2b_g1_sum_points(p2b_ctx *h, const uint8_t *points, size_t count, uint8_t out[64]) {
    return h ? sum_points(&h->c, 0, points, count, out) : P2B_EARG;
}
int p2b_g2_sum_points(p2b_ctx *h, const uint8_t *points, size_t count, uint8_t out[128]) {
    return h ? sum_points(&h->c, 1, points, count, out) : P2B_EARG;
}
}
