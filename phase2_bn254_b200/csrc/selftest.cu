// selftest.cu -- element-wise field kernels behind p2b_selftest_field: lets the parity tests drive the device multipliers directly
// (Montgomery product, dedicated squaring, the fused two-product multiplication mont_mul2, the wide product + stand-alone
// reduction) with chosen operands -- 0, 1, p - 1, R mod p, values whose limbs are all ones -- and compare with big integers.
#include "ec.cuh"
#include "p2b_internal.h"

namespace p2b {
// operands and results are RAW limbs (8 x u32 little-endian, Montgomery form is the caller's business); op: 0 a*b, 1 a^2 (dedicated),
// 2 a*b + c*d (mont_mul2), 3 reduce(wide(a, b)), 4 a + b, 5 a - b; field: 0 Fq, 1 Fr
template <class P> __global__ void k_selftest(const uint32_t *a, const uint32_t *b, const uint32_t *c, const uint32_t *d, uint32_t *out, size_t n, int op) {
#if defined(__CUDA_ARCH__)      // the PTX multipliers only exist in the device pass
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        Fp<P> x, y, z, w, r;
        for (int j = 0; j < 8; j++) { x.l[j] = a[8 * i + j]; y.l[j] = b[8 * i + j]; z.l[j] = c[8 * i + j]; w.l[j] = d[8 * i + j]; }
        if (op == 0) r = mul(x, y);
        else if (op == 1) r = sqr_ded(x);
        else if (op == 2) mont_mul2<P>(r.l, x.l, y.l, z.l, w.l);
        else if (op == 3) { uint32_t t[16]; wide_mul(t, x.l, y.l); mont_red<P>(r.l, t); }
        else if (op == 4) r = add(x, y);
        else r = sub(x, y);
        for (int j = 0; j < 8; j++) out[8 * i + j] = r.l[j];
    }
#endif
}
static int selftest(Ctx *c, int field, int op, const uint8_t *a, const uint8_t *b, const uint8_t *cc, const uint8_t *d, size_t n, uint8_t *out) {
    if (!a || !b || !cc || !d || !out || op < 0 || op > 5 || field < 0 || field > 1) return ctx_fail(c, P2B_EARG, "selftest: bad argument");
    P2B_CUDA(c, cudaSetDevice(c->device));
    int rc;
    if ((rc = dev_reserve(c, c->misc, 5 * n * 32 + 64))) return rc;
    uint32_t *da = (uint32_t *)c->misc.p, *db = da + 8 * n, *dc = db + 8 * n, *dd = dc + 8 * n, *dout = dd + 8 * n;
    const uint8_t *src[4] = {a, b, cc, d};
    uint32_t *dst[4] = {da, db, dc, dd};
    for (int k = 0; k < 4; k++) P2B_CUDA(c, cudaMemcpyAsync(dst[k], src[k], n * 32, cudaMemcpyHostToDevice, c->stream));
    const int grid = (int)((n + 127) / 128) < c->sm_count * 8 ? (int)((n + 127) / 128) : c->sm_count * 8;
    if (field == 0) k_selftest<FqP><<<grid ? grid : 1, 128, 0, c->stream>>>(da, db, dc, dd, dout, n, op);
    else k_selftest<FrP><<<grid ? grid : 1, 128, 0, c->stream>>>(da, db, dc, dd, dout, n, op);
    c->launches++;
    P2B_CUDA(c, cudaMemcpyAsync(out, dout, n * 32, cudaMemcpyDeviceToHost, c->stream));
    P2B_CUDA(c, cudaStreamSynchronize(c->stream));
    return P2B_OK;
}
}  // namespace p2b

extern "C" int p2b_selftest_field(p2b_ctx *h, int field, int op, const uint8_t *a, const uint8_t *b, const uint8_t *c, const uint8_t *d,
                                  size_t n, uint8_t *out) {
    return h ? p2b::selftest(&h->c, field, op, a, b, c, d, n, out) : P2B_EARG;
}
