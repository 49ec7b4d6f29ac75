// msm_g2x2.cuh -- G2 bucket accumulation with TWO LANES PER BUCKET: lane 2j holds the c0 components, lane 2j+1 the c1 components of
// every Fq2 value of bucket j's accumulator.
//
// Why: one thread per G2 bucket needs X, Y, ZZ, ZZZ in Fq2 (64 registers) plus the temporaries of a mixed addition in Fq2 and
// of the multiplier, and lands at 255 registers with spills and 8 warps per SM -- 60 % of the multiplier ceiling where the G1
// kernel (160 registers, 12 warps) reaches 86 %; giving the compiler fewer registers (3 blocks per SM) or prefetching only
// adds spills (measured: -10 % / -15 %).  Split over two lanes the state per thread is that of a G1
// accumulator.  An Fq2 product then is, per lane, two products under ONE interleaved Montgomery reduction (mont_mul2, fp.cuh),
//     c0 = (a0 b0 + a1 (-b1)) / R   on the even lane,      c1 = (a0 b1 + a1 b0) / R   on the odd lane,
// with the partner's components fetched by 16 shuffles (schoolbook instead of Karatsuba: 2 x 200 instead of 336 multiply-adds per
// product, the price of perfect balance between the lanes); a squaring is one ordinary Montgomery product per lane,
//     c0 = (a0 + a1)(a0 - a1),   c1 = (2 a0) a1.
// Additions, subtractions and negations are component-wise and need no communication.
#pragma once
#include "xyzz.cuh"

namespace p2b {
#if defined(__CUDA_ARCH__)      // device pass only: built on the PTX helpers of fp.cuh

struct Fq2Half {            // this lane's component of an Fq2 value (c0 on even lanes, c1 on odd lanes)
    Fq v;
};
// Every exchange names only the two lanes of the pair (pm = 3 << (lane & ~1)): the pairs of a warp walk lists of different
// lengths, so a full-warp shuffle mask would wait for lanes that have left the loop.
static __device__ __forceinline__ Fq shfl_partner(const Fq &a, unsigned pm) {
    Fq r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_xor_sync(pm, a.l[i], 1);
    return r;
}
static __device__ __forceinline__ bool both(bool mine, unsigned pm) { return mine & (__shfl_xor_sync(pm, (int)mine, 1) != 0); }
static __device__ __forceinline__ bool either(bool mine, unsigned pm) { return mine | (__shfl_xor_sync(pm, (int)mine, 1) != 0); }

static __device__ __forceinline__ Fq h2_mul(const Fq &a, const Fq &b, bool odd, unsigned pm) {
    const Fq pa = shfl_partner(a, pm), pb = shfl_partner(b, pm);
    // even lane (c0): a * b + pa * (-pb) ;  odd lane (c1): pa * b + a * pb      -- one fused two-product Montgomery multiplication
    const Fq x1 = select(odd, pa, a), x2 = select(odd, a, pa), y2 = select(odd, pb, neg(pb));
    Fq r;
    mont_mul2<FqP>(r.l, x1.l, b.l, x2.l, y2.l);
    return r;
}
static __device__ __forceinline__ Fq h2_sqr(const Fq &a, bool odd, unsigned pm) {
    const Fq pa = shfl_partner(a, pm);
    const Fq x = select(odd, dbl(pa), add(a, pa));       // odd: 2 a0 ; even: a0 + a1
    const Fq y = select(odd, a, sub(a, pa));             // odd: a1   ; even: a0 - a1
    return mul(x, y);
}

struct XyzzHalf { Fq x, y, zz, zzz; };                   // this lane's components of an XYZZ<Fq2> point; infinity <=> zz == 0 on both lanes

static __device__ __forceinline__ XyzzHalf load_xyzz_half(const uint32_t *base, size_t i, int odd) {
    const uint4 *s = reinterpret_cast<const uint4 *>(base + i * 64 + 8 * odd);
    XyzzHalf p;
    Fq *f[4] = {&p.x, &p.y, &p.zz, &p.zzz};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint4 a = s[4 * k], b = s[4 * k + 1];
        f[k]->l[0] = a.x; f[k]->l[1] = a.y; f[k]->l[2] = a.z; f[k]->l[3] = a.w;
        f[k]->l[4] = b.x; f[k]->l[5] = b.y; f[k]->l[6] = b.z; f[k]->l[7] = b.w;
    }
    return p;
}
static __device__ __forceinline__ void store_xyzz_half(uint32_t *base, size_t i, int odd, const XyzzHalf &p) {
    uint4 *d = reinterpret_cast<uint4 *>(base + i * 64 + 8 * odd);
    const Fq *f[4] = {&p.x, &p.y, &p.zz, &p.zzz};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        d[4 * k] = make_uint4(f[k]->l[0], f[k]->l[1], f[k]->l[2], f[k]->l[3]);
        d[4 * k + 1] = make_uint4(f[k]->l[4], f[k]->l[5], f[k]->l[6], f[k]->l[7]);
    }
}

// acc += (+-) the affine point whose components this lane pair loaded; complete like xyzz_madd
static __device__ __forceinline__ void h2_madd(XyzzHalf &p, const Fq &qx, const Fq &qy_in, bool neg_y, bool odd, unsigned pm) {
    const Fq qy = cneg(qy_in, neg_y);
    const Fq u2 = h2_mul(qx, p.zz, odd, pm), s2 = h2_mul(qy, p.zzz, odd, pm);
    const Fq pp_ = sub(u2, p.x), rr = sub(s2, p.y);
    const Fq pp = h2_sqr(pp_, odd, pm), ppp = h2_mul(pp_, pp, odd, pm), qq = h2_mul(p.x, pp, odd, pm);
    XyzzHalf r;
    r.x = sub(sub(sub(h2_sqr(rr, odd, pm), ppp), qq), qq);
    r.y = sub(h2_mul(rr, sub(qq, r.x), odd, pm), h2_mul(p.y, ppp, odd, pm));
    r.zz = h2_mul(p.zz, pp, odd, pm);
    r.zzz = h2_mul(p.zzz, ppp, odd, pm);
    const bool p_inf = both(is_zero(p.zz), pm);
    const bool same = !p_inf & both(is_zero(pp_), pm) & both(is_zero(rr), pm);
    const Fq pqx = shfl_partner(qx, pm), pqy = shfl_partner(qy, pm);      // (exchanged outside the rare branch below)
    if (same) {                                          // P == Q: both lanes double the full point and keep their component
        Aff<Fq2> q;
        q.x.c0 = select(odd, pqx, qx); q.x.c1 = select(odd, qx, pqx);
        q.y.c0 = select(odd, pqy, qy); q.y.c1 = select(odd, qy, pqy);
        const Xyzz<Fq2> d = xyzz_dbl_aff(q);
        r.x = odd ? d.x.c1 : d.x.c0; r.y = odd ? d.y.c1 : d.y.c0; r.zz = odd ? d.zz.c1 : d.zz.c0; r.zzz = odd ? d.zzz.c1 : d.zzz.c0;
    }
    const Fq one = select(odd, fp_zero<FqP>(), fp_one<FqP>());      // 1 = (1, 0)
    p.x = select(p_inf, qx, r.x);
    p.y = select(p_inf, qy, r.y);
    p.zz = select(p_inf, one, r.zz);
    p.zzz = select(p_inf, one, r.zzz);
}

#endif
}  // namespace p2b
