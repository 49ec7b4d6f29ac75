// smul.cuh -- per-point scalar multiplication [k]P for the batch_exp hot loop.
//
// Replaces the body of the reference's `batch_exp` closures
// (powersoftau/src/batched_accumulator.rs:1130-1181, phase2/src/parameters.rs:424-470), which run
// wNAF(4) per point (pairing/src/wnaf.rs:4-71).  The result [k]P is mathematically unique, so the
// GPU is free to use a schedule that suits SIMT execution:
//
//  * every lane executes the SAME add/double schedule: the scalar is recoded into signed ODD
//    fixed-window digits (width 4) -- digit i is read straight off the bits of k, it is never zero,
//    so there is exactly one mixed add per window per lane and no divergence;
//  * G1 uses the BN254 endomorphism phi(x,y) = (beta x, y) = [lambda](x,y): k = k1 + k2*lambda with
//    |k1|,|k2| < 2^128, halving the doubling chain (33 windows instead of 64);
//  * the 8 odd multiples {P, 3P, .., 15P} are computed with mixed adds on the curve isomorphic to E
//    in which 2P is affine, then rescaled to ONE common Z, so all ladder adds are mixed adds
//    (7M + 4S) against an affine table; the common Z is folded back into the result at the end;
//  * the table lives in shared memory, one column per thread (bank-conflict free for any
//    lane-varying digit).
//
// All functions are host+device so tests/host/ can check them against the oracle without a GPU.
#pragma once
#include "ec.cuh"

namespace p2b {

// ------------------------------------------------------------------ small multi-word helpers
// r[0..nr) = a[0..na) * b[0..nb)  (schoolbook, truncating to nr words)
P2B_HD void mp_mul(uint32_t *r, int nr, const uint32_t *a, int na, const uint32_t *b, int nb) {
    for (int i = 0; i < nr; i++) r[i] = 0;
    for (int i = 0; i < na; i++) {
        uint64_t c = 0;
        for (int j = 0; j < nb && i + j < nr; j++) {
            c += (uint64_t)a[i] * b[j] + r[i + j];
            r[i + j] = (uint32_t)c;
            c >>= 32;
        }
        for (int k = i + nb; k < nr && c; k++) { c += r[k]; r[k] = (uint32_t)c; c >>= 32; }
    }
}
P2B_HD void mp_sub(uint32_t *r, const uint32_t *a, const uint32_t *b, int n) {
    uint64_t br = 0;
    for (int i = 0; i < n; i++) { uint64_t t = (uint64_t)a[i] - b[i] - br; r[i] = (uint32_t)t; br = (t >> 63) & 1; }
}
P2B_HD void mp_add(uint32_t *r, const uint32_t *a, const uint32_t *b, int n) {
    uint64_t c = 0;
    for (int i = 0; i < n; i++) { c += (uint64_t)a[i] + b[i]; r[i] = (uint32_t)c; c >>= 32; }
}
// two's complement negate in place
P2B_HD void mp_neg(uint32_t *a, int n) {
    uint64_t c = 1;
    for (int i = 0; i < n; i++) { c += (uint32_t)~a[i]; a[i] = (uint32_t)c; c >>= 32; }
}

// ------------------------------------------------------------------ GLV decomposition
// lambda = 0xb3c4d79d41a917585bfc41088d8daaa78b17ea66b99c90dd, beta = 0x59e26bce...177fffffe (phi(G) == [lambda]G is
// asserted in tests/test_oracle.py against the big-int oracle).  Lattice basis (a1,b1), (a2,b2), a + b*lambda = 0 mod r:
//   a1 = 0x89d3256894d213e3          b1 = -0x6f4d8248eeb859fc8211bbeb7d4f1128
//   a2 = 0x6f4d8248eeb859fd0be4e1541221250b   b2 = 0x89d3256894d213e3
// c1 = floor(k*g1 / 2^256), c2 = floor(k*g2 / 2^256) with g1 = floor(2^256*b2/r), g2 = floor(2^256*(-b1)/r);
// k1 = k - c1*a1 - c2*a2, k2 = c1*(-b1) - c2*b2.  Any integers c1, c2 give k1 + k2*lambda = k (mod r); rounding only
// affects the size: |k1|, |k2| < 2^128 (tests sweep random and extreme k).
struct GlvSplit {
    uint32_t k1[5], k2[5];   // magnitudes (< 2^132)
    bool neg1, neg2;
};
P2B_HD GlvSplit glv_decompose(const uint32_t k[8]) {
    const uint32_t g1[3] = {0xc7e0b3d7u, 0xd91d232eu, 0x2u};
    const uint32_t g2[5] = {0x391eb18du, 0x7a7bd9d4u, 0xa773d2cfu, 0x4ccef014u, 0x2u};
    const uint32_t a1[2] = {0x94d213e3u, 0x89d32568u};
    const uint32_t a2[4] = {0x1221250bu, 0x0be4e154u, 0xeeb859fdu, 0x6f4d8248u};
    const uint32_t nb1[4] = {0x7d4f1128u, 0x8211bbebu, 0xeeb859fcu, 0x6f4d8248u};  // -b1
    const uint32_t b2[2] = {0x94d213e3u, 0x89d32568u};
    uint32_t t[13];
    uint32_t c1[3], c2[5];
    mp_mul(t, 11, k, 8, g1, 3);
    c1[0] = t[8]; c1[1] = t[9]; c1[2] = t[10];
    mp_mul(t, 13, k, 8, g2, 5);
    for (int i = 0; i < 5; i++) c2[i] = t[8 + i];
    // k1 = k - c1*a1 - c2*a2   (mod 2^256, two's complement)
    uint32_t u[8], v[8], k1[8], k2[8];
    mp_mul(u, 8, c1, 3, a1, 2);
    mp_mul(v, 8, c2, 5, a2, 4);
    mp_sub(k1, k, u, 8);
    mp_sub(k1, k1, v, 8);
    // k2 = c1*(-b1) - c2*b2
    mp_mul(u, 8, c1, 3, nb1, 4);
    mp_mul(v, 8, c2, 5, b2, 2);
    mp_sub(k2, u, v, 8);
    GlvSplit s;
    s.neg1 = (k1[7] >> 31) != 0;
    s.neg2 = (k2[7] >> 31) != 0;
    if (s.neg1) mp_neg(k1, 8);
    if (s.neg2) mp_neg(k2, 8);
    for (int i = 0; i < 5; i++) { s.k1[i] = k1[i]; s.k2[i] = k2[i]; }
    return s;
}

// beta (Montgomery form) with phi(x, y) = (beta x, y) = [lambda](x, y)
P2B_DEF_CONST(G1_BETA, {0xd782e155u, 0x71930c11u, 0xffbe3323u, 0xa6bb947cu, 0xd4741444u, 0xaa303344u, 0x26594943u, 0x2c3b3f0du})
P2B_HD Fq g1_beta() { Fq b; for (int i = 0; i < 8; i++) b.l[i] = P2B_C(G1_BETA, i); return b; }
// On the twist the same lambda belongs to the other cube root: (beta^2 x, y) = [lambda](x, y) for points of the order-r
// subgroup of E'(Fq2) (tests/test_oracle.py::test_glv_constants).  beta^2 in Montgomery form:
P2B_DEF_CONST(G2_BETA, {0x13e80b9cu, 0x3350c88eu, 0xdb5e56b9u, 0x7dce557cu, 0xb615564au, 0x6001b4b8u, 0x020217e0u, 0x2682e617u})
P2B_HD Fq g2_beta() { Fq b; for (int i = 0; i < 8; i++) b.l[i] = P2B_C(G2_BETA, i); return b; }
P2B_HD Fq endo_x(const Fq &x) { return mul(x, g1_beta()); }
P2B_HD Fq2 endo_x(const Fq2 &x) { return mul_fq(x, g2_beta()); }

// ------------------------------------------------------------------ table storage policies
// One column per thread: word w of entry e lives at base[(e * WORDS2 + w) * stride].
template <class F> struct StridedTable {
    uint32_t *base;
    int stride;
    static constexpr int W = FieldTraits<F>::WORDS;
    P2B_HD void put(int e, const F &x, const F &y) const {
#pragma unroll
        for (int w = 0; w < W; w++) {
            base[(e * 2 * W + w) * stride] = get_word(x, w);
            base[(e * 2 * W + W + w) * stride] = get_word(y, w);
        }
    }
    P2B_HD Aff<F> get(int e) const {
        Aff<F> r;
#pragma unroll
        for (int w = 0; w < W; w++) {
            set_word(r.x, w, base[(e * 2 * W + w) * stride]);
            set_word(r.y, w, base[(e * 2 * W + W + w) * stride]);
        }
        return r;
    }
};

// One column per thread in GLOBAL memory (L2 resident), 16-byte granules: granule g of entry e lives at
// base[(e * G + g) * stride] with G = 2 * WORDS / 4, so a warp reads 512 contiguous bytes per granule.  Used for G2,
// whose 1 KB table per thread would otherwise cap the SM at 4 warps.
#if defined(__CUDACC__)
template <class F> struct GlobalTable {
    uint4 *base;
    size_t stride;
    static constexpr int W = FieldTraits<F>::WORDS;
    static constexpr int G = 2 * W / 4;
    __device__ __forceinline__ void put(int e, const F &x, const F &y) const {
#pragma unroll
        for (int g = 0; g < G / 2; g++) {
            base[(size_t)(e * G + g) * stride] = make_uint4(get_word(x, 4 * g), get_word(x, 4 * g + 1), get_word(x, 4 * g + 2), get_word(x, 4 * g + 3));
            base[(size_t)(e * G + G / 2 + g) * stride] = make_uint4(get_word(y, 4 * g), get_word(y, 4 * g + 1), get_word(y, 4 * g + 2), get_word(y, 4 * g + 3));
        }
    }
    __device__ __forceinline__ Aff<F> get(int e) const {
        Aff<F> r;
#pragma unroll
        for (int g = 0; g < G / 2; g++) {
            uint4 a = base[(size_t)(e * G + g) * stride], b = base[(size_t)(e * G + G / 2 + g) * stride];
            set_word(r.x, 4 * g, a.x); set_word(r.x, 4 * g + 1, a.y); set_word(r.x, 4 * g + 2, a.z); set_word(r.x, 4 * g + 3, a.w);
            set_word(r.y, 4 * g, b.x); set_word(r.y, 4 * g + 1, b.y); set_word(r.y, 4 * g + 2, b.z); set_word(r.y, 4 * g + 3, b.w);
        }
        return r;
    }
};
#endif

// ------------------------------------------------------------------ odd-multiples table with one common Z
// madd-2007-bl on a point pair known to be distinct and finite, also returning Z3/Z1 = 2H.
// `bad` is raised when H == 0 (P == +-Q: only reachable for points of tiny order or off-curve garbage).
template <class F> P2B_HD Jac<F> madd_ratio(const Jac<F> &p, const Aff<F> &q, F &ratio, bool &bad) {
    F z1z1 = sqr(p.z);
    F u2 = mul(q.x, z1z1);
    F s2 = mul(mul(q.y, p.z), z1z1);
    F h = sub(u2, p.x);
    bad |= is_zero(h);
    F hh = sqr(h);
    F i = dbl(dbl(hh));
    F j = mul(h, i);
    F rr = dbl(sub(s2, p.y));
    F v = mul(p.x, i);
    Jac<F> r;
    r.x = sub(sub(sub(sqr(rr), j), v), v);
    r.y = sub(mul(rr, sub(v, r.x)), dbl(mul(p.y, j)));
    ratio = dbl(h);
    r.z = mul(p.z, ratio);
    return r;
}

// Fills tbl[j] = (2j+1)P, j = 0..7, as AFFINE points of the curve E'' isomorphic to E under
// (x, y) -> (x zg^2, y zg^3); returns zg.  A Jacobian point (X, Y, Z) computed on E'' is (X, Y, Z*zg) on E.
// The a = 0 group-law formulas never touch b, so they are valid on E'' unchanged.
// Cost: affine doubling (1M+5S) + 4 + 7 mixed adds (11) + 7 rescales (5) + 1  ~= 123 field mults.
// `zr` is per-thread scratch for the 8 z-ratios.  `bad` is raised for degenerate inputs (see madd_ratio).
template <class F, class Tbl> P2B_HD F build_odd_table(const Aff<F> &p, const Tbl &tbl, F *zr, bool &bad) {
    Jac<F> d = aff_dbl(p);                     // 2P = (Xd, Yd, Zd) on E; Zd = 2y
    bad |= is_zero(d.z);
    F zd2 = sqr(d.z);
    F zd3 = mul(zd2, d.z);
    Aff<F> d1;                                 // 2P is affine on E' = iso(E, Zd)
    d1.x = d.x; d1.y = d.y;
    Jac<F> t;                                  // P on E'
    t.x = mul(p.x, zd2); t.y = mul(p.y, zd3); t.z = FieldTraits<F>::one();
    tbl.put(0, t.x, t.y);
#pragma unroll 1
    for (int j = 1; j < 8; j++) {              // T_j = T_{j-1} + 2P  (Jacobian on E', Z_j = Z_{j-1} * zr[j])
        t = madd_ratio(t, d1, zr[j], bad);
        tbl.put(j, t.x, t.y);
    }
    F s = FieldTraits<F>::one();               // s = Z_7 / Z_j, walking down
#pragma unroll 1
    for (int j = 6; j >= 0; j--) {
        s = mul(s, zr[j + 1]);
        F s2 = sqr(s);
        Aff<F> e = tbl.get(j);
        tbl.put(j, mul(e.x, s2), mul(e.y, mul(s2, s)));
    }
    return mul(t.z, d.z);                      // zg = Z_7 * Zd
}

// ------------------------------------------------------------------ signed odd fixed-window digits
// For odd k:  k = sum_{i<n} (2 u_i - 15) 16^i + (2 top + 1) 16^n  with u_i = nibble i of (k >> 1) and
// top = k >> (4n + 1).  Every digit is odd and nonzero; |digit| = 2*idx + 1 selects table entry idx.
P2B_HD void digit_from_nibble(uint32_t u, uint32_t &idx, bool &negd) {
    negd = u < 8;
    idx = negd ? 7 - u : u - 8;
}
// m = k >> 1 over n words (k has n words + possibly more); returns nothing, helper for callers
template <int N> P2B_HD void shr1(uint32_t *m, const uint32_t *k, uint32_t hi) {
#pragma unroll
    for (int i = 0; i < N; i++) m[i] = (k[i] >> 1) | ((i + 1 < N ? k[i + 1] : hi) << 31);
}
// take the top nibble of an N-word register file and shift it left by 4
template <int N> P2B_HD uint32_t pop_nibble(uint32_t *m) {
    uint32_t u = m[N - 1] >> 28;
#pragma unroll
    for (int i = N - 1; i > 0; i--) m[i] = (m[i] << 4) | (m[i - 1] >> 28);
    m[0] <<= 4;
    return u;
}

// ------------------------------------------------------------------ generic slow path (complete, any input)
// MSB-first double-and-add with complete formulas: valid for every (x, y), on or off the curve, of any order --
// exactly the group computation the reference's wNAF performs.  Only taken for degenerate / off-curve inputs.
template <class F> P2B_HD Jac<F> mul_binary(const Aff<F> &p, const uint32_t k[8]) {
    Jac<F> acc = jac_infinity<F>();
#pragma unroll 1
    for (int i = 255; i >= 0; i--) {
        acc = jac_dbl(acc);
        if ((k[i >> 5] >> (i & 31)) & 1) acc = jac_madd(acc, p);
    }
    return acc;
}

// ------------------------------------------------------------------ G1: GLV + 2 x 33 signed windows
// k canonical (< r).  p must be on the curve (order r) -- the caller routes anything else to mul_binary.
template <class F, class Tbl> P2B_HD Jac<F> mul_glv(const Aff<F> &p, const uint32_t k[8], const Tbl &tbl, F *zr, bool &bad) {
    GlvSplit s = glv_decompose(k);
    // make both halves odd by adding lattice vectors (a1 odd, b1 even; a2 odd, b2 odd): the represented scalar
    // k1 + k2*lambda (mod r) is unchanged.  Work on signed values: v = sign * magnitude.
    {
        const uint32_t a1[5] = {0x94d213e3u, 0x89d32568u, 0, 0, 0};
        const uint32_t nb1[5] = {0x7d4f1128u, 0x8211bbebu, 0xeeb859fcu, 0x6f4d8248u, 0};  // -b1 > 0
        const uint32_t a2[5] = {0x1221250bu, 0x0be4e154u, 0xeeb859fdu, 0x6f4d8248u, 0};
        const uint32_t b2[5] = {0x94d213e3u, 0x89d32568u, 0, 0, 0};
        bool o1 = s.k1[0] & 1, o2 = s.k2[0] & 1;
        // parity (o1,o2): (0,0) -> +v2 ; (0,1) -> +v1 ; (1,0) -> +v1+v2 ; (1,1) -> nothing
        bool use_v1 = (o1 != o2) | false;       // (0,1) or (1,0)
        bool use_v2 = !o2;                       // (0,0) or (1,0)
        // signed add helper on (mag, neg): to ADD a positive w:  neg ? mag - w (may flip sign) : mag + w
        // to SUBTRACT a positive w (adding b1 = -nb1): neg ? mag + w : mag - w
        uint32_t t[5];
        // k1 += a1 (if use_v1); k1 += a2 (if use_v2)
        // k2 += b1 = -nb1 (if use_v1); k2 += b2 (if use_v2)
        // implement generically with two's complement on 6 words
        uint32_t x1[6], x2[6], w[6];
        for (int i = 0; i < 5; i++) { x1[i] = s.k1[i]; x2[i] = s.k2[i]; }
        x1[5] = 0; x2[5] = 0;
        if (s.neg1) mp_neg(x1, 6);
        if (s.neg2) mp_neg(x2, 6);
        for (int i = 0; i < 5; i++) w[i] = use_v1 ? a1[i] : 0u; w[5] = 0; mp_add(x1, x1, w, 6);
        for (int i = 0; i < 5; i++) w[i] = use_v2 ? a2[i] : 0u; w[5] = 0; mp_add(x1, x1, w, 6);
        for (int i = 0; i < 5; i++) w[i] = use_v1 ? nb1[i] : 0u; w[5] = 0; mp_sub(x2, x2, w, 6);
        for (int i = 0; i < 5; i++) w[i] = use_v2 ? b2[i] : 0u; w[5] = 0; mp_add(x2, x2, w, 6);
        s.neg1 = (x1[5] >> 31) != 0;
        s.neg2 = (x2[5] >> 31) != 0;
        if (s.neg1) mp_neg(x1, 6);
        if (s.neg2) mp_neg(x2, 6);
        for (int i = 0; i < 5; i++) { s.k1[i] = x1[i]; s.k2[i] = x2[i]; }
        (void)t;
    }
    F zg = build_odd_table<F>(p, tbl, zr, bad);
    uint32_t m1[4], m2[4];
    shr1<4>(m1, s.k1, s.k1[4]);
    shr1<4>(m2, s.k2, s.k2[4]);
    uint32_t top1 = s.k1[4] >> 1, top2 = s.k2[4] >> 1;   // k >> 129, must be < 8 (|k| < 2^132)
    bad |= (top1 > 7) | (top2 > 7);
    // top window: acc = T[top1] (+-) , then + phi(T[top2])
    Aff<F> q = tbl.get(top1 & 7);
    q.y = cneg(q.y, s.neg1);
    Jac<F> acc = jac_from_aff(q);
    q = tbl.get(top2 & 7);
    q.x = endo_x(q.x);
    q.y = cneg(q.y, s.neg2);
    acc = jac_madd(acc, q);
#pragma unroll 1
    for (int i = 0; i < 32; i++) {
#pragma unroll 1
        for (int d = 0; d < 4; d++) acc = jac_dbl(acc);
#pragma unroll 1
        for (int h = 0; h < 2; h++) {
            uint32_t u = h ? pop_nibble<4>(m2) : pop_nibble<4>(m1);
            uint32_t idx; bool negd;
            digit_from_nibble(u, idx, negd);
            q = tbl.get(idx);
            if (h) q.x = endo_x(q.x);
            q.y = cneg(q.y, negd != (h ? s.neg2 : s.neg1));
            acc = jac_madd(acc, q);
        }
    }
    acc.z = mul(acc.z, zg);                     // back from E'' to E
    return acc;
}

// ------------------------------------------------------------------ one scalar for every point (phase-2 shape)
// When all lanes multiply by the SAME scalar (MPCParameters::contribute: every H and L point times delta^-1,
// phase2/src/parameters.rs:424-470) the digit pattern is warp-uniform, so zero digits can simply be skipped without
// divergence: the host recodes the two GLV halves into width-5 NAF (odd digits in [-15, 15], on average one non-zero
// digit in six) and the kernel runs 128 doublings + ~43 mixed adds instead of 128 + 66.  Digits are LSB first.
struct UniformDigits {
    int8_t d1[136], d2[136];
    int len;
};
// m: N-word magnitude (< 2^(NDIG - 2)), destroyed; out: NDIG digits
template <int N, int NDIG> P2B_HD int wnaf5_recode(int8_t *out, uint32_t *m) {
    int len = 0;
    for (int i = 0; i < NDIG; i++) {
        int8_t d = 0;
        if (m[0] & 1u) {
            int v = (int)(m[0] & 31u);
            if (v >= 16) v -= 32;
            d = (int8_t)v;
            if (v > 0) {                                               // m -= v
                uint64_t br = (uint64_t)v;
                for (int j = 0; j < N && br; j++) { uint64_t t = (uint64_t)m[j] - br; m[j] = (uint32_t)t; br = (t >> 63) & 1; }
            } else {                                                   // m += -v
                uint64_t c = (uint64_t)(-v);
                for (int j = 0; j < N && c; j++) { c += m[j]; m[j] = (uint32_t)c; c >>= 32; }
            }
        }
        out[i] = d;
        if (d) len = i + 1;
        for (int j = 0; j < N; j++) m[j] = (m[j] >> 1) | (j + 1 < N ? m[j + 1] << 31 : 0u);
    }
    return len;
}
P2B_HD UniformDigits uniform_digits(const uint32_t k[8]) {
    GlvSplit s = glv_decompose(k);
    UniformDigits u;
    uint32_t m1[6] = {s.k1[0], s.k1[1], s.k1[2], s.k1[3], s.k1[4], 0}, m2[6] = {s.k2[0], s.k2[1], s.k2[2], s.k2[3], s.k2[4], 0};
    for (int i = 0; i < 136; i++) { u.d1[i] = 0; u.d2[i] = 0; }
    int l1 = wnaf5_recode<6, 136>(u.d1, m1), l2 = wnaf5_recode<6, 136>(u.d2, m2);   // |k1|, |k2| < 2^132
    if (s.neg1) for (int i = 0; i < l1; i++) u.d1[i] = (int8_t)-u.d1[i];
    if (s.neg2) for (int i = 0; i < l2; i++) u.d2[i] = (int8_t)-u.d2[i];
    u.len = l1 > l2 ? l1 : l2;
    return u;
}
template <class F, class Tbl> P2B_HD Jac<F> mul_glv_uniform(const Aff<F> &p, const UniformDigits &u, const Tbl &tbl, F *zr, bool &bad) {
    F zg = build_odd_table<F>(p, tbl, zr, bad);
    Jac<F> acc = jac_infinity<F>();
#pragma unroll 1
    for (int i = u.len - 1; i >= 0; i--) {
        acc = jac_dbl(acc);
        const int a = u.d1[i], b = u.d2[i];                            // warp-uniform
        if (a) {
            Aff<F> q = tbl.get(((a < 0 ? -a : a) - 1) >> 1);
            q.y = cneg(q.y, a < 0);
            acc = jac_madd(acc, q);
        }
        if (b) {
            Aff<F> q = tbl.get(((b < 0 ? -b : b) - 1) >> 1);
            q.x = endo_x(q.x);
            q.y = cneg(q.y, b < 0);
            acc = jac_madd(acc, q);
        }
    }
    acc.z = mul(acc.z, zg);
    return acc;
}

template <class Tbl> P2B_HD Jac<Fq> g1_mul_glv(const Aff<Fq> &p, const uint32_t k[8], const Tbl &tbl, Fq *zr, bool &bad) {
    return mul_glv<Fq>(p, k, tbl, zr, bad);
}

// ------------------------------------------------------------------ G2 (and generic): 64 signed windows
// k canonical (< r < 2^254).  Valid for any curve point whose small multiples are distinct (`bad` otherwise).
template <class F, class Tbl> P2B_HD Jac<F> mul_window4(const Aff<F> &p, const uint32_t k[8], const Tbl &tbl, F *zr, bool &bad) {
    uint32_t kk[8];
    bool even = !(k[0] & 1);
    {   // kk = k | 1 (k even -> k + 1); subtract P at the end
#pragma unroll
        for (int i = 0; i < 8; i++) kk[i] = k[i];
        kk[0] |= 1u;
    }
    F zg = build_odd_table<F>(p, tbl, zr, bad);
    uint32_t m[8];
    shr1<8>(m, kk, 0u);                         // 63 nibbles (bits 1..252) + top = bit 253
    uint32_t top = kk[7] >> 29;                 // k >> 253  (0 or 1 for k < 2^254)
    // align: m holds 255 bits; nibble 63 would be bits 252..255 of m = (top, 0,0,0): pop order below handles it
    Aff<F> q = tbl.get(top & 7);
    Jac<F> acc = jac_from_aff(q);
    // drop the top nibble (bits 252..255 of m, i.e. k bits 253..256) -- consumed as `top`
    (void)pop_nibble<8>(m);
#pragma unroll 1
    for (int i = 0; i < 63; i++) {
#pragma unroll 1
        for (int d = 0; d < 4; d++) acc = jac_dbl(acc);
        uint32_t u = pop_nibble<8>(m);
        uint32_t idx; bool negd;
        digit_from_nibble(u, idx, negd);
        q = tbl.get(idx);
        q.y = cneg(q.y, negd);
        acc = jac_madd(acc, q);
    }
    {   // k was even: we computed [k+1]P, subtract P (lane-wise select keeps the warp converged)
        q = tbl.get(0);
        q.y = neg(q.y);
        Jac<F> fixed = jac_madd(acc, q);
        acc = select(even, fixed, acc);
    }
    acc.z = mul(acc.z, zg);
    return acc;
}

}  // namespace p2b
