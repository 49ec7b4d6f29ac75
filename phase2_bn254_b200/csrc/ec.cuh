// ec.cuh -- Fq2 and short-Weierstrass (a = 0) Jacobian arithmetic for BN254 G1 / G2.
//
// GPU counterpart of the `curve_impl!` macro in pairing/src/bn256/ec.rs:251-629 and of
// pairing/src/bn256/fq2.rs:131-199.  The formulas are the same EFD ones the reference uses
// (dbl-2009-l ec.rs:301-358, add-2007-bl ec.rs:360-454, madd-2007-bl ec.rs:456-536); the
// exceptional cases (P = inf, P = Q, P = -Q) are resolved with lane-wise selects so a warp never
// diverges on honest inputs.  Only affine results are ever encoded, so any algorithm with the
// same group-law result is bit-identical on the wire.
#pragma once
#include "fp.cuh"

namespace p2b {

// ------------------------------------------------------------------ Fq2 = Fq[u]/(u^2+1)
struct Fq2 {
    Fq c0, c1;
};
P2B_HD Fq2 fq2_zero() { Fq2 r; r.c0 = fp_zero<FqP>(); r.c1 = fp_zero<FqP>(); return r; }
P2B_HD Fq2 fq2_one() { Fq2 r; r.c0 = fp_one<FqP>(); r.c1 = fp_zero<FqP>(); return r; }
P2B_HD bool is_zero(const Fq2 &a) { return is_zero(a.c0) & is_zero(a.c1); }
P2B_HD bool eq(const Fq2 &a, const Fq2 &b) { return eq(a.c0, b.c0) & eq(a.c1, b.c1); }
P2B_HD Fq2 add(const Fq2 &a, const Fq2 &b) { Fq2 r; r.c0 = add(a.c0, b.c0); r.c1 = add(a.c1, b.c1); return r; }
P2B_HD Fq2 sub(const Fq2 &a, const Fq2 &b) { Fq2 r; r.c0 = sub(a.c0, b.c0); r.c1 = sub(a.c1, b.c1); return r; }
P2B_HD Fq2 dbl(const Fq2 &a) { Fq2 r; r.c0 = dbl(a.c0); r.c1 = dbl(a.c1); return r; }
P2B_HD Fq2 neg(const Fq2 &a) { Fq2 r; r.c0 = neg(a.c0); r.c1 = neg(a.c1); return r; }
P2B_HD Fq2 select(bool c, const Fq2 &b, const Fq2 &a) {
    Fq2 r; r.c0 = select(c, b.c0, a.c0); r.c1 = select(c, b.c1, a.c1); return r;
}
P2B_HD Fq2 cneg(const Fq2 &a, bool c) { return select(c, neg(a), a); }
// (a0 + a1 u)(b0 + b1 u) = (a0 b0 - a1 b1) + ((a0 + a1)(b0 + b1) - a0 b0 - a1 b1) u   (fq2.rs:167-180)
P2B_HD Fq2 mul(const Fq2 &a, const Fq2 &b) {
#if defined(__CUDA_ARCH__) && defined(P2B_FQ2_SCHOOLBOOK)
    // schoolbook with two fused two-product multiplications (400 wide multiplies, no 512-bit temporaries): tuning variant
    Fq2 r;
    mont_mul2<FqP>(r.c0.l, a.c0.l, b.c0.l, a.c1.l, neg(b.c1).l);
    mont_mul2<FqP>(r.c1.l, a.c0.l, b.c1.l, a.c1.l, b.c0.l);
    return r;
#elif defined(__CUDA_ARCH__) && defined(P2B_FQ2_LAZY)
    // lazy reduction: three 512-bit products, two Montgomery reductions (336 instead of 408 wide multiply-adds)
    //   c0 = (a0 b0 - a1 b1 + q^2) / R      in (0, 2 q^2) < q 2^256
    //   c1 = ((a0 + a1)(b0 + b1) - a0 b0 - a1 b1) / R = (a0 b1 + a1 b0) / R   in [0, 2 q^2)
    uint32_t t0[16], t1[16], d[16], p2[16];
    wide_mul(t0, a.c0.l, b.c0.l);
    wide_mul(t1, a.c1.l, b.c1.l);
#pragma unroll
    for (int i = 0; i < 16; i++) p2[i] = d_FQ_P2[i];
    sub16(d, t0, t1);                    // may wrap below zero: the + q^2 brings it back (mod 2^512 arithmetic)
    add16(d, d, p2);
    Fq2 r;
    mont_red<FqP>(r.c0.l, d);
    add16(t0, t0, t1);
    uint32_t sa[8], sb[8];
    add8(sa, a.c0.l, a.c1.l);            // < 2q < 2^256: unreduced sums are fine as multiplier inputs
    add8(sb, b.c0.l, b.c1.l);
    wide_mul(t1, sa, sb);
    sub16(t1, t1, t0);
    mont_red<FqP>(r.c1.l, t1);
    return r;
#else
    Fq aa = mul(a.c0, b.c0);
    Fq bb = mul(a.c1, b.c1);
    Fq s = mul(add(a.c0, a.c1), add(b.c0, b.c1));
    Fq2 r;
    r.c1 = sub(sub(s, aa), bb);
    r.c0 = sub(aa, bb);
    return r;
#endif
}
// (a0 + a1 u)^2 = (a0 + a1)(a0 - a1) + 2 a0 a1 u   (fq2.rs:131-145)
P2B_HD Fq2 sqr(const Fq2 &a) {
    Fq ab = mul(a.c0, a.c1);
    Fq2 r;
    r.c0 = mul(add(a.c0, a.c1), sub(a.c0, a.c1));
    r.c1 = dbl(ab);
    return r;
}
P2B_HD Fq2 sqr_ded(const Fq2 &a) { return sqr(a); }
P2B_HD Fq2 mul_fq(const Fq2 &a, const Fq &b) { Fq2 r; r.c0 = mul(a.c0, b); r.c1 = mul(a.c1, b); return r; }
P2B_HD Fq2 inv(const Fq2 &a) {  // fq2.rs:182-199
    Fq t = inv(add(sqr(a.c0), sqr(a.c1)));
    Fq2 r;
    r.c0 = mul(a.c0, t);
    r.c1 = neg(mul(a.c1, t));
    return r;
}
P2B_HD Fq2 conj(const Fq2 &a) { Fq2 r; r.c0 = a.c0; r.c1 = neg(a.c1); return r; }
// y > -y in the reference's Fq2 order: c1 first, then c0 (fq2.rs:21-30)
P2B_HD bool is_lexicographically_largest(const Fq2 &y) {
    // compare (c1, c0) with (-c1, -c0): if c1 != 0 the c1 comparison decides, else c0 decides
    return is_zero(y.c1) ? is_lexicographically_largest(y.c0) : is_lexicographically_largest(y.c1);
}

template <class F> struct FieldTraits;
template <> struct FieldTraits<Fq> {
    static P2B_HD Fq zero() { return fp_zero<FqP>(); }
    static P2B_HD Fq one() { return fp_one<FqP>(); }
    static constexpr int WORDS = 8;
};
template <> struct FieldTraits<Fq2> {
    static P2B_HD Fq2 zero() { return fq2_zero(); }
    static P2B_HD Fq2 one() { return fq2_one(); }
    static constexpr int WORDS = 16;
};
// flat word access (word w of the element; Fq2 = c0 words then c1 words)
P2B_HD uint32_t get_word(const Fq &a, int w) { return a.l[w]; }
P2B_HD void set_word(Fq &a, int w, uint32_t v) { a.l[w] = v; }
P2B_HD uint32_t get_word(const Fq2 &a, int w) { return w < 8 ? a.c0.l[w] : a.c1.l[w - 8]; }
P2B_HD void set_word(Fq2 &a, int w, uint32_t v) { if (w < 8) a.c0.l[w] = v; else a.c1.l[w - 8] = v; }

// ------------------------------------------------------------------ points
template <class F> struct Aff { F x, y; };          // never the point at infinity
template <class F> struct Jac { F x, y, z; };       // infinity <=> z == 0

template <class F> P2B_HD Jac<F> jac_infinity() {
    Jac<F> r; r.x = FieldTraits<F>::zero(); r.y = FieldTraits<F>::one(); r.z = FieldTraits<F>::zero(); return r;
}
template <class F> P2B_HD Jac<F> jac_from_aff(const Aff<F> &a) {
    Jac<F> r; r.x = a.x; r.y = a.y; r.z = FieldTraits<F>::one(); return r;
}
template <class F> P2B_HD Jac<F> select(bool c, const Jac<F> &b, const Jac<F> &a) {
    Jac<F> r; r.x = select(c, b.x, a.x); r.y = select(c, b.y, a.y); r.z = select(c, b.z, a.z); return r;
}

// a b - c d: for Fq one fused two-product Montgomery multiplication (fp.cuh mont_mul2: 200 instead of 272 wide multiplies)
template <class F> P2B_HD F mul_sub2(const F &a, const F &b, const F &c, const F &d) { return sub(mul(a, b), mul(c, d)); }
template <> P2B_HD Fq mul_sub2<Fq>(const Fq &a, const Fq &b, const Fq &c, const Fq &d) { return mul2_add(a, b, c, neg(d)); }

// dbl-2009-l: 2M + 5S.  z == 0 stays 0.
template <class F> P2B_HD Jac<F> jac_dbl(const Jac<F> &p) {
    F a = sqr(p.x);
    F b = sqr(p.y);
    F c = sqr(b);
    F d = dbl(sub(sub(sqr(add(p.x, b)), a), c));
    F e = add(dbl(a), a);
    F f = sqr(e);
    Jac<F> r;
    r.z = dbl(mul(p.y, p.z));
    r.x = sub(sub(f, d), d);
    F c8 = dbl(dbl(dbl(c)));
    r.y = sub(mul(e, sub(d, r.x)), c8);
    return r;
}
// The same doubling for Fq with the fused two-product multiplication: D = 4 X B needs no C, and Y3 = E (D - X3) - (8 B) B is ONE
// mont_mul2 -- 5 multiplications + 1 fused (880 wide multiplies) instead of 7 multiplications (952).  (Not for Fq2, where the
// squaring (X + B)^2 is cheaper than the product X B.)
P2B_HD Jac<Fq> jac_dbl(const Jac<Fq> &p) {
    Fq a = sqr(p.x);
    Fq b = sqr(p.y);
    Fq d = dbl(dbl(mul(p.x, b)));
    Fq e = add(dbl(a), a);
    Fq f = sqr(e);
    Jac<Fq> r;
    r.z = dbl(mul(p.y, p.z));
    r.x = sub(sub(f, d), d);
    r.y = mul_sub2(e, sub(d, r.x), dbl(dbl(dbl(b))), b);
    return r;
}
// doubling of an affine point (Z1 = 1): mdbl-2007-bl shape, 1M + 5S
template <class F> P2B_HD Jac<F> aff_dbl(const Aff<F> &p) {
    F a = sqr(p.x);
    F b = sqr(p.y);
    F c = sqr(b);
    F d = dbl(sub(sub(sqr(add(p.x, b)), a), c));
    F e = add(dbl(a), a);
    F f = sqr(e);
    Jac<F> r;
    r.z = dbl(p.y);
    r.x = sub(sub(f, d), d);
    F c8 = dbl(dbl(dbl(c)));
    r.y = sub(mul(e, sub(d, r.x)), c8);
    return r;
}

// madd-2007-bl (7M + 4S), complete: p may be infinity, q may equal +-p.  q is never infinity.
template <class F> P2B_HD Jac<F> jac_madd(const Jac<F> &p, const Aff<F> &q) {
    F z1z1 = sqr(p.z);
    F u2 = mul(q.x, z1z1);
    F s2 = mul(mul(q.y, p.z), z1z1);
    F h = sub(u2, p.x);
    F hh = sqr(h);
    F i = dbl(dbl(hh));
    F j = mul(h, i);
    F rr = dbl(sub(s2, p.y));
    F v = mul(p.x, i);
    Jac<F> r;
    r.x = sub(sub(sub(sqr(rr), j), v), v);
    r.y = mul_sub2(rr, sub(v, r.x), p.y, dbl(j));
    r.z = sub(sub(sqr(add(p.z, h)), z1z1), hh);
    bool p_inf = is_zero(p.z);
    bool same = is_zero(h) & is_zero(rr) & !p_inf;
    if (same) r = aff_dbl(q);                       // P == Q: rare, the only divergent path
    r = select(p_inf, jac_from_aff(q), r);          // inf + Q = Q
    return r;                                       // P == -Q leaves z = 2*Z1*H = 0 = infinity
}

// add-2007-bl (11M + 5S), complete.
template <class F> P2B_HD Jac<F> jac_add(const Jac<F> &p, const Jac<F> &q) {
    F z1z1 = sqr(p.z);
    F z2z2 = sqr(q.z);
    F u1 = mul(p.x, z2z2);
    F u2 = mul(q.x, z1z1);
    F s1 = mul(mul(p.y, q.z), z2z2);
    F s2 = mul(mul(q.y, p.z), z1z1);
    F h = sub(u2, u1);
    F i = sqr(dbl(h));
    F j = mul(h, i);
    F rr = dbl(sub(s2, s1));
    F v = mul(u1, i);
    Jac<F> r;
    r.x = sub(sub(sub(sqr(rr), j), v), v);
    r.y = mul_sub2(rr, sub(v, r.x), s1, dbl(j));
    r.z = mul(sub(sub(sqr(add(p.z, q.z)), z1z1), z2z2), h);
    bool p_inf = is_zero(p.z), q_inf = is_zero(q.z);
    bool same = is_zero(h) & is_zero(rr) & !p_inf & !q_inf;
    if (same) r = jac_dbl(p);
    r = select(q_inf, p, r);
    r = select(p_inf, q, r);
    return r;
}
template <class F> P2B_HD Jac<F> jac_neg(const Jac<F> &p) { Jac<F> r = p; r.y = neg(p.y); return r; }

// curve constants b (Montgomery form): fq.rs:11-31
P2B_DEF_CONST(G1_B, {0x50ad28d7u, 0x7a17caa9u, 0xe15521b9u, 0x1f6ac17au, 0x696bd284u, 0x334bea4eu, 0xce179d8eu, 0x2a1f6744u})
P2B_DEF_CONST(G2_B0, {0x77b802a8u, 0x3bf938e3u, 0x3633535du, 0x020b1b27u, 0x49755260u, 0x26b7edf0u, 0x4384a86du, 0x2514c632u})
P2B_DEF_CONST(G2_B1, {0xd1dcff67u, 0x38e7ecccu, 0x93ce0d3eu, 0x65f0b37du, 0x22ac00aau, 0xd749d0ddu, 0x4a688d4du, 0x0141b9ceu})
P2B_HD Fq curve_b(const Fq *) { Fq b; for (int i = 0; i < 8; i++) b.l[i] = P2B_C(G1_B, i); return b; }
P2B_HD Fq2 curve_b(const Fq2 *) {
    Fq2 b;
    for (int i = 0; i < 8; i++) { b.c0.l[i] = P2B_C(G2_B0, i); b.c1.l[i] = P2B_C(G2_B1, i); }
    return b;
}
// y^2 == x^3 + b   (ec.rs:133-148)
template <class F> P2B_HD bool on_curve(const Aff<F> &p) {
    F lhs = sqr(p.y);
    F rhs = add(mul(sqr(p.x), p.x), curve_b((const F *)nullptr));
    return eq(lhs, rhs);
}

}  // namespace p2b
