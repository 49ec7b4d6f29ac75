// xyzz.cuh -- extended Jacobian (X, Y, ZZ, ZZZ) arithmetic for the MSM buckets: x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2.
// EFD "xyzz" formulas for a = 0 (madd-2008-s 8M+2S, add-2008-s 12M+2S, dbl-2008-s-1 6M+4S... counted with S = M here),
// completed with lane-wise selects for infinity / equal / opposite operands.  Host+device for unit tests.
#pragma once
#include "ec.cuh"

namespace p2b {

template <class F> struct Xyzz { F x, y, zz, zzz; };   // infinity <=> zz == 0
template <class F> P2B_HD Xyzz<F> xyzz_infinity() {
    Xyzz<F> r; r.x = FieldTraits<F>::zero(); r.y = FieldTraits<F>::zero(); r.zz = FieldTraits<F>::zero(); r.zzz = FieldTraits<F>::zero();
    return r;
}
template <class F> P2B_HD Xyzz<F> select(bool c, const Xyzz<F> &b, const Xyzz<F> &a) {
    Xyzz<F> r; r.x = select(c, b.x, a.x); r.y = select(c, b.y, a.y); r.zz = select(c, b.zz, a.zz); r.zzz = select(c, b.zzz, a.zzz);
    return r;
}
// mdbl-2008-s-1 (affine input)
template <class F> P2B_HD Xyzz<F> xyzz_dbl_aff(const Aff<F> &p) {
    F u = dbl(p.y), v = sqr(u), w = mul(u, v), s = mul(p.x, v);
    F xx = sqr(p.x), m = add(dbl(xx), xx);
    Xyzz<F> r;
    r.x = sub(sub(sqr(m), s), s);
    r.y = sub(mul(m, sub(s, r.x)), mul(w, p.y));
    r.zz = v; r.zzz = w;
    return r;
}
// dbl-2008-s-1
template <class F> P2B_HD Xyzz<F> xyzz_dbl(const Xyzz<F> &p) {
    F u = dbl(p.y), v = sqr(u), w = mul(u, v), s = mul(p.x, v);
    F xx = sqr(p.x), m = add(dbl(xx), xx);
    Xyzz<F> r;
    r.x = sub(sub(sqr(m), s), s);
    r.y = mul_sub2(m, sub(s, r.x), w, p.y);
    r.zz = mul(v, p.zz); r.zzz = mul(w, p.zzz);
    return r;                                    // infinity (zz = 0) stays infinity
}
// madd-2008-s, complete (p may be infinity, q == +-p handled); q is never infinity
template <class F> P2B_HD Xyzz<F> xyzz_madd(const Xyzz<F> &p, const Aff<F> &q) {
    F u2 = mul(q.x, p.zz), s2 = mul(q.y, p.zzz);
    F pp_ = sub(u2, p.x), rr = sub(s2, p.y);
    F pp = sqr(pp_), ppp = mul(pp_, pp), qq = mul(p.x, pp);
    Xyzz<F> r;
    r.x = sub(sub(sub(sqr(rr), ppp), qq), qq);
    r.y = sub(mul(rr, sub(qq, r.x)), mul(p.y, ppp));
    r.zz = mul(p.zz, pp); r.zzz = mul(p.zzz, ppp);
    bool p_inf = is_zero(p.zz);
    if (!p_inf & is_zero(pp_) & is_zero(rr)) r = xyzz_dbl_aff(q);
    Xyzz<F> qa; qa.x = q.x; qa.y = q.y; qa.zz = FieldTraits<F>::one(); qa.zzz = FieldTraits<F>::one();
    return select(p_inf, qa, r);
}
// the same with the two squarings on the dedicated squaring routine (tuning variant of the bucket accumulation)
template <class F> P2B_HD Xyzz<F> xyzz_madd_sq(const Xyzz<F> &p, const Aff<F> &q) {
    F u2 = mul(q.x, p.zz), s2 = mul(q.y, p.zzz);
    F pp_ = sub(u2, p.x), rr = sub(s2, p.y);
    F pp = sqr_ded(pp_), ppp = mul(pp_, pp), qq = mul(p.x, pp);
    Xyzz<F> r;
    r.x = sub(sub(sub(sqr_ded(rr), ppp), qq), qq);
    r.y = sub(mul(rr, sub(qq, r.x)), mul(p.y, ppp));
    r.zz = mul(p.zz, pp); r.zzz = mul(p.zzz, ppp);
    bool p_inf = is_zero(p.zz);
    if (!p_inf & is_zero(pp_) & is_zero(rr)) r = xyzz_dbl_aff(q);
    Xyzz<F> qa; qa.x = q.x; qa.y = q.y; qa.zz = FieldTraits<F>::one(); qa.zzz = FieldTraits<F>::one();
    return select(p_inf, qa, r);
}
// variant of xyzz_madd_sq with Y3 = R (Q - X3) + Y1 (-PPP) as ONE two-product Montgomery multiplication (G1 tuning variant)
P2B_HD Xyzz<Fq> xyzz_madd_fused(const Xyzz<Fq> &p, const Aff<Fq> &q) {
    Fq u2 = mul(q.x, p.zz), s2 = mul(q.y, p.zzz);
    Fq pp_ = sub(u2, p.x), rr = sub(s2, p.y);
    Fq pp = sqr_ded(pp_), ppp = mul(pp_, pp), qq = mul(p.x, pp);
    Xyzz<Fq> r;
    r.x = sub(sub(sub(sqr_ded(rr), ppp), qq), qq);
    r.y = mul2_add(rr, sub(qq, r.x), p.y, neg(ppp));
    r.zz = mul(p.zz, pp); r.zzz = mul(p.zzz, ppp);
    bool p_inf = is_zero(p.zz);
    if (!p_inf & is_zero(pp_) & is_zero(rr)) r = xyzz_dbl_aff(q);
    Xyzz<Fq> qa; qa.x = q.x; qa.y = q.y; qa.zz = FieldTraits<Fq>::one(); qa.zzz = FieldTraits<Fq>::one();
    return select(p_inf, qa, r);
}
// add-2008-s, complete
template <class F> P2B_HD Xyzz<F> xyzz_add(const Xyzz<F> &p, const Xyzz<F> &q) {
    F u1 = mul(p.x, q.zz), u2 = mul(q.x, p.zz), s1 = mul(p.y, q.zzz), s2 = mul(q.y, p.zzz);
    F pp_ = sub(u2, u1), rr = sub(s2, s1);
    F pp = sqr(pp_), ppp = mul(pp_, pp), qq = mul(u1, pp);
    Xyzz<F> r;
    r.x = sub(sub(sub(sqr(rr), ppp), qq), qq);
    r.y = mul_sub2(rr, sub(qq, r.x), s1, ppp);
    r.zz = mul(mul(p.zz, q.zz), pp); r.zzz = mul(mul(p.zzz, q.zzz), ppp);
    bool p_inf = is_zero(p.zz), q_inf = is_zero(q.zz);
    if (!p_inf & !q_inf & is_zero(pp_) & is_zero(rr)) r = xyzz_dbl(p);
    r = select(q_inf, p, r);
    return select(p_inf, q, r);
}

}  // namespace p2b
