// pairing.cuh -- host-side BN254 optimal ate pairing, used only for the verifier's `same_ratio` checks.
//
// The reference keeps its <= 20 pairings per verification on the CPU (powersoftau/src/utils.rs:151-159,
// phase2/src/utils.rs:48-57: `g1.0.pairing_with(&g2.1) == g1.1.pairing_with(&g2.0)`); the GPU does the linear
// combinations that feed them (merge_pairs / power_pairs = Pippenger MSMs).  This file is that host part: the tower
// Fq2 -> Fq6 -> Fq12 of pairing/src/bn256/{fq2,fq6,fq12}.rs, a Miller loop over 6u+2 (pairing/src/bn256/mod.rs:29-130)
// and the final exponentiation (mod.rs:132-240).  Only equality of pairing PRODUCTS is exposed (is the product of
// e(P_i, Q_i) the identity?), which does not depend on how the Miller function is normalised; the line functions here
// are the plain affine ones and the hard part of the final exponentiation is a square-and-multiply over (q^6+1)/r.
//
// Host code only (the field layer of fp.cuh / ec.cuh compiles for the host); nothing here runs on the GPU.
#pragma once
#include "codec.cuh"

namespace p2b {
namespace pairing {

// ------------------------------------------------------------------ Fq6 = Fq2[v]/(v^3 - xi), xi = 9 + u   (fq6.rs)
struct Fq6 { Fq2 c0, c1, c2; };
struct Fq12 { Fq6 c0, c1; };       // Fq12 = Fq6[w]/(w^2 - v)   (fq12.rs)

inline Fq2 mul_by_xi(const Fq2 &a) {     // (9 + u)(a0 + a1 u) = (9 a0 - a1) + (9 a1 + a0) u   (fq2.rs:40-58)
    Fq2 t = dbl(dbl(dbl(a)));            // 8a
    Fq2 r;
    r.c0 = sub(add(t.c0, a.c0), a.c1);
    r.c1 = add(add(t.c1, a.c1), a.c0);
    return r;
}
inline Fq6 fq6_zero() { Fq6 r; r.c0 = fq2_zero(); r.c1 = fq2_zero(); r.c2 = fq2_zero(); return r; }
inline Fq6 fq6_one() { Fq6 r = fq6_zero(); r.c0 = fq2_one(); return r; }
inline Fq6 add(const Fq6 &a, const Fq6 &b) { Fq6 r; r.c0 = add(a.c0, b.c0); r.c1 = add(a.c1, b.c1); r.c2 = add(a.c2, b.c2); return r; }
inline Fq6 sub(const Fq6 &a, const Fq6 &b) { Fq6 r; r.c0 = sub(a.c0, b.c0); r.c1 = sub(a.c1, b.c1); r.c2 = sub(a.c2, b.c2); return r; }
inline Fq6 neg(const Fq6 &a) { Fq6 r; r.c0 = neg(a.c0); r.c1 = neg(a.c1); r.c2 = neg(a.c2); return r; }
inline bool eq(const Fq6 &a, const Fq6 &b) { return eq(a.c0, b.c0) && eq(a.c1, b.c1) && eq(a.c2, b.c2); }
inline Fq6 mul_by_v(const Fq6 &a) { Fq6 r; r.c0 = mul_by_xi(a.c2); r.c1 = a.c0; r.c2 = a.c1; return r; }
// schoolbook product reduced with v^3 = xi (Karatsuba as in fq6.rs:233-283 is not worth it on the host)
inline Fq6 mul(const Fq6 &a, const Fq6 &b) {
    Fq2 a0b0 = mul(a.c0, b.c0), a1b1 = mul(a.c1, b.c1), a2b2 = mul(a.c2, b.c2);
    Fq2 t12 = sub(sub(mul(add(a.c1, a.c2), add(b.c1, b.c2)), a1b1), a2b2);   // a1 b2 + a2 b1
    Fq2 t01 = sub(sub(mul(add(a.c0, a.c1), add(b.c0, b.c1)), a0b0), a1b1);   // a0 b1 + a1 b0
    Fq2 t02 = sub(sub(mul(add(a.c0, a.c2), add(b.c0, b.c2)), a0b0), a2b2);   // a0 b2 + a2 b0
    Fq6 r;
    r.c0 = add(a0b0, mul_by_xi(t12));
    r.c1 = add(t01, mul_by_xi(a2b2));
    r.c2 = add(t02, a1b1);
    return r;
}
inline Fq6 inv(const Fq6 &a) {           // fq6.rs:285-336
    Fq2 t0 = sub(sqr(a.c0), mul_by_xi(mul(a.c1, a.c2)));
    Fq2 t1 = sub(mul_by_xi(sqr(a.c2)), mul(a.c0, a.c1));
    Fq2 t2 = sub(sqr(a.c1), mul(a.c0, a.c2));
    Fq2 d = add(mul(a.c0, t0), mul_by_xi(add(mul(a.c2, t1), mul(a.c1, t2))));
    Fq2 di = p2b::inv(d);
    Fq6 r;
    r.c0 = mul(t0, di);
    r.c1 = mul(t1, di);
    r.c2 = mul(t2, di);
    return r;
}

inline Fq12 fq12_one() { Fq12 r; r.c0 = fq6_one(); r.c1 = fq6_zero(); return r; }
inline bool eq(const Fq12 &a, const Fq12 &b) { return eq(a.c0, b.c0) && eq(a.c1, b.c1); }
inline Fq12 mul(const Fq12 &a, const Fq12 &b) {    // fq12.rs:118-131
    Fq6 aa = mul(a.c0, b.c0), bb = mul(a.c1, b.c1);
    Fq12 r;
    r.c1 = sub(sub(mul(add(a.c0, a.c1), add(b.c0, b.c1)), aa), bb);
    r.c0 = add(aa, mul_by_v(bb));
    return r;
}
inline Fq12 sqr(const Fq12 &a) {                   // fq12.rs:133-149: (c0 + c1 w)^2
    Fq6 ab = mul(a.c0, a.c1);
    Fq6 t = mul(add(a.c0, a.c1), add(a.c0, mul_by_v(a.c1)));
    Fq12 r;
    r.c0 = sub(sub(t, ab), mul_by_v(ab));
    r.c1 = add(ab, ab);
    return r;
}
inline Fq12 conj(const Fq12 &a) { Fq12 r; r.c0 = a.c0; r.c1 = neg(a.c1); return r; }     // = a^(q^6)
inline Fq12 inv(const Fq12 &a) {                   // fq12.rs:151-166
    Fq6 d = inv(sub(mul(a.c0, a.c0), mul_by_v(mul(a.c1, a.c1))));
    Fq12 r;
    r.c0 = mul(a.c0, d);
    r.c1 = neg(mul(a.c1, d));
    return r;
}

// ------------------------------------------------------------------ constants
// (q - 1) / 6 and (q^6 + 1) / r, little-endian 32-bit words
static const uint32_t Q_MINUS_1_OVER_6[8] = {0x2414d4e1u, 0x34b01759u, 0xe6bda1c2u, 0xee9591c2u, 0xc0403964u, 0xf40d60f3u, 0xd032f006u, 0x0810b7bdu};
static const uint32_t HARD_EXP[40] = {
    0x36e3f812u, 0x5250a540u, 0x96789051u, 0xa5635f15u, 0x4d5bd1d4u, 0xd1138bf5u, 0xbe36c7a2u, 0xa8ce2533u,
    0x84e09bf6u, 0x94f69f6bu, 0x50ef3644u, 0x42ad1f5eu, 0x48c3454cu, 0x0fcc420eu, 0xecc9952cu, 0x758e4408u,
    0x87c6042cu, 0xc901bf18u, 0xb14bb3b5u, 0xa733cd65u, 0xcf51b0d8u, 0xdf6d76bdu, 0x82eb59e1u, 0xca64c0fdu,
    0xe39276a1u, 0x1d2e5726u, 0xa391cae9u, 0xc2d1ea74u, 0xc82d647eu, 0x07409206u, 0xa5afdd17u, 0x051c6d1au,
    0x19667af5u, 0xb37f6019u, 0x5084015bu, 0x150e578cu, 0xc23998e4u, 0xfbdea556u, 0xc52f5b83u, 0x000fd14cu};
// 6u + 2, u = 4965661367192848881 (mod.rs:29); 65 bits
static const uint64_t SIX_U_PLUS_2_LO = 0x9d797039be763ba8ull;   // bit 64 is set on top of this

// gamma1 = xi^((q-1)/6) = FROBENIUS_COEFF_FQ12_C1[1] (fq.rs:280-432); gamma1^2 = FROBENIUS_COEFF_FQ6_C1[1] (fq.rs:121-199);
// gamma1^3 = XI_TO_Q_MINUS_1_OVER_2 (fq.rs:106-119).  Computed once; tests/test_pairing.py compares them with those tables.
struct FrobConsts {
    Fq2 g1, g2, g3;     // gamma1, gamma1^2, gamma1^3
    Fq n2, n3;          // N^2, N^3 with N = gamma1 * conj(gamma1) = xi^((q^2-1)/6)  (in Fq)
};
inline const FrobConsts &frob_consts() {
    static const FrobConsts k = [] {
        Fq2 xi;
        Fq one = fp_one<FqP>();
        Fq nine = one;
        for (int i = 0; i < 8; i++) nine = add(nine, one);
        xi.c0 = nine;
        xi.c1 = one;
        Fq2 acc = fq2_one();
        for (int i = 255; i >= 0; i--) {
            acc = sqr(acc);
            if ((Q_MINUS_1_OVER_6[i >> 5] >> (i & 31)) & 1u) acc = mul(acc, xi);
        }
        FrobConsts c;
        c.g1 = acc;
        c.g2 = sqr(acc);
        c.g3 = mul(c.g2, acc);
        Fq2 nn = mul(acc, conj(acc));            // norm: c1 == 0
        c.n2 = sqr(nn.c0);
        c.n3 = mul(c.n2, nn.c0);
        return c;
    }();
    return k;
}

// ------------------------------------------------------------------ Miller loop
// Q on the twist E': y^2 = x^3 + 3/xi over Fq2; untwist (x', y') -> (x' w^2, y' w^3).  The line through psi(T) with twist
// slope lam evaluated at P = (xp, yp):   yp  -  (lam xp) w  +  (lam xT - yT) v w      (vertical lines drop out).
inline Fq12 line_eval(const Fq2 &lam, const Fq2 &xt, const Fq2 &yt, const Aff<Fq> &p) {
    Fq12 l;
    l.c0 = fq6_zero();
    l.c1 = fq6_zero();
    l.c0.c0.c0 = p.y;
    l.c1.c0 = neg(mul_fq(lam, p.x));
    l.c1.c1 = sub(mul(lam, xt), yt);
    return l;
}
inline void dbl_step(Fq12 &f, Aff<Fq2> &t, const Aff<Fq> &p) {
    Fq2 x2 = sqr(t.x);
    Fq2 lam = mul(add(dbl(x2), x2), p2b::inv(dbl(t.y)));
    f = mul(f, line_eval(lam, t.x, t.y, p));
    Fq2 x3 = sub(sub(sqr(lam), t.x), t.x);
    Fq2 y3 = sub(mul(lam, sub(t.x, x3)), t.y);
    t.x = x3;
    t.y = y3;
}
inline void add_step(Fq12 &f, Aff<Fq2> &t, const Aff<Fq2> &q, const Aff<Fq> &p, bool update_t = true) {
    Fq2 lam = mul(sub(t.y, q.y), p2b::inv(sub(t.x, q.x)));
    f = mul(f, line_eval(lam, q.x, q.y, p));
    if (!update_t) return;
    Fq2 x3 = sub(sub(sqr(lam), t.x), q.x);
    Fq2 y3 = sub(mul(lam, sub(t.x, x3)), t.y);
    t.x = x3;
    t.y = y3;
}
// f_{6u+2,Q}(P) * l_{[6u+2]Q, pi(Q)}(P) * l_{[6u+2]Q + pi(Q), -pi^2(Q)}(P)    (mod.rs:57-130 computes the same function)
inline Fq12 miller_loop(const Aff<Fq> &p, const Aff<Fq2> &q) {
    const FrobConsts &k = frob_consts();
    Fq12 f = fq12_one();
    Aff<Fq2> t = q;
    for (int i = 63; i >= 0; i--) {               // bit 64 is the leading one
        f = sqr(f);
        dbl_step(f, t, p);
        if ((SIX_U_PLUS_2_LO >> i) & 1ull) add_step(f, t, q, p);
    }
    Aff<Fq2> q1, q2;
    q1.x = mul(conj(q.x), k.g2);                  // pi(Q)
    q1.y = mul(conj(q.y), k.g3);
    q2.x = mul_fq(q.x, k.n2);                     // -pi^2(Q)
    q2.y = neg(mul_fq(q.y, k.n3));
    add_step(f, t, q1, p);
    add_step(f, t, q2, p, false);
    return f;
}
// f^((q^12 - 1) / r) = (f^(q^6 - 1))^((q^6 + 1) / r)
inline Fq12 final_exponentiation(const Fq12 &f) {
    Fq12 g = mul(conj(f), inv(f));
    Fq12 acc = fq12_one();
    bool started = false;
    for (int i = 40 * 32 - 1; i >= 0; i--) {
        if (started) acc = sqr(acc);
        if ((HARD_EXP[i >> 5] >> (i & 31)) & 1u) {
            acc = started ? mul(acc, g) : g;
            started = true;
        }
    }
    return acc;
}

// prod_i e(P_i, Q_i) == 1 ?   Pairs with a point at infinity contribute the identity (pairing/src/lib.rs: pairing with zero is one).
inline bool pairing_product_is_one(const Aff<Fq> *ps, const bool *p_inf, const Aff<Fq2> *qs, const bool *q_inf, size_t n) {
    Fq12 f = fq12_one();
    for (size_t i = 0; i < n; i++) {
        if (p_inf[i] || q_inf[i]) continue;
        f = mul(f, miller_loop(ps[i], qs[i]));
    }
    return eq(final_exponentiation(f), fq12_one());
}

// ------------------------------------------------------------------ wire helpers (host)
inline void bytes_to_words(uint32_t *w, const uint8_t *b, int nwords) {      // big-endian words as the device sees them after a raw load
    for (int i = 0; i < nwords; i++) w[i] = (uint32_t)b[4 * i] | ((uint32_t)b[4 * i + 1] << 8) | ((uint32_t)b[4 * i + 2] << 16) | ((uint32_t)b[4 * i + 3] << 24);
}
inline void words_to_bytes(uint8_t *b, const uint32_t *w, int nwords) {
    for (int i = 0; i < nwords; i++) { b[4 * i] = (uint8_t)w[i]; b[4 * i + 1] = (uint8_t)(w[i] >> 8); b[4 * i + 2] = (uint8_t)(w[i] >> 16); b[4 * i + 3] = (uint8_t)(w[i] >> 24); }
}
template <class F> inline int decode_host(Aff<F> &a, bool &inf, const uint8_t *bytes, bool check) {
    uint32_t w[Wire<F>::WORDS_UNCOMPRESSED];
    bytes_to_words(w, bytes, Wire<F>::WORDS_UNCOMPRESSED);
    return point_decode<F>(a, inf, w, ENC_UNCOMPRESSED, check);
}
template <class F> inline void encode_host(uint8_t *bytes, const Aff<F> &a, bool inf) {
    uint32_t w[Wire<F>::WORDS_UNCOMPRESSED];
    point_encode<F>(w, a, inf, ENC_UNCOMPRESSED);
    words_to_bytes(bytes, w, Wire<F>::WORDS_UNCOMPRESSED);
}

// ------------------------------------------------------------------ hash_to_g2 (powersoftau/src/utils.rs:31-45, phase2/src/utils.rs:111-122)
// ChaChaRng of rand 0.4.6 (the crate is not vendored under /root/reference; restated from its published source):
// ChaCha20 block function, key = the 8 seed words, 128-bit block counter starting at 0 in words 12..15, no nonce;
// next_u32 walks the 16 output words of a block, next_u64 = (next_u32 << 32) | next_u32.
struct ChaChaRng {
    uint32_t state[16], buf[16];
    int index;
    ChaChaRng() : index(16) { for (int i = 0; i < 16; i++) state[i] = buf[i] = 0; }
    explicit ChaChaRng(const uint32_t seed[8]) {
        state[0] = 0x61707865u; state[1] = 0x3320646eu; state[2] = 0x79622d32u; state[3] = 0x6b206574u;
        for (int i = 0; i < 8; i++) state[4 + i] = seed[i];
        for (int i = 12; i < 16; i++) state[i] = 0;
        index = 16;
    }
    static uint32_t rotl(uint32_t v, int c) { return (v << c) | (v >> (32 - c)); }
    static void quarter(uint32_t *x, int a, int b, int c, int d) {
        x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 16);
        x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 12);
        x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 8);
        x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 7);
    }
    void update() {
        for (int i = 0; i < 16; i++) buf[i] = state[i];
        for (int r = 0; r < 10; r++) {
            quarter(buf, 0, 4, 8, 12); quarter(buf, 1, 5, 9, 13); quarter(buf, 2, 6, 10, 14); quarter(buf, 3, 7, 11, 15);
            quarter(buf, 0, 5, 10, 15); quarter(buf, 1, 6, 11, 12); quarter(buf, 2, 7, 8, 13); quarter(buf, 3, 4, 9, 14);
        }
        for (int i = 0; i < 16; i++) buf[i] += state[i];
        index = 0;
        for (int i = 12; i < 16; i++) if (++state[i] != 0) break;
    }
    uint32_t next_u32() { if (index == 16) update(); return buf[index++]; }
    uint64_t next_u64() { uint64_t hi = next_u32(); return (hi << 32) | next_u32(); }
};
// `Fq::rand` of ff_derive: four u64 limbs (least significant first), the top 2 bits shaved, rejected unless < q; the limbs
// are taken as the Montgomery representation as they are.
inline Fq fq_rand(ChaChaRng &rng) {
    for (;;) {
        Fq a;
        for (int i = 0; i < 4; i++) { uint64_t v = rng.next_u64(); a.l[2 * i] = (uint32_t)v; a.l[2 * i + 1] = (uint32_t)(v >> 32); }
        a.l[7] &= 0x3fffffffu;
        if (is_canonical(a)) return a;
    }
}
// `Fr::rand`: same recipe over r.  Returns the canonical value (the limbs are the Montgomery representation).
inline Fr fr_rand_mont(ChaChaRng &rng) {
    for (;;) {
        Fr a;
        for (int i = 0; i < 4; i++) { uint64_t v = rng.next_u64(); a.l[2 * i] = (uint32_t)v; a.l[2 * i + 1] = (uint32_t)(v >> 32); }
        a.l[7] &= 0x3fffffffu;
        if (is_canonical(a)) return a;
    }
}
template <class F> inline Jac<F> mul_bits_host(const Aff<F> &p, const uint32_t *k, int nbits) {   // CurveAffine::mul_bits: MSB-first double-and-add
    Jac<F> acc = jac_infinity<F>();
    for (int i = nbits - 1; i >= 0; i--) {
        acc = jac_dbl(acc);
        if ((k[i >> 5] >> (i & 31)) & 1u) acc = jac_madd(acc, p);
    }
    return acc;
}
template <class F> inline bool jac_to_aff_host(Aff<F> &a, const Jac<F> &p) {     // false: infinity
    if (is_zero(p.z)) return false;
    F zi = inv(p.z), zi2 = sqr(zi);
    a.x = mul(p.x, zi2);
    a.y = mul(p.y, mul(zi2, zi));
    return true;
}
// `G1::rand` (ec.rs:711-726): x = Fq::rand, greatest = bool, y chosen as above; the cofactor of G1 is 1.
inline void g1_rand(Aff<Fq> &out, ChaChaRng &rng) {
    for (;;) {
        Fq x = fq_rand(rng);
        const bool greatest = (rng.next_u32() & 1u) != 0;
        Fq y;
        if (!fq_sqrt(y, add(mul(sqr(x), x), curve_b((const Fq *)nullptr)))) continue;
        Fq ny = neg(y);
        const bool y_is_smaller = is_lexicographically_largest(ny);
        out.x = x;
        out.y = (y_is_smaller != greatest) ? y : ny;
        return;
    }
}
// G2 cofactor 2q - r (ec.rs:1347-1357)
static const uint32_t G2_COFACTOR[8] = {0xc0f9fa8du, 0x345f2299u, 0x572a2489u, 0x06ceecdau, 0x8181585eu, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
// `G2::rand` (ec.rs:1091-1106): x = Fq2 { c0: rand, c1: rand }, greatest = next_u32 as u8 & 1, y = sqrt(x^3 + b) chosen
// by ((y < -y) ^ greatest), then scale_by_cofactor.  Returns false if the result is infinity (cannot happen for x on E').
inline bool g2_rand(Aff<Fq2> &out, ChaChaRng &rng) {
    for (;;) {
        Fq2 x;
        x.c0 = fq_rand(rng);
        x.c1 = fq_rand(rng);
        const bool greatest = (rng.next_u32() & 1u) != 0;
        Fq2 y;
        if (!fq2_sqrt(y, add(mul(sqr(x), x), curve_b((const Fq2 *)nullptr)))) continue;
        Fq2 ny = neg(y);
        const bool y_is_smaller = is_lexicographically_largest(ny);    // y < -y
        Aff<Fq2> p;
        p.x = x;
        p.y = (y_is_smaller != greatest) ? y : ny;
        if (is_zero(y) && is_zero(ny)) p.y = y;
        return jac_to_aff_host(out, mul_bits_host(p, G2_COFACTOR, 254));
    }
}
inline bool hash_to_g2(Aff<Fq2> &out, const uint8_t digest32[32]) {
    uint32_t seed[8];
    for (int i = 0; i < 8; i++) seed[i] = ((uint32_t)digest32[4 * i] << 24) | ((uint32_t)digest32[4 * i + 1] << 16) | ((uint32_t)digest32[4 * i + 2] << 8) | digest32[4 * i + 3];
    ChaChaRng rng(seed);
    return g2_rand(out, rng);
}

}  // namespace pairing
}  // namespace p2b
