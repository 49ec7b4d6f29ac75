// fp.cuh -- 256-bit Montgomery prime-field arithmetic for BN254 Fq / Fr on sm_100a.
//
// Replaces, for the GPU path, what the reference gets from `#[derive(PrimeField)]`
// (pairing/src/bn256/fq.rs:4-7, fr.rs:3-6; crates ff_ce 0.7.1 / ff_derive_ce 0.5.1): canonical
// residues in Montgomery form, R = 2^256.  Same values as the reference's 4 x u64 limbs, held
// here as 8 x u32 limbs in registers (limb i of the reference = limbs 2i, 2i+1 here).
//
// The multiplier keeps two independent accumulators (even / odd limbs of the multiplicand) so
// that every 32x32->64 product is one `mad.lo.cc` + `madc.hi.cc` pair (ptxas fuses the pair into
// a single IMAD.WIDE.U32 with carry predicate) and the two carry chains give the scheduler ILP.
// Everything also compiles for the host (portable path) so the per-point logic of the kernels
// can be unit-tested on a machine without a GPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define P2B_HD __host__ __device__ __forceinline__
#define P2B_D __device__ __forceinline__
#else
#define P2B_HD inline
#define P2B_D inline
#endif

namespace p2b {

// ------------------------------------------------------------------ constants
#define P2B_DEF_CONST(name, ...)                                   \
    static const uint32_t h_##name[8] = __VA_ARGS__;               \
    P2B_CONST_DEV(name, __VA_ARGS__)
#if defined(__CUDACC__)
#define P2B_CONST_DEV(name, ...) static __device__ __constant__ uint32_t d_##name[8] = __VA_ARGS__;
#else
#define P2B_CONST_DEV(name, ...)
#endif
#if defined(__CUDA_ARCH__)
#define P2B_C(name, i) d_##name[i]
#else
#define P2B_C(name, i) h_##name[i]
#endif

// q, R mod q, R^2 mod q  (fq.rs:4-7, fq.rs:39-44; R^2 derived, checked in tests/test_constants.py)
P2B_DEF_CONST(FQ_P, {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u})
P2B_DEF_CONST(FQ_ONE, {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u, 0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u})
P2B_DEF_CONST(FQ_R2, {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u, 0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u})
// r, R mod r, R^2 mod r  (fr.rs:3-6)
P2B_DEF_CONST(FR_P, {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u})
P2B_DEF_CONST(FR_ONE, {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u})
P2B_DEF_CONST(FR_R2, {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u, 0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u})

// q^2 (16 limbs): added to a difference of two products so that it stays non-negative and below q * 2^256 (lazy reduction in Fq2)
#if defined(__CUDACC__)
static __device__ __constant__ uint32_t d_FQ_P2[16] = {0x275d69b1u, 0x3b5458a2u, 0x09eac101u, 0xa602072du, 0x6d96cadcu, 0x4a50189cu, 0x7a1242c8u, 0x04689e95u,
                                                       0x34c6b38du, 0x26edfa5cu, 0x16375606u, 0xb00b8551u, 0x0348d21cu, 0x599a6f7cu, 0x763cbf9cu, 0x0925c4b8u};
#endif
struct FqP {
    static P2B_HD uint32_t p(int i) { return P2B_C(FQ_P, i); }
    static P2B_HD uint32_t one(int i) { return P2B_C(FQ_ONE, i); }
    static P2B_HD uint32_t r2(int i) { return P2B_C(FQ_R2, i); }
    static constexpr uint32_t inv = 0xe4866389u;  // -q^-1 mod 2^32
};
struct FrP {
    static P2B_HD uint32_t p(int i) { return P2B_C(FR_P, i); }
    static P2B_HD uint32_t one(int i) { return P2B_C(FR_ONE, i); }
    static P2B_HD uint32_t r2(int i) { return P2B_C(FR_R2, i); }
    static constexpr uint32_t inv = 0xefffffffu;  // -r^-1 mod 2^32
};

// ------------------------------------------------------------------ the field element
template <class P>
struct Fp {
    uint32_t l[8];
};
using Fq = Fp<FqP>;
using Fr = Fp<FrP>;

template <class P> P2B_HD Fp<P> fp_zero() { Fp<P> r; for (int i = 0; i < 8; i++) r.l[i] = 0; return r; }
template <class P> P2B_HD Fp<P> fp_one() { Fp<P> r; for (int i = 0; i < 8; i++) r.l[i] = P::one(i); return r; }
template <class P> P2B_HD bool is_zero(const Fp<P> &a) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= a.l[i];
    return o == 0;
}
template <class P> P2B_HD bool eq(const Fp<P> &a, const Fp<P> &b) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= a.l[i] ^ b.l[i];
    return o == 0;
}
// raw limb compare a >= b
P2B_HD bool limbs_geq(const uint32_t *a, const uint32_t *b) {
    // borrow of a - b
    uint32_t borrow = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t t = (uint64_t)a[i] - b[i] - borrow;
        borrow = (uint32_t)(t >> 63);
    }
    return borrow == 0;
}

// ------------------------------------------------------------------ add / sub
#if defined(__CUDA_ARCH__)
// r = a + b  (8 limbs, carry chain; no carry out for inputs < 2^255)
P2B_D void add8(uint32_t *r, const uint32_t *a, const uint32_t *b) {
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, %23;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
}
// r = a - b, returns borrow as 0 / 0xffffffff
P2B_D uint32_t sub8(uint32_t *r, const uint32_t *a, const uint32_t *b) {
    uint32_t br;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(br)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    return br;
}
#else
P2B_HD void add8(uint32_t *r, const uint32_t *a, const uint32_t *b) {
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) { c += (uint64_t)a[i] + b[i]; r[i] = (uint32_t)c; c >>= 32; }
}
P2B_HD uint32_t sub8(uint32_t *r, const uint32_t *a, const uint32_t *b) {
    uint64_t br = 0;
    for (int i = 0; i < 8; i++) { uint64_t t = (uint64_t)a[i] - b[i] - br; r[i] = (uint32_t)t; br = (t >> 63) & 1; }
    return br ? 0xffffffffu : 0u;
}
#endif

// conditional final subtraction: r in [0, 2p) -> [0, p)
template <class P> P2B_HD void reduce_once(uint32_t *r) {
    uint32_t t[8], m[8];
#pragma unroll
    for (int i = 0; i < 8; i++) m[i] = P::p(i);
    uint32_t br = sub8(t, r, m);
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = br ? r[i] : t[i];
}

template <class P> P2B_HD Fp<P> add(const Fp<P> &a, const Fp<P> &b) {
    Fp<P> r;
    add8(r.l, a.l, b.l);
    reduce_once<P>(r.l);
    return r;
}
template <class P> P2B_HD Fp<P> sub(const Fp<P> &a, const Fp<P> &b) {
    Fp<P> r;
    uint32_t br = sub8(r.l, a.l, b.l);
    uint32_t m[8];
#pragma unroll
    for (int i = 0; i < 8; i++) m[i] = P::p(i) & br;
    add8(r.l, r.l, m);
    return r;
}
template <class P> P2B_HD Fp<P> dbl(const Fp<P> &a) { return add(a, a); }
template <class P> P2B_HD Fp<P> neg(const Fp<P> &a) {
    Fp<P> r;
    uint32_t m[8];
#pragma unroll
    for (int i = 0; i < 8; i++) m[i] = P::p(i);
    sub8(r.l, m, a.l);
    bool z = is_zero(a);
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = z ? 0u : r.l[i];
    return r;
}
// a if !c else b (lane-wise select, no divergence)
template <class P> P2B_HD Fp<P> select(bool c, const Fp<P> &b, const Fp<P> &a) {
    Fp<P> r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = c ? b.l[i] : a.l[i];
    return r;
}
template <class P> P2B_HD Fp<P> cneg(const Fp<P> &a, bool c) { return select(c, neg(a), a); }

// ------------------------------------------------------------------ Montgomery multiplication
#if defined(__CUDA_ARCH__)
// acc[0..7] = {x0,x1,x2,x3} * y laid out as four 64-bit products, acc[8] = 0
P2B_D void row_mul(uint32_t *acc, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t y) {
    asm("mul.lo.u32 %0, %8, %12;\n\t"
        "mul.hi.u32 %1, %8, %12;\n\t"
        "mul.lo.u32 %2, %9, %12;\n\t"
        "mul.hi.u32 %3, %9, %12;\n\t"
        "mul.lo.u32 %4, %10, %12;\n\t"
        "mul.hi.u32 %5, %10, %12;\n\t"
        "mul.lo.u32 %6, %11, %12;\n\t"
        "mul.hi.u32 %7, %11, %12;"
        : "=r"(acc[0]), "=r"(acc[1]), "=r"(acc[2]), "=r"(acc[3]), "=r"(acc[4]), "=r"(acc[5]), "=r"(acc[6]), "=r"(acc[7])
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(y));
    acc[8] = 0;
}
// acc[0..7] += {x0..x3} * y ; acc[8] += carry
P2B_D void row_mad(uint32_t *acc, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t y) {
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
          "+r"(acc[7]), "+r"(acc[8])
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(y));
}
// lo += mrg (the overlapping word of the other accumulator), the carry enters the chain:
// acc[0..7] += {x0..x3} * y + carry ; acc[8] = carry out      (acc[8] is a fresh word)
P2B_D void row_mad_merge(uint32_t &lo, uint32_t mrg, uint32_t *acc, uint32_t x0, uint32_t x1, uint32_t x2,
                         uint32_t x3, uint32_t y) {
    asm("add.cc.u32 %9, %9, %10;\n\t"
        "madc.lo.cc.u32 %0, %11, %15, %0;\n\t"
        "madc.hi.cc.u32 %1, %11, %15, %1;\n\t"
        "madc.lo.cc.u32 %2, %12, %15, %2;\n\t"
        "madc.hi.cc.u32 %3, %12, %15, %3;\n\t"
        "madc.lo.cc.u32 %4, %13, %15, %4;\n\t"
        "madc.hi.cc.u32 %5, %13, %15, %5;\n\t"
        "madc.lo.cc.u32 %6, %14, %15, %6;\n\t"
        "madc.hi.cc.u32 %7, %14, %15, %7;\n\t"
        "addc.u32 %8, 0, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
          "+r"(acc[7]), "=r"(acc[8]), "+r"(lo)
        : "r"(mrg), "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(y));
}

// r = a * b / R mod p.
// E and O are two plain multi-word accumulators indexed by ABSOLUTE word position (fully unrolled into registers);
// the running total is E + O.  A 64-bit product a_j*b_i lands on words (i+j, i+j+1): rows whose first word is even go
// to E, rows whose first word is odd go to O, so E's register pairs are always (even, odd) and O's (odd, even) --
// IMAD.WIDE needs aligned pairs, and this keeps ptxas from shuffling registers between iterations.  In iteration i
// the accumulator whose pair starts at word i ("A") absorbs the other one's word i; the carry of that merge rides
// into the other accumulator's ("B") chain, which starts at word i+1.
template <class P>
P2B_D void mont_step(uint32_t *A, uint32_t *B, int i, const uint32_t *a, uint32_t bi) {
    B[i + 8] = 0;
    row_mad_merge(A[i], B[i], &B[i + 1], a[1], a[3], a[5], a[7], bi);       // B[i+1..i+8] += a_odd * b_i (+ merge carry)
    row_mad(&A[i], a[0], a[2], a[4], a[6], bi);                            // A[i..i+7]   += a_even * b_i
    uint32_t m = A[i] * P::inv;
    row_mad(&A[i], P::p(0), P::p(2), P::p(4), P::p(6), m);                 // A[i] becomes 0
    row_mad(&B[i + 1], P::p(1), P::p(3), P::p(5), P::p(7), m);
}
template <class P> P2B_D void mont_mul(uint32_t *r, const uint32_t *a, const uint32_t *b) {
    uint32_t E[18], O[18];
    row_mul(&E[0], a[0], a[2], a[4], a[6], b[0]);
    row_mul(&O[1], a[1], a[3], a[5], a[7], b[0]);
    {
        uint32_t m = E[0] * P::inv;
        row_mad(&E[0], P::p(0), P::p(2), P::p(4), P::p(6), m);
        row_mad(&O[1], P::p(1), P::p(3), P::p(5), P::p(7), m);
    }
    mont_step<P>(O, E, 1, a, b[1]);
    mont_step<P>(E, O, 2, a, b[2]);
    mont_step<P>(O, E, 3, a, b[3]);
    mont_step<P>(E, O, 4, a, b[4]);
    mont_step<P>(O, E, 5, a, b[5]);
    mont_step<P>(E, O, 6, a, b[6]);
    mont_step<P>(O, E, 7, a, b[7]);
    add8(r, &E[8], &O[8]);
    reduce_once<P>(r);
}

// r = (a * b + c * d) / R mod p: two products under ONE interleaved reduction (192 + 8 instead of 2 x 136 wide multiplies).  The
// sum of the products stays below 2 p^2 < p 2^256, so the result is below 2p before the final conditional subtraction.  Used by
// the two-lanes-per-element Fq2 arithmetic (msm_g2x2.cuh): each lane needs a0 b0 - a1 b1 or a0 b1 + a1 b0.
template <class P>
P2B_D void mont_step2(uint32_t *A, uint32_t *B, int i, const uint32_t *a, uint32_t bi, const uint32_t *c, uint32_t di) {
    B[i + 8] = 0;
    row_mad_merge(A[i], B[i], &B[i + 1], a[1], a[3], a[5], a[7], bi);       // B[i+1..i+8] += a_odd * b_i (+ merge carry), B[i+9] = carry
    row_mad(&B[i + 1], c[1], c[3], c[5], c[7], di);
    row_mad(&A[i], a[0], a[2], a[4], a[6], bi);
    row_mad(&A[i], c[0], c[2], c[4], c[6], di);
    uint32_t m = A[i] * P::inv;
    row_mad(&A[i], P::p(0), P::p(2), P::p(4), P::p(6), m);                 // A[i] becomes 0
    row_mad(&B[i + 1], P::p(1), P::p(3), P::p(5), P::p(7), m);
}
template <class P> P2B_D void mont_mul2(uint32_t *r, const uint32_t *a, const uint32_t *b, const uint32_t *c, const uint32_t *d) {
    uint32_t E[18], O[18];
    row_mul(&E[0], a[0], a[2], a[4], a[6], b[0]);
    row_mul(&O[1], a[1], a[3], a[5], a[7], b[0]);
    row_mad(&E[0], c[0], c[2], c[4], c[6], d[0]);
    row_mad(&O[1], c[1], c[3], c[5], c[7], d[0]);
    {
        uint32_t m = E[0] * P::inv;
        row_mad(&E[0], P::p(0), P::p(2), P::p(4), P::p(6), m);
        row_mad(&O[1], P::p(1), P::p(3), P::p(5), P::p(7), m);
    }
    mont_step2<P>(O, E, 1, a, b[1], c, d[1]);
    mont_step2<P>(E, O, 2, a, b[2], c, d[2]);
    mont_step2<P>(O, E, 3, a, b[3], c, d[3]);
    mont_step2<P>(E, O, 4, a, b[4], c, d[4]);
    mont_step2<P>(O, E, 5, a, b[5], c, d[5]);
    mont_step2<P>(E, O, 6, a, b[6], c, d[6]);
    mont_step2<P>(O, E, 7, a, b[7], c, d[7]);
    add8(r, &E[8], &O[8]);
    reduce_once<P>(r);
}

// ---- dedicated squaring: 28 cross products (doubled) + 8 squares + the 64 products of the reduction = 100 wide
// multiplies instead of 128.  The cross products a_i * a_j (i < j) land on words (i+j, i+j+1); those with i+j even are
// summed in E, those with i+j odd in O (aligned register pairs again).  Rows are issued in an order in which the word
// above each row is still untouched, so a row's carry chain ends in a fresh word.
P2B_D void row_mad1(uint32_t *acc, uint32_t x0, uint32_t y) {
    asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
        "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
        "addc.u32 %2, %2, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2])
        : "r"(x0), "r"(y));
}
P2B_D void row_mad2(uint32_t *acc, uint32_t x0, uint32_t x1, uint32_t y) {
    asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"
        "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
        "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
        "addc.u32 %4, %4, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4])
        : "r"(x0), "r"(x1), "r"(y));
}
P2B_D void row_mad3(uint32_t *acc, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t y) {
    asm("mad.lo.cc.u32 %0, %7, %10, %0;\n\t"
        "madc.hi.cc.u32 %1, %7, %10, %1;\n\t"
        "madc.lo.cc.u32 %2, %8, %10, %2;\n\t"
        "madc.hi.cc.u32 %3, %8, %10, %3;\n\t"
        "madc.lo.cc.u32 %4, %9, %10, %4;\n\t"
        "madc.hi.cc.u32 %5, %9, %10, %5;\n\t"
        "addc.u32 %6, %6, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6])
        : "r"(x0), "r"(x1), "r"(x2), "r"(y));
}
// r[0..7] = a[0..7] + b[0..7] + cin (0 / 1), returns the carry out
P2B_D uint32_t add8c(uint32_t *r, const uint32_t *a, const uint32_t *b, uint32_t cin) {
    uint32_t co;
    asm("add.cc.u32 %8, %25, 0xffffffff;\n\t"
        "addc.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=&r"(co)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]), "r"(cin));
    return co;
}
// t[0..15] += a_i^2 at words (2i, 2i+1): one carry chain over all 16 words
P2B_D void add_squares(uint32_t *t, const uint32_t *a) {
    asm("mad.lo.cc.u32 %0, %16, %16, %0;\n\t"
        "madc.hi.cc.u32 %1, %16, %16, %1;\n\t"
        "madc.lo.cc.u32 %2, %17, %17, %2;\n\t"
        "madc.hi.cc.u32 %3, %17, %17, %3;\n\t"
        "madc.lo.cc.u32 %4, %18, %18, %4;\n\t"
        "madc.hi.cc.u32 %5, %18, %18, %5;\n\t"
        "madc.lo.cc.u32 %6, %19, %19, %6;\n\t"
        "madc.hi.cc.u32 %7, %19, %19, %7;\n\t"
        "madc.lo.cc.u32 %8, %20, %20, %8;\n\t"
        "madc.hi.cc.u32 %9, %20, %20, %9;\n\t"
        "madc.lo.cc.u32 %10, %21, %21, %10;\n\t"
        "madc.hi.cc.u32 %11, %21, %21, %11;\n\t"
        "madc.lo.cc.u32 %12, %22, %22, %12;\n\t"
        "madc.hi.cc.u32 %13, %22, %22, %13;\n\t"
        "madc.lo.cc.u32 %14, %23, %23, %14;\n\t"
        "madc.hi.u32 %15, %23, %23, %15;"
        : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]),
          "+r"(t[8]), "+r"(t[9]), "+r"(t[10]), "+r"(t[11]), "+r"(t[12]), "+r"(t[13]), "+r"(t[14]), "+r"(t[15])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]));
}
// one word-serial reduction step on the low half (no multiplicand rows): same accumulator discipline as mont_step
template <class P> P2B_D void red_step(uint32_t *A, uint32_t *B, int i) {
    B[i + 8] = 0;
    const uint32_t m = (A[i] + B[i]) * P::inv;
    row_mad_merge(A[i], B[i], &B[i + 1], P::p(1), P::p(3), P::p(5), P::p(7), m);
    row_mad(&A[i], P::p(0), P::p(2), P::p(4), P::p(6), m);                 // A[i] becomes 0
}
template <class P> P2B_D void mont_sqr(uint32_t *r, const uint32_t *a) {
    uint32_t E[18], O[18];
#pragma unroll
    for (int i = 0; i < 18; i++) { E[i] = 0; O[i] = 0; }
    // i + j odd
    row_mad1(&O[1], a[0], a[1]);
    row_mad1(&O[3], a[1], a[2]);
    row_mad2(&O[3], a[0], a[2], a[3]);
    row_mad2(&O[5], a[1], a[3], a[4]);
    row_mad3(&O[5], a[0], a[2], a[4], a[5]);
    row_mad3(&O[7], a[1], a[3], a[5], a[6]);
    row_mad(&O[7], a[0], a[2], a[4], a[6], a[7]);
    // i + j even
    row_mad1(&E[2], a[0], a[2]);
    row_mad1(&E[4], a[1], a[3]);
    row_mad2(&E[4], a[0], a[2], a[4]);
    row_mad2(&E[6], a[1], a[3], a[5]);
    row_mad3(&E[6], a[0], a[2], a[4], a[6]);
    row_mad3(&E[8], a[1], a[3], a[5], a[7]);
    // t = 2 (E + O) + squares
    uint32_t s[16], t[16];
    uint32_t c = add8c(&s[0], &E[0], &O[0], 0);
    add8c(&s[8], &E[8], &O[8], c);
    t[0] = s[0] << 1;
#pragma unroll
    for (int i = 1; i < 16; i++) t[i] = __funnelshift_l(s[i - 1], s[i], 1);
    add_squares(t, a);
    // reduce the low half word-serially; the high half joins at the end:  a^2 / R = H + (L + sum m_i p 2^(32 i)) / R
#pragma unroll
    for (int i = 0; i < 8; i++) { E[i] = t[i]; O[i] = 0; }
#pragma unroll
    for (int i = 8; i < 18; i++) { E[i] = 0; O[i] = 0; }
    {
        const uint32_t m = E[0] * P::inv;
        row_mad(&E[0], P::p(0), P::p(2), P::p(4), P::p(6), m);
        row_mad(&O[1], P::p(1), P::p(3), P::p(5), P::p(7), m);
    }
    red_step<P>(O, E, 1);
    red_step<P>(E, O, 2);
    red_step<P>(O, E, 3);
    red_step<P>(E, O, 4);
    red_step<P>(O, E, 5);
    red_step<P>(E, O, 6);
    red_step<P>(O, E, 7);
    uint32_t u[8];
    add8(u, &E[8], &O[8]);
    add8(r, u, &t[8]);
    reduce_once<P>(r);
}
#else
template <class P> P2B_HD void mont_mul(uint32_t *r, const uint32_t *a, const uint32_t *b) {
    uint32_t t[10] = {0};
    for (int i = 0; i < 8; i++) {
        uint64_t c = 0;
        for (int j = 0; j < 8; j++) { c += (uint64_t)a[j] * b[i] + t[j]; t[j] = (uint32_t)c; c >>= 32; }
        c += t[8]; t[8] = (uint32_t)c; t[9] = (uint32_t)(c >> 32);
        uint32_t m = t[0] * P::inv;
        c = ((uint64_t)m * P::p(0) + t[0]) >> 32;
        for (int j = 1; j < 8; j++) { c += (uint64_t)m * P::p(j) + t[j]; t[j - 1] = (uint32_t)c; c >>= 32; }
        c += t[8]; t[7] = (uint32_t)c; t[8] = t[9] + (uint32_t)(c >> 32);
    }
    for (int i = 0; i < 8; i++) r[i] = t[i];
    reduce_once<P>(r);
}
#endif

template <class P> P2B_HD Fp<P> mul(const Fp<P> &a, const Fp<P> &b) {
    Fp<P> r;
    mont_mul<P>(r.l, a.l, b.l);
    return r;
}
template <class P> P2B_HD Fp<P> sqr(const Fp<P> &a) {
    // The dedicated squaring (100 wide multiplies instead of 128) measured no faster than mul(a, a) on B200: the kernels
    // run 2-3 warps per scheduler and are limited by the latency of the carry chains, and the squaring's phases (cross
    // products, doubling, squares, reduction) expose less instruction-level parallelism than the interleaved multiplier.
    // Kept for tuning builds (-DP2B_USE_SQR); see DESIGN.md section 4.
#if defined(__CUDA_ARCH__) && defined(P2B_USE_SQR)
    Fp<P> r;
    mont_sqr<P>(r.l, a.l);
    return r;
#else
    return mul(a, a);
#endif
}

// ---- wide (unreduced) products and a stand-alone Montgomery reduction: the building blocks of lazy reduction in Fq2, where
// (a0 + a1 u)(b0 + b1 u) needs three 512-bit products but only TWO reductions (csrc/ec.cuh mul(Fq2, Fq2)).
#if defined(__CUDA_ARCH__)
// t[0..15] = a * b.  Same two-accumulator discipline as mont_mul (aligned register pairs for IMAD.WIDE), no reduction rows.
P2B_D void wide_mul(uint32_t *t, const uint32_t *a, const uint32_t *b) {
    uint32_t E[18], O[18];
    row_mul(&E[0], a[0], a[2], a[4], a[6], b[0]);
    row_mul(&O[1], a[1], a[3], a[5], a[7], b[0]);
#pragma unroll
    for (int i = 9; i < 18; i++) E[i] = 0;
    O[0] = 0;
#pragma unroll
    for (int i = 10; i < 18; i++) O[i] = 0;
#pragma unroll
    for (int i = 1; i < 8; i++) {
        if (i & 1) {
            row_mad(&O[i], a[0], a[2], a[4], a[6], b[i]);
            row_mad(&E[i + 1], a[1], a[3], a[5], a[7], b[i]);
        } else {
            row_mad(&E[i], a[0], a[2], a[4], a[6], b[i]);
            row_mad(&O[i + 1], a[1], a[3], a[5], a[7], b[i]);
        }
    }
    const uint32_t c = add8c(&t[0], &E[0], &O[0], 0);
    add8c(&t[8], &E[8], &O[8], c);
}
// r[0..15] = a + b / a - b over 16 words (callers guarantee no overflow / no negative result)
P2B_D void add16(uint32_t *r, const uint32_t *a, const uint32_t *b) {
    const uint32_t c = add8c(&r[0], &a[0], &b[0], 0);
    add8c(&r[8], &a[8], &b[8], c);
}
P2B_D void sub16(uint32_t *r, const uint32_t *a, const uint32_t *b) {
    asm("sub.cc.u32 %0, %16, %32;\n\t"
        "subc.cc.u32 %1, %17, %33;\n\t"
        "subc.cc.u32 %2, %18, %34;\n\t"
        "subc.cc.u32 %3, %19, %35;\n\t"
        "subc.cc.u32 %4, %20, %36;\n\t"
        "subc.cc.u32 %5, %21, %37;\n\t"
        "subc.cc.u32 %6, %22, %38;\n\t"
        "subc.cc.u32 %7, %23, %39;\n\t"
        "subc.cc.u32 %8, %24, %40;\n\t"
        "subc.cc.u32 %9, %25, %41;\n\t"
        "subc.cc.u32 %10, %26, %42;\n\t"
        "subc.cc.u32 %11, %27, %43;\n\t"
        "subc.cc.u32 %12, %28, %44;\n\t"
        "subc.cc.u32 %13, %29, %45;\n\t"
        "subc.cc.u32 %14, %30, %46;\n\t"
        "subc.u32 %15, %31, %47;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(a[8]), "r"(a[9]), "r"(a[10]),
          "r"(a[11]), "r"(a[12]), "r"(a[13]), "r"(a[14]), "r"(a[15]), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]),
          "r"(b[6]), "r"(b[7]), "r"(b[8]), "r"(b[9]), "r"(b[10]), "r"(b[11]), "r"(b[12]), "r"(b[13]), "r"(b[14]), "r"(b[15]));
}
// r = t / R mod p for t < p * 2^256 (result < 2p before the final conditional subtraction): the reduction phase of mont_sqr
template <class P> P2B_D void mont_red(uint32_t *r, const uint32_t *t) {
    uint32_t E[18], O[18];
#pragma unroll
    for (int i = 0; i < 8; i++) { E[i] = t[i]; O[i] = 0; }
#pragma unroll
    for (int i = 8; i < 18; i++) { E[i] = 0; O[i] = 0; }
    {
        const uint32_t m = E[0] * P::inv;
        row_mad(&E[0], P::p(0), P::p(2), P::p(4), P::p(6), m);
        row_mad(&O[1], P::p(1), P::p(3), P::p(5), P::p(7), m);
    }
    red_step<P>(O, E, 1);
    red_step<P>(E, O, 2);
    red_step<P>(O, E, 3);
    red_step<P>(E, O, 4);
    red_step<P>(O, E, 5);
    red_step<P>(E, O, 6);
    red_step<P>(O, E, 7);
    uint32_t u[8];
    add8(u, &E[8], &O[8]);
    add8(r, u, &t[8]);
    reduce_once<P>(r);
}
#endif

// dedicated squaring wherever the caller asks for it explicitly (device: 100 instead of 128 wide multiplies)
template <class P> P2B_HD Fp<P> sqr_ded(const Fp<P> &a) {
#if defined(__CUDA_ARCH__)
    Fp<P> r;
    mont_sqr<P>(r.l, a.l);
    return r;
#else
    return mul(a, a);
#endif
}

// a * b + c * d with one reduction (device: mont_mul2, 200 instead of 272 wide multiplies)
template <class P> P2B_HD Fp<P> mul2_add(const Fp<P> &a, const Fp<P> &b, const Fp<P> &c, const Fp<P> &d) {
#if defined(__CUDA_ARCH__)
    Fp<P> r;
    mont_mul2<P>(r.l, a.l, b.l, c.l, d.l);
    return r;
#else
    return add(mul(a, b), mul(c, d));
#endif
}

// canonical <-> Montgomery
template <class P> P2B_HD Fp<P> to_mont(const Fp<P> &a) {
    Fp<P> r2;
#pragma unroll
    for (int i = 0; i < 8; i++) r2.l[i] = P::r2(i);
    return mul(a, r2);
}
template <class P> P2B_HD Fp<P> from_mont(const Fp<P> &a) {
    Fp<P> one = fp_zero<P>();
    one.l[0] = 1;
    return mul(a, one);
}
// true if the raw limbs are a canonical residue (< p)
template <class P> P2B_HD bool is_canonical(const Fp<P> &a) {
    uint32_t m[8];
#pragma unroll
    for (int i = 0; i < 8; i++) m[i] = P::p(i);
    return !limbs_geq(a.l, m);
}
// a > b comparing canonical integer values (inputs Montgomery) -- ff_derive Ord
template <class P> P2B_HD bool gt_canonical(const Fp<P> &a, const Fp<P> &b) {
    Fp<P> ca = from_mont(a), cb = from_mont(b);
    return !limbs_geq(cb.l, ca.l);  // a > b  <=>  !(b >= a)
}
// y > -y on canonical values, i.e. y > (p - y)  (the compressed-point sign bit, ec.rs:934-941)
template <class P> P2B_HD bool is_lexicographically_largest(const Fp<P> &y_mont) {
    Fp<P> c = from_mont(y_mont);
    if (is_zero(c)) return false;
    uint32_t m[8], t[8];
#pragma unroll
    for (int i = 0; i < 8; i++) m[i] = P::p(i);
    sub8(t, m, c.l);                   // p - y
    return !limbs_geq(t, c.l);         // y > p - y
}

// a^e for an exponent held as 8 little-endian u32 limbs in memory (MSB first; uniform across lanes)
template <class P> P2B_HD Fp<P> pow_limbs(const Fp<P> &a, const uint32_t *e) {
    Fp<P> res = fp_one<P>();
    bool started = false;
    for (int i = 255; i >= 0; i--) {
        bool bit = (e[i >> 5] >> (i & 31)) & 1;
        if (started) res = sqr(res);
        if (bit) { res = started ? mul(res, a) : a; started = true; }
    }
    return res;
}
// Fermat inverse a^(p-2); inverse(0) = 0.  (p-2 differs from p only in limb 0.)
template <class P> P2B_HD Fp<P> inv(const Fp<P> &a) {
    Fp<P> res = fp_one<P>();
    bool started = false;
    for (int i = 255; i >= 0; i--) {
        uint32_t w = (i < 32) ? (P::p(0) - 2u) : P::p(i >> 5);
        bool bit = (w >> (i & 31)) & 1;
        if (started) res = sqr(res);
        if (bit) { res = started ? mul(res, a) : a; started = true; }
    }
    return started ? res : res;
}
// a^k for a 64-bit exponent (tau^index)
template <class P> P2B_HD Fp<P> pow_u64(const Fp<P> &a, uint64_t k) {
    Fp<P> res = fp_one<P>();
    bool started = false;
    for (int i = 63; i >= 0; i--) {
        bool bit = (k >> i) & 1;
        if (started) res = sqr(res);
        if (bit) { res = started ? mul(res, a) : a; started = true; }
    }
    return res;
}

// ------------------------------------------------------------------ byte codecs (big-endian wire form)
P2B_HD uint32_t bswap32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(v, 0, 0x0123);
#else
    return (v >> 24) | ((v >> 8) & 0xff00u) | ((v << 8) & 0xff0000u) | (v << 24);
#endif
}
// 32 BE bytes at a 4-byte aligned address, already loaded as 8 native words w[0..7] (w[0] = first 4 bytes)
template <class P> P2B_HD Fp<P> limbs_from_be_words(const uint32_t *w) {
    Fp<P> r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = bswap32(w[7 - i]);
    return r;
}
template <class P> P2B_HD void limbs_to_be_words(const Fp<P> &a, uint32_t *w) {
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = bswap32(a.l[7 - i]);
}

}  // namespace p2b
