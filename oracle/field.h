/* oracle/field.h -- CPU restatement of the reference's prime-field arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/README.md): never linked into the product library.
 *
 * The reference's Fq / Fr are `#[derive(PrimeField)]` types (pairing/src/bn256/fq.rs:4-7,
 * pairing/src/bn256/fr.rs:3-6) expanded by the crates ff_ce 0.7.1 + ff_derive_ce 0.5.1, which
 * are NOT vendored under /root/reference (powersoftau/Cargo.lock:147-160,511-512).  Their
 * published algorithm is restated here: 4 x u64 little-endian limbs, Montgomery form with
 * R = 2^256, every stored value a canonical residue < modulus:
 *   add_assign   add_nocarry, then subtract the modulus if the result is not < modulus
 *   sub_assign   if other > self add the modulus first, then sub_noborrow
 *   negate       modulus - self (0 stays 0)
 *   mul_assign   4x4 schoolbook into 8 limbs, then mont_reduce (4 rounds of k = r[i]*inv)
 *   square       same result as mul_assign(self, self)
 *   inverse      binary extended Euclid (Guajardo-Kumar-Paar-Pelzl alg. 16) on the Montgomery
 *                residue, started with b = R^2 so the result is again in Montgomery form
 *   pow          MSB-first square-and-multiply over a little-endian u64 exponent
 *   from_repr    reject >= modulus, then multiply by R^2;   into_repr = mont_reduce(self, 0)
 *   sqrt (Fq)    q = 3 mod 4: a1 = a^((q-3)/4), a0 = a1^2 * a; a0 == -1 -> None else a1 * a
 * The constants are pinned by the values hard-coded in the reference (fq.rs:11-16 = 3R,
 * fq.rs:39-50 = R and 2R, fq.rs:434-439 = -R); tests/test_oracle.py checks them.
 */
#ifndef P2B_ORACLE_FIELD_H
#define P2B_ORACLE_FIELD_H
#include <stdint.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fe;

typedef struct {
    uint64_t m[4];    /* modulus */
    uint64_t inv;     /* -m^-1 mod 2^64 */
    uint64_t one[4];  /* R mod m */
    uint64_t r2[4];   /* R^2 mod m */
} fparams;

static const fparams FQ = {
    {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL},
    0x87d20782e4866389ULL,
    {0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL},
    {0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL}};

static const fparams FR = {
    {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL},
    0xc2e1f593efffffffULL,
    {0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL},
    {0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL}};

/* ---- raw 256-bit helpers (PrimeFieldRepr) ---- */
static inline int r_is_zero(const uint64_t *a) { return (a[0] | a[1] | a[2] | a[3]) == 0; }
static inline int r_cmp(const uint64_t *a, const uint64_t *b) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] < b[i]) return -1;
        if (a[i] > b[i]) return 1;
    }
    return 0;
}
static inline uint64_t r_add(uint64_t *a, const uint64_t *b) { /* add_nocarry */
    u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)a[i] + b[i]; a[i] = (uint64_t)c; c >>= 64; }
    return (uint64_t)c;
}
static inline uint64_t r_sub(uint64_t *a, const uint64_t *b) { /* sub_noborrow */
    uint64_t br = 0;
    for (int i = 0; i < 4; i++) {
        u128 t = (u128)a[i] - b[i] - br;
        a[i] = (uint64_t)t;
        br = (uint64_t)(t >> 64) & 1;
    }
    return br;
}
static inline void r_div2(uint64_t *a) {
    for (int i = 0; i < 3; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 63);
    a[3] >>= 1;
}
static inline void r_shr(uint64_t *a, unsigned n) {
    if (n >= 256) { memset(a, 0, 32); return; }
    while (n >= 64) { a[0] = a[1]; a[1] = a[2]; a[2] = a[3]; a[3] = 0; n -= 64; }
    if (n) {
        for (int i = 0; i < 3; i++) a[i] = (a[i] >> n) | (a[i + 1] << (64 - n));
        a[3] >>= n;
    }
}
static inline void r_from_be(uint64_t *a, const uint8_t *b) { /* read_be */
    for (int i = 0; i < 4; i++) {
        uint64_t v = 0;
        for (int j = 0; j < 8; j++) v = (v << 8) | b[(3 - i) * 8 + j];
        a[i] = v;
    }
}
static inline void r_to_be(const uint64_t *a, uint8_t *b) { /* write_be */
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 8; j++) b[(3 - i) * 8 + j] = (uint8_t)(a[i] >> (56 - 8 * j));
}

/* ---- field ops ---- */
static inline fe f_zero(void) { fe z = {{0, 0, 0, 0}}; return z; }
static inline fe f_one(const fparams *p) { fe o; memcpy(o.l, p->one, 32); return o; }
static inline int f_is_zero(const fe *a) { return r_is_zero(a->l); }
static inline int f_eq(const fe *a, const fe *b) { return r_cmp(a->l, b->l) == 0; }

static inline void f_add(fe *a, const fe *b, const fparams *p) {
    r_add(a->l, b->l);                               /* no carry: 2*(m-1) < 2^256 */
    if (r_cmp(a->l, p->m) >= 0) r_sub(a->l, p->m);
}
static inline void f_dbl(fe *a, const fparams *p) { fe t = *a; f_add(a, &t, p); }
static inline void f_sub(fe *a, const fe *b, const fparams *p) {
    if (r_cmp(b->l, a->l) > 0) r_add(a->l, p->m);
    r_sub(a->l, b->l);
}
static inline void f_neg(fe *a, const fparams *p) {
    if (!f_is_zero(a)) { fe t; memcpy(t.l, p->m, 32); r_sub(t.l, a->l); *a = t; }
}
static inline void f_mont_reduce(fe *out, uint64_t r[8], const fparams *p) {
    uint64_t carry2 = 0;
    for (int i = 0; i < 4; i++) {
        uint64_t k = r[i] * p->inv;
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)k * p->m[j] + r[i + j];
            r[i + j] = (uint64_t)c;
            c >>= 64;
        }
        c += (u128)r[i + 4] + carry2;
        r[i + 4] = (uint64_t)c;
        carry2 = (uint64_t)(c >> 64);
    }
    memcpy(out->l, r + 4, 32);
    if (carry2 || r_cmp(out->l, p->m) >= 0) r_sub(out->l, p->m);
}
static inline void f_mul(fe *a, const fe *b, const fparams *p) {
    uint64_t r[8] = {0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)a->l[i] * b->l[j] + r[i + j];
            r[i + j] = (uint64_t)c;
            c >>= 64;
        }
        r[i + 4] = (uint64_t)c;
    }
    f_mont_reduce(a, r, p);
}
static inline void f_sqr(fe *a, const fparams *p) { fe t = *a; f_mul(a, &t, p); }

static inline int f_from_repr(fe *out, const uint64_t repr[4], const fparams *p) {
    if (r_cmp(repr, p->m) >= 0) return 0;            /* PrimeFieldDecodingError::NotInField */
    fe r2; memcpy(r2.l, p->r2, 32);
    memcpy(out->l, repr, 32);
    f_mul(out, &r2, p);
    return 1;
}
static inline void f_into_repr(uint64_t repr[4], const fe *a, const fparams *p) {
    uint64_t r[8] = {a->l[0], a->l[1], a->l[2], a->l[3], 0, 0, 0, 0};
    fe t; f_mont_reduce(&t, r, p);
    memcpy(repr, t.l, 32);
}
/* pow over little-endian u64 limbs, MSB first (ff Field::pow) */
static inline fe f_pow(const fe *a, const uint64_t *e, int nlimbs, const fparams *p) {
    fe res = f_one(p);
    int found = 0;
    for (int i = nlimbs * 64 - 1; i >= 0; i--) {
        int bit = (e[i / 64] >> (i % 64)) & 1;
        if (found) f_sqr(&res, p); else found = bit;
        if (bit) f_mul(&res, a, p);
    }
    return res;
}
/* binary extended Euclid, ff_derive `inverse` */
static inline int f_inv(fe *out, const fe *a, const fparams *p) {
    if (f_is_zero(a)) return 0;
    uint64_t one[4] = {1, 0, 0, 0};
    uint64_t u[4], v[4];
    memcpy(u, a->l, 32); memcpy(v, p->m, 32);
    fe b, c = f_zero();
    memcpy(b.l, p->r2, 32);                          /* avoids a final Montgomery fix-up */
    while (r_cmp(u, one) != 0 && r_cmp(v, one) != 0) {
        while ((u[0] & 1) == 0) {
            r_div2(u);
            if (b.l[0] & 1) { uint64_t cy = r_add(b.l, p->m); r_div2(b.l); b.l[3] |= cy << 63; }
            else r_div2(b.l);
        }
        while ((v[0] & 1) == 0) {
            r_div2(v);
            if (c.l[0] & 1) { uint64_t cy = r_add(c.l, p->m); r_div2(c.l); c.l[3] |= cy << 63; }
            else r_div2(c.l);
        }
        if (r_cmp(v, u) < 0) { r_sub(u, v); f_sub(&b, &c, p); }
        else { r_sub(v, u); f_sub(&c, &b, p); }
    }
    *out = (r_cmp(u, one) == 0) ? b : c;
    return 1;
}
/* canonical-value ordering (ff_derive Ord compares into_repr()) */
static inline int f_cmp(const fe *a, const fe *b, const fparams *p) {
    uint64_t ra[4], rb[4];
    f_into_repr(ra, a, p); f_into_repr(rb, b, p);
    return r_cmp(ra, rb);
}
static inline int fq_sqrt(fe *out, const fe *a) {
    /* (q-3)/4, pairing/src/bn256/fq2.rs:218-223 holds the same limbs */
    static const uint64_t e[4] = {0x4f082305b61f3f51ULL, 0x65e05aa45a1c72a3ULL,
                                  0x6e14116da0605617ULL, 0x0c19139cb84c680aULL};
    fe a1 = f_pow(a, e, 4, &FQ);
    fe a0 = a1; f_sqr(&a0, &FQ); f_mul(&a0, a, &FQ);
    fe m1 = f_one(&FQ); f_neg(&m1, &FQ);
    if (f_eq(&a0, &m1)) return 0;
    f_mul(&a1, a, &FQ);
    *out = a1;
    return 1;
}

/* ---- uniform wrappers so ec_tmpl.h can be instantiated over Fq and Fq2 ---- */
typedef fe fq;
static inline fq fq_zero(void) { return f_zero(); }
static inline fq fq_one(void) { return f_one(&FQ); }
static inline int fq_is_zero(const fq *a) { return f_is_zero(a); }
static inline int fq_eq(const fq *a, const fq *b) { return f_eq(a, b); }
static inline void fq_add(fq *a, const fq *b) { f_add(a, b, &FQ); }
static inline void fq_sub(fq *a, const fq *b) { f_sub(a, b, &FQ); }
static inline void fq_dbl(fq *a) { f_dbl(a, &FQ); }
static inline void fq_neg(fq *a) { f_neg(a, &FQ); }
static inline void fq_mul(fq *a, const fq *b) { f_mul(a, b, &FQ); }
static inline void fq_sqr(fq *a) { f_sqr(a, &FQ); }
static inline int fq_inv(fq *o, const fq *a) { return f_inv(o, a, &FQ); }
/* a > b on canonical values */
static inline int fq_gt(const fq *a, const fq *b) { return f_cmp(a, b, &FQ) > 0; }

/* Fq2 = Fq[u]/(u^2+1), pairing/src/bn256/fq2.rs */
typedef struct { fq c0, c1; } fq2;
static inline fq2 fq2_zero(void) { fq2 z = {f_zero(), f_zero()}; return z; }
static inline fq2 fq2_one(void) { fq2 o = {f_one(&FQ), f_zero()}; return o; }
static inline int fq2_is_zero(const fq2 *a) { return f_is_zero(&a->c0) && f_is_zero(&a->c1); }
static inline int fq2_eq(const fq2 *a, const fq2 *b) { return f_eq(&a->c0, &b->c0) && f_eq(&a->c1, &b->c1); }
static inline void fq2_add(fq2 *a, const fq2 *b) { fq_add(&a->c0, &b->c0); fq_add(&a->c1, &b->c1); }
static inline void fq2_sub(fq2 *a, const fq2 *b) { fq_sub(&a->c0, &b->c0); fq_sub(&a->c1, &b->c1); }
static inline void fq2_dbl(fq2 *a) { fq_dbl(&a->c0); fq_dbl(&a->c1); }
static inline void fq2_neg(fq2 *a) { fq_neg(&a->c0); fq_neg(&a->c1); }
static inline void fq2_mul(fq2 *a, const fq2 *b) {       /* fq2.rs:167-180 (Karatsuba) */
    fq aa = a->c0; fq_mul(&aa, &b->c0);
    fq bb = a->c1; fq_mul(&bb, &b->c1);
    fq o = b->c0; fq_add(&o, &b->c1);
    fq_add(&a->c1, &a->c0);
    fq_mul(&a->c1, &o);
    fq_sub(&a->c1, &aa);
    fq_sub(&a->c1, &bb);
    a->c0 = aa; fq_sub(&a->c0, &bb);
}
static inline void fq2_sqr(fq2 *a) {                     /* fq2.rs:131-145 (complex squaring) */
    fq ab = a->c0; fq_mul(&ab, &a->c1);
    fq c0c1 = a->c0; fq_add(&c0c1, &a->c1);
    fq c0 = a->c1; fq_neg(&c0); fq_add(&c0, &a->c0);
    fq_mul(&c0, &c0c1);
    fq_sub(&c0, &ab);
    a->c1 = ab; fq_add(&a->c1, &ab);
    fq_add(&c0, &ab);
    a->c0 = c0;
}
static inline int fq2_inv(fq2 *o, const fq2 *a) {        /* fq2.rs:182-199 */
    fq t1 = a->c1; fq_sqr(&t1);
    fq t0 = a->c0; fq_sqr(&t0);
    fq_add(&t0, &t1);
    fq t;
    if (!fq_inv(&t, &t0)) return 0;
    *o = *a;
    fq_mul(&o->c0, &t); fq_mul(&o->c1, &t); fq_neg(&o->c1);
    return 1;
}
static inline int fq2_gt(const fq2 *a, const fq2 *b) {   /* fq2.rs:21-30: c1 first, then c0 */
    int c = f_cmp(&a->c1, &b->c1, &FQ);
    if (c != 0) return c > 0;
    return f_cmp(&a->c0, &b->c0, &FQ) > 0;
}
static inline fq2 fq2_pow(const fq2 *a, const uint64_t *e, int nlimbs) {
    fq2 res = fq2_one();
    int found = 0;
    for (int i = nlimbs * 64 - 1; i >= 0; i--) {
        int bit = (e[i / 64] >> (i % 64)) & 1;
        if (found) fq2_sqr(&res); else found = bit;
        if (bit) fq2_mul(&res, a);
    }
    return res;
}
static inline int fq2_sqrt(fq2 *out, const fq2 *a) {     /* fq2.rs:206-262, Algorithm 9 of eprint 2012/685 */
    if (fq2_is_zero(a)) { *out = fq2_zero(); return 1; }
    static const uint64_t e1[4] = {0x4f082305b61f3f51ULL, 0x65e05aa45a1c72a3ULL,
                                   0x6e14116da0605617ULL, 0x0c19139cb84c680aULL}; /* (q-3)/4 */
    static const uint64_t e2[4] = {0x9e10460b6c3e7ea3ULL, 0xcbc0b548b438e546ULL,
                                   0xdc2822db40c0ac2eULL, 0x183227397098d014ULL}; /* (q-1)/2 */
    fq2 a1 = fq2_pow(a, e1, 4);
    fq2 alpha = a1; fq2_sqr(&alpha); fq2_mul(&alpha, a);
    fq2 a0 = alpha; fq_neg(&a0.c1);                      /* frobenius_map(1) = conjugation */
    fq2_mul(&a0, &alpha);
    fq2 neg1 = fq2_one(); fq_neg(&neg1.c0);
    if (fq2_eq(&a0, &neg1)) return 0;
    fq2_mul(&a1, a);
    if (fq2_eq(&alpha, &neg1)) {
        fq2 u = {f_zero(), f_one(&FQ)};
        fq2_mul(&a1, &u);
    } else {
        fq2 one = fq2_one();
        fq2_add(&alpha, &one);
        alpha = fq2_pow(&alpha, e2, 4);
        fq2_mul(&a1, &alpha);
    }
    *out = a1;
    return 1;
}
#endif
