/* oracle/p2b_oracle.c -- CPU restatement of the phase2/powersoftau contribution hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library, and only as the checker or the timed CPU
 * baseline.  The product library (phase2_bn254_b200/csrc -> libp2b.so) never links it.
 *
 * PARITY PIN -- "parity unpinned" by reference-produced output bytes: the reference holds no stored output
 * bytes for this path and cannot be run here.  This oracle is pinned by (1) the Montgomery / generator / curve constants hard-coded in
 * pairing/src/bn256/fq.rs, fq2.rs, ec.rs, fr.rs; (2) an independent big-int implementation
 * (oracle/bn254_ref.py) on randomised cases; (3) the reference tests' relations (curve.rs
 * add/double/mul/wnaf/encoding round trips, domain.rs fft∘ifft = id, multiexp == naive sum).
 * The reference itself (Rust, crates.io dependencies) cannot be built here: no rustc/cargo.
 *
 * What is restated (reference file:line):
 *   wnaf_form                          pairing/src/wnaf.rs:18-43
 *   point codecs                       pairing/src/bn256/ec.rs:763-946, 1136-1344, 110-148
 *   batch_exp (phase 1)                powersoftau/src/batched_accumulator.rs:1130-1181
 *   tau powers                         powersoftau/src/batched_accumulator.rs:1201-1216
 *   transform chunk loops              powersoftau/src/batched_accumulator.rs:1187-1289
 *   read_points_chunk error rules      powersoftau/src/batched_accumulator.rs:889-1001
 *   file positions                     powersoftau/src/batched_accumulator.rs:96-178
 *   batch_exp (phase 2)                phase2/src/parameters.rs:424-470
 *   multiexp (Pippenger)               bellman/src/multiexp.rs:53-157, 330-355
 *   serial_fft / parallel_fft / ifft   bellman/src/domain.rs:154-195, 263-376
 * Threading follows the reference: static chunks of len/ncpus on threads (batch_exp, FFT); the
 * Pippenger tasks (window x point-chunk) are handed out dynamically like bellman's CpuPool futures.
 * (4) golden vectors tests/golden/vectors.json generated from the big-int implementation.
 */
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include "field.h"

/* ---- wNAF recoding, wnaf.rs:18-43 ---- */
static int wnaf_form(int64_t *out, const uint64_t k[4], int window) {
    uint64_t c[4] = {k[0], k[1], k[2], k[3]};
    int len = 0;
    while (!r_is_zero(c)) {
        int64_t u;
        if (c[0] & 1) {
            u = (int64_t)(c[0] % (1ULL << (window + 1)));
            if (u > (1LL << window)) u -= 1LL << (window + 1);
            uint64_t t[4] = {0, 0, 0, 0};
            if (u > 0) { t[0] = (uint64_t)u; r_sub(c, t); }
            else { t[0] = (uint64_t)(-u); r_add(c, t); }
        } else u = 0;
        out[len++] = u;
        r_div2(c);
    }
    return len;
}

#define F fq
#define FN(x) fq_##x
#define EC(x) g1_##x
#include "ec_tmpl.h"
#undef F
#undef FN
#undef EC
#define F fq2
#define FN(x) fq2_##x
#define EC(x) g2_##x
#include "ec_tmpl.h"
#undef F
#undef FN
#undef EC

/* curve constants in Montgomery form */
static fq g1_b(void) {   /* fq.rs:11-16 */
    fq b = {{0x7a17caa950ad28d7ULL, 0x1f6ac17ae15521b9ULL, 0x334bea4e696bd284ULL, 0x2a1f6744ce179d8eULL}};
    return b;
}
static fq2 g2_b(void) {  /* fq.rs:18-31 */
    fq2 b = {{{0x3bf938e377b802a8ULL, 0x020b1b273633535dULL, 0x26b7edf049755260ULL, 0x2514c6324384a86dULL}},
             {{0x38e7ecccd1dcff67ULL, 0x65f0b37d93ce0d3eULL, 0xd749d0dd22ac00aaULL, 0x0141b9ce4a688d4dULL}}};
    return b;
}
static g1_aff g1_generator(void) {  /* fq.rs:39-50 */
    g1_aff g = {{{0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL}},
                {{0xa6ba871b8b1e1b3aULL, 0x14f1d651eb8e167bULL, 0xccdd46def0f28c58ULL, 0x1c14ef83340fbe5eULL}}, 0};
    return g;
}
static g2_aff g2_generator(void) {  /* fq.rs:60-83 */
    g2_aff g;
    g.x.c0 = (fq){{0x8e83b5d102bc2026ULL, 0xdceb1935497b0172ULL, 0xfbb8264797811adfULL, 0x19573841af96503bULL}};
    g.x.c1 = (fq){{0xafb4737da84c6140ULL, 0x6043dd5a5802d8c4ULL, 0x09e950fc52a02f86ULL, 0x14fef0833aea7b6bULL}};
    g.y.c0 = (fq){{0x619dfa9d886be9f6ULL, 0xfe7fd297f59e9b78ULL, 0xff9e1a62231b7dfeULL, 0x28fd7eebae9e4206ULL}};
    g.y.c1 = (fq){{0x64095b56c71856eeULL, 0xdc57f922327d3cbbULL, 0x55f935be33351076ULL, 0x0da4a0e693fd6482ULL}};
    g.inf = 0;
    return g;
}

/* ---- error codes shared with include/p2b.h (kept numerically identical) ---- */
enum {
    ORC_OK = 0, ORC_EARG = 1, ORC_EDECODE = 2, ORC_EINFINITY_IN = 3, ORC_EINFINITY_OUT = 4,
    /* decode sub-codes, GroupDecodingError pairing/src/lib.rs:280-291 */
    ORC_D_NOT_ON_CURVE = 1, ORC_D_COORD = 2, ORC_D_UNEXPECTED_INFO = 3, ORC_D_UNEXPECTED_COMPRESSION = 4
};
enum { ENC_UNCOMPRESSED = 0, ENC_COMPRESSED = 1, ENC_RAW_MONT_LE = 2 };
static int enc_bytes(int g2, int enc) { int full = g2 ? 128 : 64; return enc == ENC_COMPRESSED ? full / 2 : full; }

/* ---- codecs ---- */
static int all_zero(const uint8_t *b, int n) { uint8_t a = 0; for (int i = 0; i < n; i++) a |= b[i]; return a == 0; }
static int fq_read(fq *out, const uint8_t *b) { uint64_t r[4]; r_from_be(r, b); return f_from_repr(out, r, &FQ); }
static void fq_write(const fq *a, uint8_t *b) { uint64_t r[4]; f_into_repr(r, a, &FQ); r_to_be(r, b); }

static int g1_on_curve(const g1_aff *p) {   /* ec.rs:133-148 */
    if (p->inf) return 1;
    fq y2 = p->y; fq_sqr(&y2);
    fq x3 = p->x; fq_sqr(&x3); fq_mul(&x3, &p->x);
    fq b = g1_b(); fq_add(&x3, &b);
    return fq_eq(&y2, &x3);
}
static int g2_on_curve(const g2_aff *p) {
    if (p->inf) return 1;
    fq2 y2 = p->y; fq2_sqr(&y2);
    fq2 x3 = p->x; fq2_sqr(&x3); fq2_mul(&x3, &p->x);
    fq2 b = g2_b(); fq2_add(&x3, &b);
    return fq2_eq(&y2, &x3);
}
/* RawEncodable (pairing/src/bn256/ec.rs:653-706): coordinates as Montgomery limbs, little-endian (`into_raw_repr().write_le`),
 * all-zero bytes = the point at infinity; `from_raw_repr` rejects limbs >= q (CoordinateDecodingError). */
static int fq_read_raw(fq *out, const uint8_t *b) {
    for (int i = 0; i < 4; i++) { uint64_t v = 0; for (int j = 7; j >= 0; j--) v = (v << 8) | b[8 * i + j]; out->l[i] = v; }
    return r_cmp(out->l, FQ.m) < 0;
}
static void fq_write_raw(const fq *a, uint8_t *b) {
    for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) b[8 * i + j] = (uint8_t)(a->l[i] >> (8 * j));
}
static int g1_on_curve(const g1_aff *p);
static int g2_on_curve(const g2_aff *p);
static int g1_decode_raw(g1_aff *out, const uint8_t *src, int checked) {
    if (all_zero(src, 64)) { *out = g1_aff_zero(); return 0; }
    out->inf = 0;
    if (!fq_read_raw(&out->x, src) || !fq_read_raw(&out->y, src + 32)) return ORC_D_COORD;
    if (checked && !g1_on_curve(out)) return ORC_D_NOT_ON_CURVE;      /* from_raw_uncompressed_le, ec.rs:696-704 */
    return 0;
}
/* G2 has no RawEncodable impl in the reference; include/p2b.h defines the same layout over x.c0, x.c1, y.c0, y.c1 */
static int g2_decode_raw(g2_aff *out, const uint8_t *src, int checked) {
    if (all_zero(src, 128)) { *out = g2_aff_zero(); return 0; }
    out->inf = 0;
    if (!fq_read_raw(&out->x.c0, src) || !fq_read_raw(&out->x.c1, src + 32) || !fq_read_raw(&out->y.c0, src + 64) ||
        !fq_read_raw(&out->y.c1, src + 96)) return ORC_D_COORD;
    if (checked && !g2_on_curve(out)) return ORC_D_NOT_ON_CURVE;
    return 0;
}
/* returns 0 or a decode sub-code */
static int g1_decode(g1_aff *out, const uint8_t *src, int enc, int checked) {
    if (enc == ENC_RAW_MONT_LE) return g1_decode_raw(out, src, checked);
    int n = enc == ENC_COMPRESSED ? 32 : 64;
    uint8_t b[64]; memcpy(b, src, n);
    if (b[0] & 0x40) {
        b[0] &= 0x3f;
        if (!all_zero(b, n)) return ORC_D_UNEXPECTED_INFO;
        *out = g1_aff_zero();
        return 0;
    }
    if (enc == ENC_UNCOMPRESSED) {
        if (b[0] & 0x80) return ORC_D_UNEXPECTED_INFO;
        b[0] &= 0x3f;
        out->inf = 0;
        if (!fq_read(&out->x, b) || !fq_read(&out->y, b + 32)) return ORC_D_COORD;
        if (checked && !g1_on_curve(out)) return ORC_D_NOT_ON_CURVE;
        return 0;
    }
    int greatest = (b[0] & 0x80) != 0;
    b[0] &= 0x3f;
    fq x;
    if (!fq_read(&x, b)) return ORC_D_COORD;
    fq x3b = x; fq_sqr(&x3b); fq_mul(&x3b, &x);
    fq bb = g1_b(); fq_add(&x3b, &bb);
    fq y;
    if (!fq_sqrt(&y, &x3b)) return ORC_D_NOT_ON_CURVE;
    fq negy = y; fq_neg(&negy);
    out->x = x; out->inf = 0;
    out->y = (fq_gt(&negy, &y) ^ greatest) ? y : negy;   /* (y < negy) ^ greatest, ec.rs:123 */
    return 0;
}
static void g1_encode(const g1_aff *p, uint8_t *dst, int enc) {
    int n = enc == ENC_COMPRESSED ? 32 : 64;
    memset(dst, 0, n);
    if (enc == ENC_RAW_MONT_LE) { if (!p->inf) { fq_write_raw(&p->x, dst); fq_write_raw(&p->y, dst + 32); } return; }
    if (p->inf) { dst[0] |= 0x40; return; }
    fq_write(&p->x, dst);
    if (enc == ENC_UNCOMPRESSED) { fq_write(&p->y, dst + 32); return; }
    fq negy = p->y; fq_neg(&negy);
    if (fq_gt(&p->y, &negy)) dst[0] |= 0x80;
}
static int g2_decode(g2_aff *out, const uint8_t *src, int enc, int checked) {
    if (enc == ENC_RAW_MONT_LE) return g2_decode_raw(out, src, checked);
    int n = enc == ENC_COMPRESSED ? 64 : 128;
    uint8_t b[128]; memcpy(b, src, n);
    if (enc == ENC_UNCOMPRESSED && (b[0] & 0x80)) return ORC_D_UNEXPECTED_COMPRESSION;
    if (b[0] & 0x40) {
        b[0] &= 0x3f;
        if (!all_zero(b, n)) return ORC_D_UNEXPECTED_INFO;
        *out = g2_aff_zero();
        return 0;
    }
    int greatest = (b[0] & 0x80) != 0;
    b[0] &= 0x3f;
    out->inf = 0;
    if (!fq_read(&out->x.c1, b) || !fq_read(&out->x.c0, b + 32)) return ORC_D_COORD;
    if (enc == ENC_UNCOMPRESSED) {
        if (!fq_read(&out->y.c1, b + 64) || !fq_read(&out->y.c0, b + 96)) return ORC_D_COORD;
        if (checked && !g2_on_curve(out)) return ORC_D_NOT_ON_CURVE;
        return 0;
    }
    fq2 x3b = out->x; fq2_sqr(&x3b); fq2_mul(&x3b, &out->x);
    fq2 bb = g2_b(); fq2_add(&x3b, &bb);
    fq2 y;
    if (!fq2_sqrt(&y, &x3b)) return ORC_D_NOT_ON_CURVE;
    fq2 negy = y; fq2_neg(&negy);
    out->y = (fq2_gt(&negy, &y) ^ greatest) ? y : negy;
    return 0;
}
static void g2_encode(const g2_aff *p, uint8_t *dst, int enc) {
    int n = enc == ENC_COMPRESSED ? 64 : 128;
    memset(dst, 0, n);
    if (enc == ENC_RAW_MONT_LE) {
        if (!p->inf) { fq_write_raw(&p->x.c0, dst); fq_write_raw(&p->x.c1, dst + 32); fq_write_raw(&p->y.c0, dst + 64); fq_write_raw(&p->y.c1, dst + 96); }
        return;
    }
    if (p->inf) { dst[0] |= 0x40; return; }
    fq_write(&p->x.c1, dst); fq_write(&p->x.c0, dst + 32);
    if (enc == ENC_UNCOMPRESSED) { fq_write(&p->y.c1, dst + 64); fq_write(&p->y.c0, dst + 96); return; }
    fq2 negy = p->y; fq2_neg(&negy);
    if (fq2_gt(&p->y, &negy)) dst[0] |= 0x80;
}

/* ---- threading helper: static chunks like crossbeam::scope + chunks(len / ncpus) ---- */
typedef void (*range_fn)(void *ctx, size_t lo, size_t hi, int tid);
typedef struct { range_fn fn; void *ctx; size_t lo, hi; int tid; } range_job;
static void *range_tramp(void *p) { range_job *j = (range_job *)p; j->fn(j->ctx, j->lo, j->hi, j->tid); return NULL; }
static void parallel_ranges(size_t n, int threads, range_fn fn, void *ctx) {
    if (threads < 1) threads = 1;
    size_t chunk = n / (size_t)threads;
    if (chunk == 0) chunk = 1;          /* reference panics on chunks_mut(0); see DESIGN.md */
    size_t njobs = (n + chunk - 1) / chunk;
    if (njobs <= 1) { if (n) fn(ctx, 0, n, 0); return; }
    pthread_t *th = (pthread_t *)malloc(njobs * sizeof(pthread_t));
    range_job *jobs = (range_job *)malloc(njobs * sizeof(range_job));
    for (size_t j = 0; j < njobs; j++) {
        jobs[j] = (range_job){fn, ctx, j * chunk, (j + 1) * chunk < n ? (j + 1) * chunk : n, (int)j};
        pthread_create(&th[j], NULL, range_tramp, &jobs[j]);
    }
    for (size_t j = 0; j < njobs; j++) pthread_join(th[j], NULL);
    free(th); free(jobs);
}

/* ---- batch_exp over wire bytes ----
 * scalars: n_scalars == n (one per point) or 1 (broadcast; phase-2 shape), 32-byte BE canonical.
 * coeff (optional, 32-byte BE): exp[i] *= coeff first (batched_accumulator.rs:1152-1155).
 * Each thread: wNAF(4) per point, then batch_normalization of its chunk, then encode. */
typedef struct {
    int g2, in_enc, out_enc, checked, reject_inf_in, reject_inf_out;
    const uint8_t *in; uint8_t *out; size_t n;
    const fe *exps; size_t n_exps;     /* Montgomery Fr */
    const fe *coeff;
    int err, sub; size_t err_index;
    pthread_mutex_t mu;
} bexp_ctx;

static void bexp_set_err(bexp_ctx *c, int err, int sub, size_t idx) {
    pthread_mutex_lock(&c->mu);
    if (!c->err || idx < c->err_index) { c->err = err; c->sub = sub; c->err_index = idx; }
    pthread_mutex_unlock(&c->mu);
}

static void bexp_range(void *vp, size_t lo, size_t hi, int tid) {
    (void)tid;
    bexp_ctx *c = (bexp_ctx *)vp;
    size_t m = hi - lo;
    int isz = enc_bytes(c->g2, c->in_enc), osz = enc_bytes(c->g2, c->out_enc);
    if (!c->g2) {
        g1_jac *proj = (g1_jac *)malloc(m * sizeof(g1_jac));
        fq *scratch = (fq *)malloc(m * sizeof(fq));
        for (size_t i = 0; i < m; i++) {
            g1_aff p;
            int sub = g1_decode(&p, c->in + (lo + i) * isz, c->in_enc, c->checked);
            if (sub) { bexp_set_err(c, ORC_EDECODE, sub, lo + i); p = g1_aff_zero(); }
            else if (p.inf && c->reject_inf_in) bexp_set_err(c, ORC_EINFINITY_IN, 0, lo + i);
            fe e = c->exps[c->n_exps == 1 ? 0 : lo + i];
            if (c->coeff) f_mul(&e, c->coeff, &FR);
            uint64_t k[4]; f_into_repr(k, &e, &FR);
            proj[i] = g1_wnaf_mul(&p, k);
        }
        g1_batch_normalize(proj, m, scratch);
        for (size_t i = 0; i < m; i++) {
            g1_aff a = g1_to_aff(&proj[i]);
            if (a.inf && c->reject_inf_out) bexp_set_err(c, ORC_EINFINITY_OUT, 0, lo + i);
            g1_encode(&a, c->out + (lo + i) * osz, c->out_enc);
        }
        free(proj); free(scratch);
    } else {
        g2_jac *proj = (g2_jac *)malloc(m * sizeof(g2_jac));
        fq2 *scratch = (fq2 *)malloc(m * sizeof(fq2));
        for (size_t i = 0; i < m; i++) {
            g2_aff p;
            int sub = g2_decode(&p, c->in + (lo + i) * isz, c->in_enc, c->checked);
            if (sub) { bexp_set_err(c, ORC_EDECODE, sub, lo + i); p = g2_aff_zero(); }
            else if (p.inf && c->reject_inf_in) bexp_set_err(c, ORC_EINFINITY_IN, 0, lo + i);
            fe e = c->exps[c->n_exps == 1 ? 0 : lo + i];
            if (c->coeff) f_mul(&e, c->coeff, &FR);
            uint64_t k[4]; f_into_repr(k, &e, &FR);
            proj[i] = g2_wnaf_mul(&p, k);
        }
        g2_batch_normalize(proj, m, scratch);
        for (size_t i = 0; i < m; i++) {
            g2_aff a = g2_to_aff(&proj[i]);
            if (a.inf && c->reject_inf_out) bexp_set_err(c, ORC_EINFINITY_OUT, 0, lo + i);
            g2_encode(&a, c->out + (lo + i) * osz, c->out_enc);
        }
        free(proj); free(scratch);
    }
}

static int fr_read(fe *out, const uint8_t *be) { uint64_t r[4]; r_from_be(r, be); return f_from_repr(out, r, &FR); }

static int batch_exp_core(int g2, const uint8_t *in, uint8_t *out, size_t n, const fe *exps, size_t n_exps,
                          const fe *coeff, int in_enc, int out_enc, int checked, int reject_in, int reject_out,
                          int threads, uint64_t *err_index, int *err_sub) {
    bexp_ctx c = {g2, in_enc, out_enc, checked, reject_in, reject_out, in, out, n, exps, n_exps, coeff, 0, 0, 0,
                  PTHREAD_MUTEX_INITIALIZER};
    parallel_ranges(n, threads, bexp_range, &c);
    if (err_index) *err_index = c.err_index;
    if (err_sub) *err_sub = c.sub;
    return c.err;
}

/* exported: scalars as 32-byte BE canonical values */
int orc_batch_mul(int g2, const uint8_t *in, uint8_t *out, size_t n, const uint8_t *scalars_be, size_t n_scalars,
                  int in_enc, int out_enc, int checked, int reject_inf, int threads,
                  uint64_t *err_index, int *err_sub) {
    if (n_scalars != 1 && n_scalars != n) return ORC_EARG;
    fe *exps = (fe *)malloc((n_scalars ? n_scalars : 1) * sizeof(fe));
    for (size_t i = 0; i < n_scalars; i++)
        if (!fr_read(&exps[i], scalars_be + 32 * i)) { free(exps); return ORC_EARG; }
    int rc = batch_exp_core(g2, in, out, n, exps, n_scalars, NULL, in_enc, out_enc, checked, reject_inf, reject_inf,
                            threads, err_index, err_sub);
    free(exps);
    return rc;
}

/* tau powers for [start, start+n): per thread one pow() then a running product (1201-1216) */
typedef struct { fe *out; fe tau; uint64_t start; } taup_ctx;
static void taup_range(void *vp, size_t lo, size_t hi, int tid) {
    (void)tid;
    taup_ctx *c = (taup_ctx *)vp;
    uint64_t e = c->start + lo;
    fe acc = f_pow(&c->tau, &e, 1, &FR);
    for (size_t i = lo; i < hi; i++) { c->out[i] = acc; f_mul(&acc, &c->tau, &FR); }
}

int orc_batch_mul_powers(int g2, const uint8_t *in, uint8_t *out, size_t n, const uint8_t tau_be[32],
                         const uint8_t *coeff_be, uint64_t start, int in_enc, int out_enc, int checked,
                         int threads, uint64_t *err_index, int *err_sub) {
    fe tau, coeff;
    if (!fr_read(&tau, tau_be)) return ORC_EARG;
    if (coeff_be && !fr_read(&coeff, coeff_be)) return ORC_EARG;
    fe *exps = (fe *)malloc((n ? n : 1) * sizeof(fe));
    taup_ctx tc = {exps, tau, start};
    parallel_ranges(n, threads, taup_range, &tc);
    int rc = batch_exp_core(g2, in, out, n, exps, n, coeff_be ? &coeff : NULL, in_enc, out_enc, checked, 1, 1,
                            threads, err_index, err_sub);
    free(exps);
    return rc;
}

/* ---- ceremony geometry, powersoftau/src/parameters.rs:72-120 ---- */
typedef struct { uint64_t powers, powers_g1, g1i, g2i, g1o, g2o; } geom;
enum { EL_TAU_G1, EL_TAU_G2, EL_ALPHA_G1, EL_BETA_G1, EL_BETA_G2 };
static uint64_t geom_pos(const geom *g, uint64_t index, int el, int out) {
    uint64_t s1 = out ? g->g1o : g->g1i, s2 = out ? g->g2o : g->g2i, p = 0;
    switch (el) {
        case EL_TAU_G1: p = s1 * index; break;
        case EL_TAU_G2: p = s1 * g->powers_g1 + s2 * index; break;
        case EL_ALPHA_G1: p = s1 * g->powers_g1 + s2 * g->powers + s1 * index; break;
        case EL_BETA_G1: p = s1 * g->powers_g1 + s2 * g->powers + s1 * g->powers + s1 * index; break;
        case EL_BETA_G2: p = s1 * g->powers_g1 + s2 * g->powers + 2 * s1 * g->powers; break;
    }
    return p + 64;
}
uint64_t orc_accumulator_size(uint32_t size_log2, int compressed) {
    uint64_t p = 1ULL << size_log2, pg1 = 2 * p - 1, s1 = compressed ? 32 : 64, s2 = compressed ? 64 : 128;
    return pg1 * s1 + p * s2 + 2 * p * s1 + s2 + 64;
}

/* initial accumulator: batched_accumulator.rs:1295-1347 (hash prefix is the caller's) */
int orc_pot_generate_initial(uint8_t *out, uint64_t out_len, uint32_t size_log2, int compressed) {
    if (out_len < orc_accumulator_size(size_log2, compressed)) return ORC_EARG;
    geom g = {1ULL << size_log2, (2ULL << size_log2) - 1, 0, 0, compressed ? 32u : 64u, compressed ? 64u : 128u};
    g1_aff a = g1_generator(); g2_aff b = g2_generator();
    uint8_t e1[64], e2[128];
    g1_encode(&a, e1, compressed); g2_encode(&b, e2, compressed);
    for (uint64_t i = 0; i < g.powers_g1; i++) memcpy(out + geom_pos(&g, i, EL_TAU_G1, 1), e1, g.g1o);
    for (uint64_t i = 0; i < g.powers; i++) {
        memcpy(out + geom_pos(&g, i, EL_TAU_G2, 1), e2, g.g2o);
        memcpy(out + geom_pos(&g, i, EL_ALPHA_G1, 1), e1, g.g1o);
        memcpy(out + geom_pos(&g, i, EL_BETA_G1, 1), e1, g.g1o);
    }
    memcpy(out + geom_pos(&g, 0, EL_BETA_G2, 1), e2, g.g2o);
    return ORC_OK;
}

/* transform: writes the accumulator region [64, accumulator_size(out)) of `response`;
 * bytes [0,64) and the public key tail belong to the caller (compute_constrained.rs:155-161,207-209) */
int orc_pot_transform(const uint8_t *challenge, uint64_t challenge_len, uint8_t *response, uint64_t response_len,
                      uint32_t size_log2, uint32_t batch_size, int in_compressed, int out_compressed, int check_input,
                      const uint8_t tau_be[32], const uint8_t alpha_be[32], const uint8_t beta_be[32], int threads,
                      uint64_t *err_index, int *err_sub) {
    if (challenge_len < orc_accumulator_size(size_log2, in_compressed) ||
        response_len < orc_accumulator_size(size_log2, out_compressed) || batch_size == 0)
        return ORC_EARG;
    geom g = {1ULL << size_log2, (2ULL << size_log2) - 1, in_compressed ? 32u : 64u, in_compressed ? 64u : 128u,
              out_compressed ? 32u : 64u, out_compressed ? 64u : 128u};
    fe beta;
    if (!fr_read(&beta, beta_be)) return ORC_EARG;
    int rc;
    /* two loops as in the reference: [0, powers) with all five element types, then
     * [powers, powers_g1) with tau_g1 only; each chunked by batch_size from its own start */
    for (int phase = 0; phase < 2; phase++) {
        uint64_t lo = phase ? g.powers : 0, hi = phase ? g.powers_g1 : g.powers;
        for (uint64_t start = lo; start < hi; start += batch_size) {
            uint64_t n = start + batch_size <= hi ? batch_size : hi - start;
            rc = orc_batch_mul_powers(0, challenge + geom_pos(&g, start, EL_TAU_G1, 0), response + geom_pos(&g, start, EL_TAU_G1, 1),
                                      n, tau_be, NULL, start, in_compressed, out_compressed, check_input, threads, err_index, err_sub);
            if (rc) return rc;
            if (phase) continue;
            rc = orc_batch_mul_powers(1, challenge + geom_pos(&g, start, EL_TAU_G2, 0), response + geom_pos(&g, start, EL_TAU_G2, 1),
                                      n, tau_be, NULL, start, in_compressed, out_compressed, check_input, threads, err_index, err_sub);
            if (rc) return rc;
            rc = orc_batch_mul_powers(0, challenge + geom_pos(&g, start, EL_ALPHA_G1, 0), response + geom_pos(&g, start, EL_ALPHA_G1, 1),
                                      n, tau_be, alpha_be, start, in_compressed, out_compressed, check_input, threads, err_index, err_sub);
            if (rc) return rc;
            rc = orc_batch_mul_powers(0, challenge + geom_pos(&g, start, EL_BETA_G1, 0), response + geom_pos(&g, start, EL_BETA_G1, 1),
                                      n, tau_be, beta_be, start, in_compressed, out_compressed, check_input, threads, err_index, err_sub);
            if (rc) return rc;
            /* beta_g2 = beta_g2.mul(beta), recomputed from the input every chunk (1230-1234) */
            g2_aff b2;
            int sub = g2_decode(&b2, challenge + geom_pos(&g, 0, EL_BETA_G2, 0), in_compressed, check_input);
            if (sub) { if (err_sub) *err_sub = sub; if (err_index) *err_index = 0; return ORC_EDECODE; }
            if (b2.inf) return ORC_EINFINITY_IN;
            uint64_t k[4]; f_into_repr(k, &beta, &FR);
            g2_jac r = g2_mul_bits(&b2, k);
            g2_aff ra = g2_to_aff(&r);
            if (ra.inf) return ORC_EINFINITY_OUT;
            g2_encode(&ra, response + geom_pos(&g, 0, EL_BETA_G2, 1), out_compressed);
        }
    }
    return ORC_OK;
}

/* ---- Pippenger, bellman/src/multiexp.rs:53-157,330-355: one task per window ---- */
typedef struct {
    int g2; const uint8_t *points; int enc; const uint64_t (*ks)[4]; size_t n; unsigned c; unsigned nwin;
    void *accs; const void *affs;
    size_t nchunks, next_task; pthread_mutex_t mu;
} msm_ctx;
/* Tasks are (window, point-chunk) pairs handed out dynamically, as bellman's CpuPool does for its per-window
 * futures (multiexp.rs:75-145); with more workers than windows the points are also split into chunks whose
 * per-window partial sums are merged under a mutex, as dense_multiexp_inner does (multiexp.rs:393-446). */
static void msm_window(void *vp, size_t lo_unused, size_t hi_unused, int tid) {
    (void)tid; (void)lo_unused; (void)hi_unused;
    msm_ctx *m = (msm_ctx *)vp;
    for (;;) {
        size_t task = __atomic_fetch_add(&m->next_task, 1, __ATOMIC_RELAXED);
        if (task >= (size_t)m->nwin * m->nchunks) break;
        size_t w = task / m->nchunks, ch = task % m->nchunks;
        size_t per = (m->n + m->nchunks - 1) / m->nchunks, p_lo = ch * per, p_hi = p_lo + per < m->n ? p_lo + per : m->n;
        unsigned skip = (unsigned)w * m->c;
        size_t nb = ((size_t)1 << m->c) - 1;
        static const uint64_t one[4] = {1, 0, 0, 0};
        if (!m->g2) {
            const g1_aff *aff = (const g1_aff *)m->affs;
            g1_jac acc = g1_jac_zero();
            g1_jac *b = (g1_jac *)malloc(nb * sizeof(g1_jac));
            for (size_t i = 0; i < nb; i++) b[i] = g1_jac_zero();
            for (size_t i = p_lo; i < p_hi; i++) {
                if (r_is_zero(m->ks[i])) continue;
                if (r_cmp(m->ks[i], one) == 0) { if (w == 0) g1_madd(&acc, &aff[i]); continue; }
                uint64_t e[4] = {m->ks[i][0], m->ks[i][1], m->ks[i][2], m->ks[i][3]};
                r_shr(e, skip);
                uint64_t d = e[0] % (1ULL << m->c);
                if (d) g1_madd(&b[d - 1], &aff[i]);
            }
            g1_jac run = g1_jac_zero();
            for (size_t i = nb; i-- > 0;) { g1_add(&run, &b[i]); g1_add(&acc, &run); }
            pthread_mutex_lock(&m->mu);
            g1_add(&((g1_jac *)m->accs)[w], &acc);
            pthread_mutex_unlock(&m->mu);
            free(b);
        } else {
            const g2_aff *aff = (const g2_aff *)m->affs;
            g2_jac acc = g2_jac_zero();
            g2_jac *b = (g2_jac *)malloc(nb * sizeof(g2_jac));
            for (size_t i = 0; i < nb; i++) b[i] = g2_jac_zero();
            for (size_t i = p_lo; i < p_hi; i++) {
                if (r_is_zero(m->ks[i])) continue;
                if (r_cmp(m->ks[i], one) == 0) { if (w == 0) g2_madd(&acc, &aff[i]); continue; }
                uint64_t e[4] = {m->ks[i][0], m->ks[i][1], m->ks[i][2], m->ks[i][3]};
                r_shr(e, skip);
                uint64_t d = e[0] % (1ULL << m->c);
                if (d) g2_madd(&b[d - 1], &aff[i]);
            }
            g2_jac run = g2_jac_zero();
            for (size_t i = nb; i-- > 0;) { g2_add(&run, &b[i]); g2_add(&acc, &run); }
            pthread_mutex_lock(&m->mu);
            g2_add(&((g2_jac *)m->accs)[w], &acc);
            pthread_mutex_unlock(&m->mu);
            free(b);
        }
    }
}
typedef struct { int g2; const uint8_t *points; void *affs; int bad; } dec_ctx;
static void dec_range(void *vp, size_t lo, size_t hi, int tid) {
    (void)tid;
    dec_ctx *d = (dec_ctx *)vp;
    for (size_t i = lo; i < hi; i++) {
        int sub = d->g2 ? g2_decode(&((g2_aff *)d->affs)[i], d->points + 128 * i, 0, 0)
                        : g1_decode(&((g1_aff *)d->affs)[i], d->points + 64 * i, 0, 0);
        if (sub) d->bad = 1;
    }
}
/* points: uncompressed wire; scalars: 32-byte BE canonical (< r); out: uncompressed wire */
typedef struct { const uint8_t *be; uint64_t (*ks)[4]; int bad; } ksp_ctx;
static void ksp_range(void *vp, size_t lo, size_t hi, int tid) {
    (void)tid;
    ksp_ctx *k = (ksp_ctx *)vp;
    for (size_t i = lo; i < hi; i++) {
        r_from_be(k->ks[i], k->be + 32 * i);
        if (r_cmp(k->ks[i], FR.m) >= 0) k->bad = 1;
    }
}
static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec; }
/* seconds[0] = wire decode of the inputs (not part of the reference's multiexp, which takes G1Affine / FrRepr values),
 * seconds[1] = the Pippenger proper (bucket phase on all workers + Horner over the windows + into_affine) */
int orc_msm_timed(int g2, const uint8_t *points, const uint8_t *scalars_be, size_t n, uint8_t *out, int threads, double *seconds);
int orc_msm(int g2, const uint8_t *points, const uint8_t *scalars_be, size_t n, uint8_t *out, int threads) {
    return orc_msm_timed(g2, points, scalars_be, n, out, threads, NULL);
}
int orc_msm_timed(int g2, const uint8_t *points, const uint8_t *scalars_be, size_t n, uint8_t *out, int threads, double *seconds) {
    unsigned c = n < 32 ? 3u : (unsigned)ceil(log((double)n));
    unsigned nwin = (254 + c - 1) / c;   /* regions while skip < Fr::NUM_BITS */
    const double t0 = now_s();
    uint64_t(*ks)[4] = (uint64_t(*)[4])malloc((n ? n : 1) * 32);
    ksp_ctx kc = {scalars_be, ks, 0};
    parallel_ranges(n, threads, ksp_range, &kc);
    if (kc.bad) { free(ks); return ORC_EARG; }
    void *affs = malloc((n ? n : 1) * (g2 ? sizeof(g2_aff) : sizeof(g1_aff)));
    dec_ctx dc = {g2, points, affs, 0};
    parallel_ranges(n, threads, dec_range, &dc);
    if (dc.bad) { free(ks); free(affs); return ORC_EDECODE; }
    const double t1 = now_s();
    void *accs = malloc(nwin * (g2 ? sizeof(g2_jac) : sizeof(g1_jac)));
    for (unsigned w = 0; w < nwin; w++) {
        if (g2) ((g2_jac *)accs)[w] = g2_jac_zero(); else ((g1_jac *)accs)[w] = g1_jac_zero();
    }
    if (threads < 1) threads = 1;
    size_t nchunks = threads > (int)nwin && n >= 4096 ? ((size_t)threads + nwin - 1) / nwin : 1;
    msm_ctx m = {g2, points, 0, (const uint64_t(*)[4])ks, n, c, nwin, accs, affs, nchunks, 0, PTHREAD_MUTEX_INITIALIZER};
    parallel_ranges((size_t)threads, threads, msm_window, &m);     /* `threads` workers pulling tasks */
    if (!g2) {
        g1_jac *a = (g1_jac *)accs;
        g1_jac hi = a[nwin - 1];
        for (int w = (int)nwin - 2; w >= 0; w--) { for (unsigned k = 0; k < c; k++) g1_dbl(&hi); g1_add(&hi, &a[w]); }
        g1_aff r = g1_to_aff(&hi); g1_encode(&r, out, 0);
    } else {
        g2_jac *a = (g2_jac *)accs;
        g2_jac hi = a[nwin - 1];
        for (int w = (int)nwin - 2; w >= 0; w--) { for (unsigned k = 0; k < c; k++) g2_dbl(&hi); g2_add(&hi, &a[w]); }
        g2_aff r = g2_to_aff(&hi); g2_encode(&r, out, 0);
    }
    free(ks); free(affs); free(accs);
    if (seconds) { seconds[0] = t1 - t0; seconds[1] = now_s() - t1; }
    return ORC_OK;
}

/* sum of encoded points (used to check the multi-GPU partial combination) */
int orc_sum_points(int g2, const uint8_t *points, size_t n, uint8_t *out) {
    if (!g2) {
        g1_jac acc = g1_jac_zero();
        for (size_t i = 0; i < n; i++) { g1_aff p; if (g1_decode(&p, points + 64 * i, 0, 1)) return ORC_EDECODE; g1_madd(&acc, &p); }
        g1_aff r = g1_to_aff(&acc); g1_encode(&r, out, 0);
    } else {
        g2_jac acc = g2_jac_zero();
        for (size_t i = 0; i < n; i++) { g2_aff p; if (g2_decode(&p, points + 128 * i, 0, 1)) return ORC_EDECODE; g2_madd(&acc, &p); }
        g2_aff r = g2_to_aff(&acc); g2_encode(&r, out, 0);
    }
    return ORC_OK;
}

/* ---- Fr FFT, bellman/src/domain.rs ---- */
static unsigned bitrev(unsigned n, unsigned l) { unsigned r = 0; for (unsigned i = 0; i < l; i++) { r = (r << 1) | (n & 1); n >>= 1; } return r; }
static fe fr_pow_u64(const fe *a, uint64_t e) { return f_pow(a, &e, 1, &FR); }
static void serial_fft(fe *a, const fe *omega, unsigned log_n) {   /* domain.rs:274-317 */
    uint32_t n = 1u << log_n;
    for (uint32_t k = 0; k < n; k++) { uint32_t rk = bitrev(k, log_n); if (k < rk) { fe t = a[rk]; a[rk] = a[k]; a[k] = t; } }
    uint32_t m = 1;
    for (unsigned s = 0; s < log_n; s++) {
        fe w_m = fr_pow_u64(omega, n / (2 * m));
        for (uint32_t k = 0; k < n; k += 2 * m) {
            fe w = f_one(&FR);
            for (uint32_t j = 0; j < m; j++) {
                fe t = a[k + j + m]; f_mul(&t, &w, &FR);
                fe tmp = a[k + j]; f_sub(&tmp, &t, &FR);
                a[k + j + m] = tmp;
                f_add(&a[k + j], &t, &FR);
                f_mul(&w, &w_m, &FR);
            }
        }
        m *= 2;
    }
}
typedef struct { const fe *a; fe **tmp; const fe *omega; fe new_omega; unsigned log_n, log_cpus, log_new_n; } pfft_ctx;
static void pfft_sub(void *vp, size_t lo, size_t hi, int tid) {     /* domain.rs:329-359 */
    (void)tid;
    pfft_ctx *c = (pfft_ctx *)vp;
    for (size_t j = lo; j < hi; j++) {
        fe *tmp = c->tmp[j];
        fe omega_j = fr_pow_u64(c->omega, j);
        fe omega_step = fr_pow_u64(c->omega, (uint64_t)j << c->log_new_n);
        fe elt = f_one(&FR);
        size_t num_cpus = (size_t)1 << c->log_cpus;
        for (size_t i = 0; i < ((size_t)1 << c->log_new_n); i++) {
            for (size_t s = 0; s < num_cpus; s++) {
                size_t idx = (i + (s << c->log_new_n)) % ((size_t)1 << c->log_n);
                fe t = c->a[idx]; f_mul(&t, &elt, &FR);
                f_add(&tmp[i], &t, &FR);
                f_mul(&elt, &omega_step, &FR);
            }
            f_mul(&elt, &omega_j, &FR);
        }
        serial_fft(tmp, &c->new_omega, c->log_new_n);
    }
}
static void best_fft(fe *a, const fe *omega, unsigned log_n, int threads) {
    unsigned log_cpus = 0;
    while ((2u << log_cpus) <= (unsigned)threads) log_cpus++;      /* log2_floor(num_cpus), multicore.rs:33-39 */
    if (log_n <= log_cpus) { serial_fft(a, omega, log_n); return; }
    size_t num_cpus = (size_t)1 << log_cpus, sub = (size_t)1 << (log_n - log_cpus);
    fe **tmp = (fe **)malloc(num_cpus * sizeof(fe *));
    for (size_t j = 0; j < num_cpus; j++) tmp[j] = (fe *)calloc(sub, sizeof(fe));
    pfft_ctx c = {a, tmp, omega, fr_pow_u64(omega, num_cpus), log_n, log_cpus, log_n - log_cpus};
    parallel_ranges(num_cpus, (int)num_cpus, pfft_sub, &c);
    size_t mask = num_cpus - 1;
    for (size_t idx = 0; idx < ((size_t)1 << log_n); idx++) a[idx] = tmp[idx & mask][idx >> log_cpus];
    for (size_t j = 0; j < num_cpus; j++) free(tmp[j]);
    free(tmp);
}
typedef struct { fe *a; fe g; fe scale; int use_g; } scale_ctx;
static void scale_range(void *vp, size_t lo, size_t hi, int tid) {   /* distribute_powers / minv, domain.rs:159-189 */
    (void)tid;
    scale_ctx *c = (scale_ctx *)vp;
    if (c->use_g) {
        fe u = fr_pow_u64(&c->g, lo);
        for (size_t i = lo; i < hi; i++) { f_mul(&c->a[i], &u, &FR); f_mul(&u, &c->g, &FR); }
    } else
        for (size_t i = lo; i < hi; i++) f_mul(&c->a[i], &c->scale, &FR);
}
/* data: 2^log_n scalars, 32-byte BE canonical, transformed in place */
int orc_fr_fft(uint8_t *data, uint32_t log_n, int inverse, int coset, int threads) {
    if (log_n > 28) return ORC_EARG;
    size_t n = (size_t)1 << log_n;
    fe *a = (fe *)malloc(n * sizeof(fe));
    for (size_t i = 0; i < n; i++) if (!fr_read(&a[i], data + 32 * i)) { free(a); return ORC_EARG; }
    /* root_of_unity = 7^((r-1)/2^28), then squared down (domain.rs:80-89) */
    static const uint64_t t[4] = {0x9b9709143e1f593fULL, 0x181585d2833e8487ULL, 0x131a029b85045b68ULL, 0x000000030644e72eULL};
    fe gen; uint64_t seven[4] = {7, 0, 0, 0}; f_from_repr(&gen, seven, &FR);
    fe omega = f_pow(&gen, t, 4, &FR);
    for (unsigned i = log_n; i < 28; i++) f_sqr(&omega, &FR);
    fe geninv; f_inv(&geninv, &gen, &FR);
    if (coset && !inverse) { scale_ctx s = {a, gen, f_zero(), 1}; parallel_ranges(n, threads, scale_range, &s); }
    if (inverse) { fe oi; f_inv(&oi, &omega, &FR); omega = oi; }
    best_fft(a, &omega, log_n, threads);
    if (inverse) {
        uint64_t mrepr[4] = {n, 0, 0, 0};
        fe m, minv; f_from_repr(&m, mrepr, &FR); f_inv(&minv, &m, &FR);
        scale_ctx s = {a, f_zero(), minv, 0}; parallel_ranges(n, threads, scale_range, &s);
        if (coset) { scale_ctx s2 = {a, geninv, f_zero(), 1}; parallel_ranges(n, threads, scale_range, &s2); }
    }
    for (size_t i = 0; i < n; i++) { uint64_t r[4]; f_into_repr(r, &a[i], &FR); r_to_be(r, data + 32 * i); }
    free(a);
    return ORC_OK;
}

/* ---- small helpers for tests ---- */
/* out = [k] * P for one point via the affine mul_bits path (ec.rs:96-103,179-182) */
int orc_point_mul(int g2, const uint8_t *in, const uint8_t k_be[32], uint8_t *out, int in_enc, int out_enc) {
    uint64_t k[4]; r_from_be(k, k_be);
    if (!g2) {
        g1_aff p; int sub = g1_decode(&p, in, in_enc, 1); if (sub) return ORC_EDECODE;
        g1_jac r = g1_mul_bits(&p, k); g1_aff a = g1_to_aff(&r); g1_encode(&a, out, out_enc);
    } else {
        g2_aff p; int sub = g2_decode(&p, in, in_enc, 1); if (sub) return ORC_EDECODE;
        g2_jac r = g2_mul_bits(&p, k); g2_aff a = g2_to_aff(&r); g2_encode(&a, out, out_enc);
    }
    return ORC_OK;
}
/* re-encode (decompress / compress / check); returns decode sub-code in *err_sub */
int orc_point_recode(int g2, const uint8_t *in, uint8_t *out, int in_enc, int out_enc, int checked, int *err_sub) {
    int sub;
    if (!g2) { g1_aff p; sub = g1_decode(&p, in, in_enc, checked); if (!sub) g1_encode(&p, out, out_enc); }
    else { g2_aff p; sub = g2_decode(&p, in, in_enc, checked); if (!sub) g2_encode(&p, out, out_enc); }
    if (err_sub) *err_sub = sub;
    return sub ? ORC_EDECODE : ORC_OK;
}
/* field KAT access: op 0 mul, 1 add, 2 sub, 3 inv(a), 4 sqrt(a) (Fq only); field 0 = Fq, 1 = Fr.
 * Operands / result are 32-byte BE canonical values. returns 0 if no result (inv 0 / non-residue). */
int orc_field_op(int field, int op, const uint8_t *a_be, const uint8_t *b_be, uint8_t *out_be) {
    const fparams *p = field ? &FR : &FQ;
    uint64_t ra[4], rb[4];
    fe a, b;
    r_from_be(ra, a_be); r_from_be(rb, b_be);
    if (!f_from_repr(&a, ra, p) || !f_from_repr(&b, rb, p)) return 0;
    int ok = 1;
    switch (op) {
        case 0: f_mul(&a, &b, p); break;
        case 1: f_add(&a, &b, p); break;
        case 2: f_sub(&a, &b, p); break;
        case 3: { fe o; ok = f_inv(&o, &a, p); if (ok) a = o; break; }
        case 4: { fe o; ok = field == 0 && fq_sqrt(&o, &a); if (ok) a = o; break; }
        default: return 0;
    }
    if (!ok) return 0;
    uint64_t r[4]; f_into_repr(r, &a, p); r_to_be(r, out_be);
    return 1;
}
/* Fq2 KAT access: out = (c0 + c1 u)^e and the product a * b, operands as 32-byte BE canonical coordinates, results as RAW
 * MONTGOMERY LIMBS (8 x u64: c0 then c1) so that they compare directly with the reference's hard-coded tables
 * (pairing/src/bn256/fq.rs:87-432: XI_TO_Q_MINUS_1_OVER_2, FROBENIUS_COEFF_FQ6_C1 / C2, FROBENIUS_COEFF_FQ12_C1 are powers
 * of 9 + u computed with the Fq2 arithmetic under test). */
int orc_fq2_pow_raw(const uint8_t *c0_be, const uint8_t *c1_be, const uint8_t *e_be, uint64_t *out_limbs) {
    fq2 a;
    if (!fq_read(&a.c0, c0_be) || !fq_read(&a.c1, c1_be)) return 0;
    uint64_t e[4]; r_from_be(e, e_be);
    fq2 r = fq2_pow(&a, e, 4);
    memcpy(out_limbs, r.c0.l, 32); memcpy(out_limbs + 4, r.c1.l, 32);
    return 1;
}
int orc_fq2_mul_raw(const uint64_t *a_limbs, const uint64_t *b_limbs, uint64_t *out_limbs) {
    fq2 a, b;
    memcpy(a.c0.l, a_limbs, 32); memcpy(a.c1.l, a_limbs + 4, 32); memcpy(b.c0.l, b_limbs, 32); memcpy(b.c1.l, b_limbs + 4, 32);
    fq2_mul(&a, &b);
    memcpy(out_limbs, a.c0.l, 32); memcpy(out_limbs + 4, a.c1.l, 32);
    return 1;
}
/* raw Montgomery limbs of constants, for pinning against the reference's hard-coded tables */
void orc_constants(uint64_t *out /* 4 x {R mod q, R^2 mod q, R mod r, R^2 mod r} */) {
    memcpy(out, FQ.one, 32); memcpy(out + 4, FQ.r2, 32); memcpy(out + 8, FR.one, 32); memcpy(out + 12, FR.r2, 32);
}
