// tests/host/host_sim.cpp -- compiles the product's per-point device logic (fp.cuh / ec.cuh / smul.cuh /
// codec.cuh are host+device) for the CPU so that the GLV split, the signed-window ladder, the common-Z table
// and the codecs can be checked against the oracle on a machine without a GPU.  Test-only.
#include <cstring>
#include "../../phase2_bn254_b200/csrc/codec.cuh"
#include "../../phase2_bn254_b200/csrc/smul.cuh"
#include "../../phase2_bn254_b200/csrc/xyzz.cuh"
using namespace p2b;

static void be_to_words(const uint8_t *b, uint32_t *w, int nwords) { memcpy(w, b, nwords * 4); }
static void words_to_be(const uint32_t *w, uint8_t *b, int nwords) { memcpy(b, w, nwords * 4); }
static void k_from_be(const uint8_t *be, uint32_t k[8]) {
    uint32_t w[8]; memcpy(w, be, 32);
    for (int i = 0; i < 8; i++) k[i] = bswap32(w[7 - i]);
}
template <class F> static void to_affine(const Jac<F> &j, Aff<F> &a, bool &inf) {
    inf = is_zero(j.z);
    if (inf) return;
    F zi = inv(j.z);
    F zi2 = sqr(zi);
    a.x = mul(j.x, zi2);
    a.y = mul(j.y, mul(zi2, zi));
}

extern "C" {
// path: 0 = production fast path (G1: GLV; G2: window4), 1 = window4 generic, 2 = binary
int sim_point_mul(int g2, const uint8_t *in, const uint8_t *k_be, uint8_t *out, int in_enc, int out_enc, int path, int *bad_out) {
    uint32_t k[8]; k_from_be(k_be, k);
    bool bad = false;
    if (!g2) {
        uint32_t w[16]; be_to_words(in, w, wire_words<Fq>(in_enc));
        Aff<Fq> p; bool inf;
        int rc = point_decode<Fq>(p, inf, w, in_enc, true);
        if (rc) return rc;
        uint32_t tblmem[8 * 16];
        StridedTable<Fq> tbl{tblmem, 1};
        Fq zr[8];
        Jac<Fq> r;
        if (inf) r = jac_infinity<Fq>();
        else if (path == 0) r = g1_mul_glv(p, k, tbl, zr, bad);
        else if (path == 1) r = mul_window4<Fq>(p, k, tbl, zr, bad);
        else if (path == 4) { UniformDigits u = uniform_digits(k); r = mul_glv_uniform<Fq>(p, u, tbl, zr, bad); }
        else r = mul_binary<Fq>(p, k);
        Aff<Fq> a; bool oinf; to_affine(r, a, oinf);
        uint32_t ow[16]; point_encode<Fq>(ow, a, oinf, out_enc);
        words_to_be(ow, out, wire_words<Fq>(out_enc));
    } else {
        uint32_t w[32]; be_to_words(in, w, wire_words<Fq2>(in_enc));
        Aff<Fq2> p; bool inf;
        int rc = point_decode<Fq2>(p, inf, w, in_enc, true);
        if (rc) return rc;
        uint32_t tblmem[8 * 32];
        StridedTable<Fq2> tbl{tblmem, 1};
        Fq2 zr[8];
        Jac<Fq2> r;
        if (inf) r = jac_infinity<Fq2>();
        else if (path == 0 || path == 1) r = mul_window4<Fq2>(p, k, tbl, zr, bad);
        else if (path == 3) r = mul_glv<Fq2>(p, k, tbl, zr, bad);
        else r = mul_binary<Fq2>(p, k);
        Aff<Fq2> a; bool oinf; to_affine(r, a, oinf);
        uint32_t ow[32]; point_encode<Fq2>(ow, a, oinf, out_enc);
        words_to_be(ow, out, wire_words<Fq2>(out_enc));
    }
    if (bad_out) *bad_out = bad;
    return 0;
}
int sim_recode(int g2, const uint8_t *in, uint8_t *out, int in_enc, int out_enc, int check) {
    if (!g2) {
        uint32_t w[16]; be_to_words(in, w, wire_words<Fq>(in_enc));
        Aff<Fq> p; bool inf;
        int rc = point_decode<Fq>(p, inf, w, in_enc, check);
        if (rc) return rc;
        uint32_t ow[16]; point_encode<Fq>(ow, p, inf, out_enc);
        words_to_be(ow, out, wire_words<Fq>(out_enc));
    } else {
        uint32_t w[32]; be_to_words(in, w, wire_words<Fq2>(in_enc));
        Aff<Fq2> p; bool inf;
        int rc = point_decode<Fq2>(p, inf, w, in_enc, check);
        if (rc) return rc;
        uint32_t ow[32]; point_encode<Fq2>(ow, p, inf, out_enc);
        words_to_be(ow, out, wire_words<Fq2>(out_enc));
    }
    return 0;
}
// GLV split: outputs k1, k2 magnitudes as 5 LE words each + signs (after the parity fix is NOT applied here)
void sim_glv(const uint8_t *k_be, uint32_t *k1, uint32_t *k2, int *neg1, int *neg2) {
    uint32_t k[8]; k_from_be(k_be, k);
    GlvSplit s = glv_decompose(k);
    memcpy(k1, s.k1, 20); memcpy(k2, s.k2, 20);
    *neg1 = s.neg1; *neg2 = s.neg2;
}
// width-5 NAF recoding of the GLV halves (the uniform-scalar path): digits LSB first, returns the common length
int sim_uniform_digits(const uint8_t *k_be, int8_t *d1, int8_t *d2) {
    uint32_t k[8]; k_from_be(k_be, k);
    UniformDigits u = uniform_digits(k);
    memcpy(d1, u.d1, 136); memcpy(d2, u.d2, 136);
    return u.len;
}
// field op on canonical BE operands: field 0 = Fq, 1 = Fr; op 0 mul 1 add 2 sub 3 inv 4 neg
void sim_field(int field, int op, const uint8_t *a_be, const uint8_t *b_be, uint8_t *out_be) {
    uint32_t wa[8], wb[8], wo[8];
    memcpy(wa, a_be, 32); memcpy(wb, b_be, 32);
    if (field == 0) {
        Fq a = to_mont(limbs_from_be_words<FqP>(wa)), b = to_mont(limbs_from_be_words<FqP>(wb)), r;
        r = op == 0 ? mul(a, b) : op == 1 ? add(a, b) : op == 2 ? sub(a, b) : op == 3 ? inv(a) : neg(a);
        limbs_to_be_words(from_mont(r), wo);
    } else {
        Fr a = to_mont(limbs_from_be_words<FrP>(wa)), b = to_mont(limbs_from_be_words<FrP>(wb)), r;
        r = op == 0 ? mul(a, b) : op == 1 ? add(a, b) : op == 2 ? sub(a, b) : op == 3 ? inv(a) : neg(a);
        limbs_to_be_words(from_mont(r), wo);
    }
    memcpy(out_be, wo, 32);
}

// XYZZ check: sum_i k_i P_i by plain double-and-add in XYZZ coordinates (exercises xyzz_madd / xyzz_add / xyzz_dbl)
int sim_xyzz_msm(const uint8_t *points, const uint8_t *scalars_be, int n, uint8_t *out) {
    Xyzz<Fq> total = xyzz_infinity<Fq>();
    for (int i = 0; i < n; i++) {
        uint32_t w[16]; be_to_words(points + 64 * i, w, 16);
        Aff<Fq> p; bool inf;
        if (point_decode<Fq>(p, inf, w, ENC_UNCOMPRESSED, true)) return 1;
        uint32_t k[8]; k_from_be(scalars_be + 32 * i, k);
        Xyzz<Fq> acc = xyzz_infinity<Fq>();
        for (int b = 255; b >= 0; b--) {
            acc = xyzz_dbl(acc);
            if (!inf && ((k[b >> 5] >> (b & 31)) & 1)) acc = xyzz_madd(acc, p);
        }
        total = xyzz_add(total, acc);
    }
    bool inf = is_zero(total.zz);
    Fq t = inv(mul(total.zz, total.zzz));
    Aff<Fq> a;
    a.x = mul(total.x, mul(t, total.zzz));
    a.y = mul(total.y, mul(t, total.zz));
    uint32_t ow[16]; point_encode<Fq>(ow, a, inf, ENC_UNCOMPRESSED);
    words_to_be(ow, out, 16);
    return 0;
}
}
