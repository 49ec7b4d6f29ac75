// codec.cuh -- BN254 point wire codecs (big-endian, flag bits in byte 0) on the GPU.
//
// Mirrors pairing/src/bn256/ec.rs: G1Uncompressed 763-843, G1Compressed 866-946, G2Uncompressed
// 1136-1232, G2Compressed 1255-1344, get_point_from_x 110-131, RawEncodable 653-706, and the square
// roots they rely on (Fq: ff_derive sqrt for q = 3 mod 4; Fq2: fq2.rs:206-262, Algorithm 9 of
// eprint 2012/685).  Error sub-codes follow GroupDecodingError (pairing/src/lib.rs:280-291).
#pragma once
#include "ec.cuh"

namespace p2b {

enum : int { ENC_UNCOMPRESSED = 0, ENC_COMPRESSED = 1, ENC_RAW_MONT_LE = 2 };
enum : int { DEC_OK = 0, DEC_NOT_ON_CURVE = 1, DEC_COORD = 2, DEC_UNEXPECTED_INFO = 3, DEC_UNEXPECTED_COMPRESSION = 4 };

template <class F> struct Wire;
template <> struct Wire<Fq> {
    static constexpr int WORDS_UNCOMPRESSED = 16, WORDS_COMPRESSED = 8;
};
template <> struct Wire<Fq2> {
    static constexpr int WORDS_UNCOMPRESSED = 32, WORDS_COMPRESSED = 16;
};
template <class F> P2B_HD constexpr int wire_words(int enc) {
    return enc == ENC_COMPRESSED ? Wire<F>::WORDS_COMPRESSED : Wire<F>::WORDS_UNCOMPRESSED;
}

// ---- field element <-> 8 native words holding 32 big-endian bytes -------------------------------------------------
// returns false when the value is not a canonical residue (PrimeFieldDecodingError::NotInField)
P2B_HD bool fq_from_wire(Fq &out, const uint32_t *w, uint32_t first_word_mask) {
    uint32_t t[8];
#pragma unroll
    for (int i = 0; i < 8; i++) t[i] = w[i];
    t[0] &= first_word_mask;
    Fq c = limbs_from_be_words<FqP>(t);
    bool ok = is_canonical(c);
    out = to_mont(c);
    return ok;
}
P2B_HD void fq_to_wire(const Fq &a_mont, uint32_t *w) { limbs_to_be_words(from_mont(a_mont), w); }

// mask that clears the two flag bits of byte 0 when applied to the first native (little-endian) word
static constexpr uint32_t FLAG_CLEAR_MASK = 0xffffff3fu;
P2B_HD uint32_t flags_of(uint32_t first_word) { return first_word & 0xc0u; }   // bit7 = 0x80, bit6 = 0x40 of byte 0

// ---- square roots ----------------------------------------------------------------------------------------------------
P2B_DEF_CONST(FQ_EXP_QM3D4, {0xb61f3f51u, 0x4f082305u, 0x5a1c72a3u, 0x65e05aa4u, 0xa0605617u, 0x6e14116du, 0xb84c680au, 0x0c19139cu})
P2B_DEF_CONST(FQ_EXP_QM1D2, {0x6c3e7ea3u, 0x9e10460bu, 0xb438e546u, 0xcbc0b548u, 0x40c0ac2eu, 0xdc2822dbu, 0x7098d014u, 0x18322739u})

// a1 = a^((q-3)/4); a0 = a1^2 a; a0 == -1 -> no root; else root = a1 a
P2B_HD bool fq_sqrt(Fq &out, const Fq &a) {
    Fq a1 = fp_one<FqP>();
    bool started = false;
#pragma unroll 1
    for (int i = 255; i >= 0; i--) {
        bool bit = (P2B_C(FQ_EXP_QM3D4, i >> 5) >> (i & 31)) & 1;
        if (started) a1 = sqr(a1);
        if (bit) { a1 = started ? mul(a1, a) : a; started = true; }
    }
    Fq a0 = mul(sqr(a1), a);
    Fq m1 = neg(fp_one<FqP>());
    out = mul(a1, a);
    return !eq(a0, m1);
}
template <int WHICH> P2B_HD Fq2 fq2_pow_const(const Fq2 &a) {
    Fq2 res = fq2_one();
    bool started = false;
#pragma unroll 1
    for (int i = 255; i >= 0; i--) {
        uint32_t w = WHICH == 0 ? P2B_C(FQ_EXP_QM3D4, i >> 5) : P2B_C(FQ_EXP_QM1D2, i >> 5);
        bool bit = (w >> (i & 31)) & 1;
        if (started) res = sqr(res);
        if (bit) { res = started ? mul(res, a) : a; started = true; }
    }
    return res;
}
P2B_HD bool fq2_sqrt(Fq2 &out, const Fq2 &a) {
    if (is_zero(a)) { out = fq2_zero(); return true; }
    Fq2 a1 = fq2_pow_const<0>(a);
    Fq2 alpha = mul(sqr(a1), a);
    Fq2 a0 = mul(conj(alpha), alpha);
    Fq2 neg1 = fq2_one();
    neg1.c0 = neg(neg1.c0);
    if (eq(a0, neg1)) return false;
    a1 = mul(a1, a);
    if (eq(alpha, neg1)) {
        Fq2 u; u.c0 = fp_zero<FqP>(); u.c1 = fp_one<FqP>();
        a1 = mul(a1, u);
    } else {
        alpha = fq2_pow_const<1>(add(alpha, fq2_one()));
        a1 = mul(a1, alpha);
    }
    out = a1;
    return true;
}
P2B_HD bool field_sqrt(Fq &o, const Fq &a) { return fq_sqrt(o, a); }
P2B_HD bool field_sqrt(Fq2 &o, const Fq2 &a) { return fq2_sqrt(o, a); }

// ---- coordinates <-> wire words ---------------------------------------------------------------------------------------
// Fq: one 32-byte value.  Fq2: c1 first, then c0 (ec.rs:1187-1190, 1224-1227).
P2B_HD bool coord_from_wire(Fq &out, const uint32_t *w, uint32_t mask0) { return fq_from_wire(out, w, mask0); }
P2B_HD bool coord_from_wire(Fq2 &out, const uint32_t *w, uint32_t mask0) {
    bool ok1 = fq_from_wire(out.c1, w, mask0);
    bool ok0 = fq_from_wire(out.c0, w + 8, 0xffffffffu);
    return ok1 & ok0;
}
P2B_HD void coord_to_wire(const Fq &a, uint32_t *w) { fq_to_wire(a, w); }
P2B_HD void coord_to_wire(const Fq2 &a, uint32_t *w) { fq_to_wire(a.c1, w); fq_to_wire(a.c0, w + 8); }

P2B_HD bool raw_in_field(const Fq &a) { return is_canonical(a); }
P2B_HD bool raw_in_field(const Fq2 &a) { return is_canonical(a.c0) & is_canonical(a.c1); }

// ---- decode -----------------------------------------------------------------------------------------------------------
// w: the encoding as native words (wire_words<F>(enc) of them).  On success fills p / inf.  `check` = is_on_curve for
// uncompressed input (CheckForCorrectness::Yes).  Returns a DEC_* code.
template <class F> P2B_HD int point_decode(Aff<F> &p, bool &inf, const uint32_t *w, int enc, bool check) {
    constexpr bool IS_G2 = FieldTraits<F>::WORDS == 16;
    const int n = wire_words<F>(enc);
    inf = false;
    p.x = FieldTraits<F>::zero();
    p.y = FieldTraits<F>::one();
    if (enc == ENC_RAW_MONT_LE) {
        // RawEncodable (G1 only): x || y as Montgomery limbs, little-endian; all-zero = infinity (ec.rs:653-706)
        uint32_t any = 0;
        for (int i = 0; i < n; i++) any |= w[i];
        if (!any) { inf = true; return DEC_OK; }
        for (int i = 0; i < FieldTraits<F>::WORDS; i++) { set_word(p.x, i, w[i]); set_word(p.y, i, w[FieldTraits<F>::WORDS + i]); }
        if (!(raw_in_field(p.x) & raw_in_field(p.y))) return DEC_COORD;        // Fq::from_raw_repr rejects limbs >= q
        if (check && !on_curve(p)) return DEC_NOT_ON_CURVE;
        return DEC_OK;
    }
    uint32_t fl = flags_of(w[0]);
    if (IS_G2 && enc == ENC_UNCOMPRESSED && (fl & 0x80u)) return DEC_UNEXPECTED_COMPRESSION;   // ec.rs:1158-1161
    if (fl & 0x40u) {                                                                           // infinity flag
        uint32_t any = w[0] & FLAG_CLEAR_MASK;
        for (int i = 1; i < n; i++) any |= w[i];
        if (any) return DEC_UNEXPECTED_INFO;
        inf = true;
        return DEC_OK;
    }
    if (enc == ENC_UNCOMPRESSED) {
        if (fl & 0x80u) return DEC_UNEXPECTED_INFO;                                            // ec.rs:797-801 (G1)
        bool okx = coord_from_wire(p.x, w, FLAG_CLEAR_MASK);
        bool oky = coord_from_wire(p.y, w + n / 2, 0xffffffffu);
        if (!(okx & oky)) return DEC_COORD;
        if (check && !on_curve(p)) return DEC_NOT_ON_CURVE;
        return DEC_OK;
    }
    bool greatest = (fl & 0x80u) != 0;
    if (!coord_from_wire(p.x, w, FLAG_CLEAR_MASK)) return DEC_COORD;
    F x3b = add(mul(sqr(p.x), p.x), curve_b((const F *)nullptr));
    F y;
    if (!field_sqrt(y, x3b)) return DEC_NOT_ON_CURVE;
    // (y < -y) ^ greatest ? y : -y      (ec.rs:123-127)
    bool y_is_larger = is_lexicographically_largest(y);
    bool y_lt_negy = !y_is_larger && !is_zero(y);
    p.y = (y_lt_negy != greatest) ? y : neg(y);
    return DEC_OK;
}

// ---- encode -----------------------------------------------------------------------------------------------------------
template <class F> P2B_HD void point_encode(uint32_t *w, const Aff<F> &p, bool inf, int enc) {
    const int n = wire_words<F>(enc);
    if (enc == ENC_RAW_MONT_LE) {
        for (int i = 0; i < FieldTraits<F>::WORDS; i++) {
            w[i] = inf ? 0u : get_word(p.x, i);
            w[FieldTraits<F>::WORDS + i] = inf ? 0u : get_word(p.y, i);
        }
        return;
    }
    if (inf) {
        for (int i = 0; i < n; i++) w[i] = 0;
        w[0] = 0x40u;
        return;
    }
    coord_to_wire(p.x, w);
    if (enc == ENC_UNCOMPRESSED) { coord_to_wire(p.y, w + n / 2); return; }
    if (is_lexicographically_largest(p.y)) w[0] |= 0x80u;
}

}  // namespace p2b
