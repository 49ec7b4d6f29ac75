// host_verify.cu -- the verifier's host side: pairing checks, hash_to_g2, key-generation RNG.
//
// The reference does these few operations per verification on the CPU as well (<= 20 pairings, a handful of single
// scalar multiplications); the bulk work that feeds them -- merge_pairs / power_pairs over millions of points -- is
// the GPU MSM (p2b_g{1,2}_msm).  No CUDA calls in this file, no ctx: usable on a machine without a GPU.
#include <cstring>
#include <memory>
#include <vector>
#include "../../include/p2b.h"
#include "p2b_internal.h"
#include "pairing.cuh"

using namespace p2b;
using namespace p2b::pairing;

namespace {
struct RngState {            // layout of the caller-owned 136-byte p2b_rng state
    uint32_t state[16], buf[16];
    int32_t index;
    uint32_t magic;
};
static_assert(sizeof(RngState) == P2B_RNG_STATE_BYTES, "p2b_rng state size");
constexpr uint32_t RNG_MAGIC = 0x70326272u;
bool rng_load(ChaChaRng &r, const uint8_t *st) {
    RngState s;
    memcpy(&s, st, sizeof s);
    if (s.magic != RNG_MAGIC || s.index < 0 || s.index > 16) return false;
    memcpy(r.state, s.state, sizeof s.state);
    memcpy(r.buf, s.buf, sizeof s.buf);
    r.index = s.index;
    return true;
}
void rng_store(uint8_t *st, const ChaChaRng &r) {
    RngState s;
    memcpy(s.state, r.state, sizeof s.state);
    memcpy(s.buf, r.buf, sizeof s.buf);
    s.index = r.index;
    s.magic = RNG_MAGIC;
    memcpy(st, &s, sizeof s);
}
void fr_to_be(uint8_t out[32], const Fr &canonical) {
    for (int i = 0; i < 8; i++) {
        const uint32_t v = canonical.l[7 - i];
        out[4 * i] = (uint8_t)(v >> 24); out[4 * i + 1] = (uint8_t)(v >> 16); out[4 * i + 2] = (uint8_t)(v >> 8); out[4 * i + 3] = (uint8_t)v;
    }
}
bool scalar_from_be(uint32_t k[8], const uint8_t be[32]) {
    for (int i = 0; i < 8; i++) k[7 - i] = ((uint32_t)be[4 * i] << 24) | ((uint32_t)be[4 * i + 1] << 16) | ((uint32_t)be[4 * i + 2] << 8) | be[4 * i + 3];
    Fr f;
    for (int i = 0; i < 8; i++) f.l[i] = k[i];
    return is_canonical(f);
}
template <class F> int host_mul(const uint8_t *point, const uint8_t scalar_be[32], uint8_t *out) {
    Aff<F> a;
    bool inf;
    if (int rc = decode_host<F>(a, inf, point, true)) { (void)rc; return P2B_EDECODE; }
    uint32_t k[8];
    if (!scalar_from_be(k, scalar_be)) return P2B_EARG;
    Aff<F> r = a;
    bool r_inf = true;
    if (!inf) r_inf = !jac_to_aff_host(r, mul_bits_host(a, k, 256));
    encode_host<F>(out, r, r_inf);
    return P2B_OK;
}
}  // namespace

extern "C" {

int p2b_pairing_check(const uint8_t *g1_points, const uint8_t *g2_points, size_t n, int *is_one) { P2B_RANGE("p2b_pairing_check");
    if (!is_one || (n && (!g1_points || !g2_points))) return P2B_EARG;
    std::vector<Aff<Fq>> ps(n);
    std::vector<Aff<Fq2>> qs(n);
    std::unique_ptr<bool[]> pi(new bool[n + 1]), qi(new bool[n + 1]);
    for (size_t i = 0; i < n; i++) {
        bool inf;
        if (decode_host<Fq>(ps[i], inf, g1_points + 64 * i, true)) return P2B_EDECODE;
        pi[i] = inf;
        if (decode_host<Fq2>(qs[i], inf, g2_points + 128 * i, true)) return P2B_EDECODE;
        qi[i] = inf;
    }
    *is_one = pairing_product_is_one(ps.data(), pi.get(), qs.data(), qi.get(), n) ? 1 : 0;
    return P2B_OK;
}

int p2b_same_ratio(const uint8_t g1_a[64], const uint8_t g1_b[64], const uint8_t g2_a[128], const uint8_t g2_b[128], int *same) { P2B_RANGE("p2b_same_ratio");
    if (!same || !g1_a || !g1_b || !g2_a || !g2_b) return P2B_EARG;
    Aff<Fq> p[2];
    Aff<Fq2> q[2];
    bool pinf[2], qinf[2];
    // e(g1_a, g2_b) == e(g1_b, g2_a)   <=>   e(g1_a, g2_b) * e(-g1_b, g2_a) == 1
    if (decode_host<Fq>(p[0], pinf[0], g1_a, true) || decode_host<Fq>(p[1], pinf[1], g1_b, true)) return P2B_EDECODE;
    if (decode_host<Fq2>(q[0], qinf[0], g2_b, true) || decode_host<Fq2>(q[1], qinf[1], g2_a, true)) return P2B_EDECODE;
    if (pinf[0] || pinf[1] || qinf[0] || qinf[1]) { *same = 0; return P2B_OK; }    // utils.rs:155-157
    p[1].y = neg(p[1].y);
    *same = pairing_product_is_one(p, pinf, q, qinf, 2) ? 1 : 0;
    return P2B_OK;
}

int p2b_hash_to_g2(const uint8_t digest[32], uint8_t out[128]) { P2B_RANGE("p2b_hash_to_g2");
    if (!digest || !out) return P2B_EARG;
    Aff<Fq2> a;
    const bool ok = hash_to_g2(a, digest);
    encode_host<Fq2>(out, a, !ok);
    return P2B_OK;
}

int p2b_rng_seed(uint8_t state[P2B_RNG_STATE_BYTES], const uint32_t seed[8]) {
    if (!state || !seed) return P2B_EARG;
    ChaChaRng r(seed);
    memset(r.buf, 0, sizeof r.buf);
    rng_store(state, r);
    return P2B_OK;
}
int p2b_rng_u32(uint8_t state[P2B_RNG_STATE_BYTES], uint32_t *out) {
    ChaChaRng r;
    if (!state || !out || !rng_load(r, state)) return P2B_EARG;
    *out = r.next_u32();
    rng_store(state, r);
    return P2B_OK;
}
int p2b_rng_fr(uint8_t state[P2B_RNG_STATE_BYTES], uint8_t out_be32[32]) {
    ChaChaRng r;
    if (!state || !out_be32 || !rng_load(r, state)) return P2B_EARG;
    fr_to_be(out_be32, from_mont(fr_rand_mont(r)));
    rng_store(state, r);
    return P2B_OK;
}
int p2b_rng_g1(uint8_t state[P2B_RNG_STATE_BYTES], uint8_t out[64]) {
    ChaChaRng r;
    if (!state || !out || !rng_load(r, state)) return P2B_EARG;
    Aff<Fq> a;
    g1_rand(a, r);
    encode_host<Fq>(out, a, false);
    rng_store(state, r);
    return P2B_OK;
}
int p2b_rng_g2(uint8_t state[P2B_RNG_STATE_BYTES], uint8_t out[128]) {
    ChaChaRng r;
    if (!state || !out || !rng_load(r, state)) return P2B_EARG;
    Aff<Fq2> a;
    const bool ok = g2_rand(a, r);
    encode_host<Fq2>(out, a, !ok);
    rng_store(state, r);
    return P2B_OK;
}

int p2b_host_g1_mul(const uint8_t point[64], const uint8_t scalar_be32[32], uint8_t out[64]) { P2B_RANGE("p2b_host_g1_mul");
    if (!point || !scalar_be32 || !out) return P2B_EARG;
    return host_mul<Fq>(point, scalar_be32, out);
}
int p2b_host_g2_mul(const uint8_t point[128], const uint8_t scalar_be32[32], uint8_t out[128]) { P2B_RANGE("p2b_host_g2_mul");
    if (!point || !scalar_be32 || !out) return P2B_EARG;
    return host_mul<Fq2>(point, scalar_be32, out);
}

// test hook: gamma1, gamma1^2, gamma1^3 as 3 x 64 bytes of Montgomery limbs (c0 then c1, little-endian words), to be
// compared with the reference's hard-coded Frobenius tables (pairing/src/bn256/fq.rs:106-119,121-199,280-432)
int p2b_pairing_constants(uint8_t out[192]) {
    if (!out) return P2B_EARG;
    const FrobConsts &k = frob_consts();
    const Fq2 *g[3] = {&k.g1, &k.g2, &k.g3};
    for (int i = 0; i < 3; i++) {
        memcpy(out + 64 * i, g[i]->c0.l, 32);
        memcpy(out + 64 * i + 32, g[i]->c1.l, 32);
    }
    return P2B_OK;
}

}  // extern "C"
