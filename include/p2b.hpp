// p2b.hpp -- C++ host-side mirror of the reference's interfaces on the contribution hot path, over the C ABI of p2b.h.
//
// The reference is Rust; there is no Rust toolchain in the build image, so this header is the compiled-language host side
// above the boundary: same names, argument meaning and error behaviour as the Rust items it mirrors, header-only, no CUDA
// or torch types.  (phase2_bn254_b200/*.py is the same mirror for Python callers.)
//
//   p2b::powersoftau::CeremonyParams                powersoftau/src/parameters.rs:38-120
//   p2b::powersoftau::{UseCompression, CheckForCorrectness, DeserializationError, PrivateKey}
//                                                   parameters.rs:127-170, keypair.rs:47-51
//   p2b::powersoftau::BatchedAccumulator::{transform, decompress, generate_initial}
//                                                   batched_accumulator.rs:1119-1292, 543-618, 1295-1347
//   p2b::powersoftau::prepare_phase2                powersoftau/src/bin/prepare_phase2.rs:62-241 (one m)
//   p2b::phase2::MPCParameters::{read, write, contribute}
//                                                   phase2/src/parameters.rs:414-522, 663-703
//   p2b::bellman::{dense_multiexp, EvaluationDomain}   bellman/src/multiexp.rs:361-475, domain.rs:52-205
//   p2b::bellman::{merge_pairs_g1, power_pairs_g1/g2}  powersoftau/src/utils.rs:112-135 (one pass on the GPU)
//
// Field elements cross as 32-byte big-endian canonical values (`into_repr().write_be()`), points as the wire encodings
// of pairing/src/bn256/ec.rs; maps are (pointer, length) views of the mmaps.
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include "p2b.h"

namespace p2b {

using Scalar = std::array<uint8_t, 32>;   // Fr, big-endian canonical

struct Error : std::runtime_error {
    int code, sub;
    uint64_t index;
    Error(int c, const std::string &m, uint64_t i = 0, int s = 0) : std::runtime_error(m), code(c), sub(s), index(i) {}
};

// One compute context per GPU (p2b_init / p2b_destroy); every mirror below takes one.
class Context {
   public:
    explicit Context(int device = 0) {
        int rc = p2b_init(device, &h_);
        if (rc != P2B_OK) throw Error(rc, "p2b_init failed: no CUDA device / driver (there is no CPU fallback)");
    }
    ~Context() { p2b_destroy(h_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    p2b_ctx *get() const { return h_; }
    void check(int rc) const {
        if (rc == P2B_OK) return;
        uint64_t idx = 0;
        int sub = 0;
        p2b_error_detail(h_, &idx, &sub);
        throw Error(rc, p2b_last_error(h_), idx, sub);
    }

   private:
    p2b_ctx *h_ = nullptr;
};

namespace powersoftau {

enum class UseCompression { Yes, No };            // parameters.rs:127-131
enum class CheckForCorrectness { Yes, No };       // parameters.rs:136-140

// parameters.rs:143-170
struct DeserializationError : std::runtime_error {
    enum Kind { IoError, DecodingError, PointAtInfinity } kind;
    int group_decoding_error;   // GroupDecodingError: 1 NotOnCurve, 2 CoordinateDecodingError, 3 UnexpectedInformation, 4 UnexpectedCompressionMode
    uint64_t index;
    DeserializationError(Kind k, const std::string &m, uint64_t i = 0, int g = 0)
        : std::runtime_error(m), kind(k), group_decoding_error(g), index(i) {}
};

// Bn256 geometry (parameters.rs:21-34, 72-120)
struct CeremonyParams {
    size_t size, batch_size;
    size_t g1 = 64, g2 = 128, g1_compressed = 32, g2_compressed = 64;
    size_t powers_length, powers_g1_length, accumulator_size, public_key_size, contribution_size, hash_size = 64;
    CeremonyParams(size_t size_, size_t batch_size_) : size(size_), batch_size(batch_size_) {
        powers_length = (size_t)1 << size;
        powers_g1_length = (powers_length << 1) - 1;
        accumulator_size = powers_g1_length * g1 + powers_length * g2 + powers_length * g1 + powers_length * g1 + g2 + hash_size;
        public_key_size = 3 * g2 + 6 * g1;
        contribution_size = powers_g1_length * g1_compressed + powers_length * g2_compressed + powers_length * g1_compressed +
                            powers_length * g1_compressed + g2_compressed + hash_size + public_key_size;
    }
};

struct PrivateKey {   // keypair.rs:47-51
    Scalar tau, alpha, beta;
};

namespace detail {
[[noreturn]] inline void rethrow(const Error &e) {
    if (e.code == P2B_EDECODE)
        throw DeserializationError(DeserializationError::DecodingError, std::string("DecodingError: ") + e.what(), e.index, e.sub);
    if (e.code == P2B_EINFINITY_IN) throw DeserializationError(DeserializationError::PointAtInfinity, e.what(), e.index);
    if (e.code == P2B_EINFINITY_OUT)
        throw std::logic_error("your contribution happened to produce a point at infinity, please re-run");   // the assert at :1176-1179
    throw e;
}
}  // namespace detail

struct BatchedAccumulator {
    // Transforms the accumulator with a private key (batched_accumulator.rs:1119-1292).  Writes output_map[64, accumulator
    // end); the hash prefix and the public key stay with the caller (compute_constrained.rs:155-161,207-209).
    static void transform(const Context &ctx, const uint8_t *input_map, size_t input_len, uint8_t *output_map, size_t output_len,
                          UseCompression input_is_compressed, UseCompression compress_the_output,
                          CheckForCorrectness check_input_for_correctness, const PrivateKey &key, const CeremonyParams &parameters,
                          uint32_t shard_index = 0, uint32_t shard_count = 1, bool g2_in_subgroup = false) {
        try {
            ctx.check(p2b_pot_transform(ctx.get(), input_map, input_len, output_map, output_len, (uint32_t)parameters.size,
                                        (uint32_t)parameters.batch_size, input_is_compressed == UseCompression::Yes,
                                        compress_the_output == UseCompression::Yes,
                                        (check_input_for_correctness == CheckForCorrectness::Yes ? 1 : 0) |
                                            (g2_in_subgroup ? P2B_G2_SUBGROUP : 0),
                                        key.tau.data(), key.alpha.data(), key.beta.data(), shard_index, shard_count));
        } catch (const Error &e) {
            detail::rethrow(e);
        }
    }
    // Compressed response -> uncompressed accumulator (batched_accumulator.rs:543-618).
    static void decompress(const Context &ctx, const uint8_t *input_map, size_t input_len, uint8_t *output_map, size_t output_len,
                           CheckForCorrectness check_input_for_correctness, const CeremonyParams &parameters) {
        try {
            ctx.check(p2b_pot_decompress(ctx.get(), input_map, input_len, output_map, output_len, (uint32_t)parameters.size,
                                         check_input_for_correctness == CheckForCorrectness::Yes, 0, 1));
        } catch (const Error &e) {
            detail::rethrow(e);
        }
    }
    // Initial accumulator: every element is the generator (batched_accumulator.rs:1295-1347); uncompressed output.
    static void generate_initial(uint8_t *output_map, size_t output_len, const CeremonyParams &p) {
        if (output_len < p.accumulator_size) throw std::invalid_argument("output map too small");
        uint8_t g1[64] = {0}, g2[128];
        g1[31] = 1;
        g1[63] = 2;                                                        // (1, 2), ec.rs:1013-1051
        static const char *g2hex =                                          // fq.rs:54-83, x.c1 | x.c0 | y.c1 | y.c0
            "198e9393920d483a7260bfb731fb5d25f1aa493335a9e71297e485b7aef312c2"
            "1800deef121f1e76426a00665e5c4479674322d4f75edadd46debd5cd992f6ed"
            "090689d0585ff075ec9e99ad690c3395bc4b313370b38ef355acdadcd122975b"
            "12c85ea5db8c6deb4aab71808dcb408fe3d1e7690c43d37b4ce6cc0166fa7daa";
        for (int i = 0; i < 128; i++) {
            auto nib = [](char ch) { return ch <= '9' ? ch - '0' : ch - 'a' + 10; };
            g2[i] = (uint8_t)(nib(g2hex[2 * i]) << 4 | nib(g2hex[2 * i + 1]));
        }
        uint8_t *o = output_map + 64;
        for (size_t i = 0; i < p.powers_g1_length; i++, o += 64) memcpy(o, g1, 64);
        for (size_t i = 0; i < p.powers_length; i++, o += 128) memcpy(o, g2, 128);
        for (size_t i = 0; i < 2 * p.powers_length; i++, o += 64) memcpy(o, g1, 64);
        memcpy(o, g2, 128);
    }
};

// One iteration of bin/prepare_phase2.rs:62-241: the bytes of the file phase1radix2m{m}.
inline std::vector<uint8_t> prepare_phase2(const Context &ctx, const uint8_t *accumulator_map, size_t len, const CeremonyParams &parameters,
                                           uint32_t m, UseCompression input_is_compressed = UseCompression::Yes,
                                           CheckForCorrectness check = CheckForCorrectness::Yes) {
    std::vector<uint8_t> out(p2b_pot_radix_file_size(m));
    try {
        ctx.check(p2b_pot_prepare_phase2(ctx.get(), accumulator_map, len, (uint32_t)parameters.size,
                                         input_is_compressed == UseCompression::Yes, check == CheckForCorrectness::Yes, m, out.data(),
                                         out.size(), 0));
    } catch (const Error &e) {
        detail::rethrow(e);
    }
    return out;
}

}  // namespace powersoftau

namespace phase2 {

// MPCParameters kept in its serialized form (MPCParameters::write, parameters.rs:663-677): the GPU path consumes and
// produces wire bytes.
class MPCParameters {
   public:
    std::vector<uint8_t> data;

    static MPCParameters read(const uint8_t *bytes, size_t len) {
        MPCParameters p;
        p.data.assign(bytes, bytes + len);
        return p;
    }
    void write(std::vector<uint8_t> &writer) const { writer.insert(writer.end(), data.begin(), data.end()); }

    // contribute (parameters.rs:414-522).  The reference draws delta, s = G1::rand and r = hash_to_g2(transcript) from its
    // RNG (keypair, :860-908); here they are explicit.  Returns the 64-byte contribution hash.
    std::array<uint8_t, 64> contribute(const Context &ctx, const Scalar &delta, const uint8_t s_g1[64], const uint8_t r_g2[128]) {
        std::vector<uint8_t> out(data.size() + 384);
        std::array<uint8_t, 64> hash;
        ctx.check(p2b_phase2_contribute(ctx.get(), data.data(), data.size(), out.data(), out.size(), delta.data(), s_g1, r_g2, hash.data()));
        data.swap(out);
        return hash;
    }
    std::array<uint8_t, 64> transcript(const Context &ctx, const Scalar &delta, const uint8_t s_g1[64]) const {
        std::array<uint8_t, 64> t;
        ctx.check(p2b_phase2_transcript(ctx.get(), data.data(), data.size(), delta.data(), s_g1, t.data()));
        return t;
    }
};


// eval() of MPCParameters::new (parameters.rs:244-300) as one sparse linear map over group elements: CSR rows = variables,
// (coeff, lag) entries of the A / B / C matrices; out[i] = sum_j coeffs[j] * bases[cols[j]] (uncompressed, infinity allowed).
inline std::vector<uint8_t> sparse_eval_g1(const Context &ctx, const uint8_t *bases, size_t n_bases, const std::vector<uint64_t> &row_offsets,
                                           const std::vector<uint32_t> &cols, const std::vector<Scalar> &coeffs) {
    if (row_offsets.empty() || cols.size() != coeffs.size() || row_offsets.back() != cols.size()) throw std::invalid_argument("sparse_eval: inconsistent CSR arrays");
    std::vector<uint8_t> out((row_offsets.size() - 1) * 64), k(coeffs.size() * 32);
    for (size_t i = 0; i < coeffs.size(); i++) memcpy(&k[32 * i], coeffs[i].data(), 32);
    ctx.check(p2b_g1_sparse_mul(ctx.get(), bases, n_bases, row_offsets.data(), cols.data(), k.data(), row_offsets.size() - 1, out.data()));
    return out;
}
inline std::vector<uint8_t> sparse_eval_g2(const Context &ctx, const uint8_t *bases, size_t n_bases, const std::vector<uint64_t> &row_offsets,
                                           const std::vector<uint32_t> &cols, const std::vector<Scalar> &coeffs) {
    if (row_offsets.empty() || cols.size() != coeffs.size() || row_offsets.back() != cols.size()) throw std::invalid_argument("sparse_eval: inconsistent CSR arrays");
    std::vector<uint8_t> out((row_offsets.size() - 1) * 128), k(coeffs.size() * 32);
    for (size_t i = 0; i < coeffs.size(); i++) memcpy(&k[32 * i], coeffs[i].data(), 32);
    ctx.check(p2b_g2_sparse_mul(ctx.get(), bases, n_bases, row_offsets.data(), cols.data(), k.data(), row_offsets.size() - 1, out.data()));
    return out;
}

}  // namespace phase2

// ---- host side of the verifiers and of key generation (CPU code in libp2b.so, no Context needed) ----
// same_ratio (powersoftau/src/utils.rs:151-159, phase2/src/utils.rs:48-57): e(g1.0, g2.1) == e(g1.1, g2.0); false if any point is zero.
inline bool same_ratio(const uint8_t g1_0[64], const uint8_t g1_1[64], const uint8_t g2_0[128], const uint8_t g2_1[128]) {
    int same = 0;
    if (p2b_same_ratio(g1_0, g1_1, g2_0, g2_1, &same) != P2B_OK) throw std::invalid_argument("same_ratio: a point does not decode");
    return same != 0;
}
// hash_to_g2 (utils.rs:31-45): uncompressed G2 point from the first 32 bytes of `digest`.
inline std::array<uint8_t, 128> hash_to_g2(const uint8_t *digest) {
    std::array<uint8_t, 128> out;
    if (p2b_hash_to_g2(digest, out.data()) != P2B_OK) throw std::invalid_argument("hash_to_g2");
    return out;
}
// rand 0.4.6 ChaChaRng::from_seed(&[u32; 8]) with the samplers keypair() uses (keypair.rs:54-103, parameters.rs:860-908).
class ChaChaRng {
    uint8_t st_[P2B_RNG_STATE_BYTES];
   public:
    explicit ChaChaRng(const uint32_t seed[8]) { p2b_rng_seed(st_, seed); }
    uint32_t next_u32() { uint32_t v = 0; p2b_rng_u32(st_, &v); return v; }
    Scalar gen_fr() { Scalar k; p2b_rng_fr(st_, k.data()); return k; }
    std::array<uint8_t, 64> gen_g1() { std::array<uint8_t, 64> p; p2b_rng_g1(st_, p.data()); return p; }
    std::array<uint8_t, 128> gen_g2() { std::array<uint8_t, 128> p; p2b_rng_g2(st_, p.data()); return p; }
};

namespace bellman {

// dense_multiexp (multiexp.rs:361-475): sum exponents[i] * bases[i]; bases uncompressed wire, exponents 32-byte BE repr.
inline std::array<uint8_t, 64> dense_multiexp_g1(const Context &ctx, const uint8_t *bases, const uint8_t *exponents, size_t n) {
    std::array<uint8_t, 64> out;
    ctx.check(p2b_g1_msm(ctx.get(), bases, exponents, n, out.data()));
    return out;
}
inline std::array<uint8_t, 128> dense_multiexp_g2(const Context &ctx, const uint8_t *bases, const uint8_t *exponents, size_t n) {
    std::array<uint8_t, 128> out;
    ctx.check(p2b_g2_msm(ctx.get(), bases, exponents, n, out.data()));
    return out;
}

// merge_pairs / power_pairs of the verifiers (powersoftau/src/utils.rs:112-135, phase2/src/utils.rs:59-105): the random linear
// combination (sum rho_i v1_i, sum rho_i v2_i) in ONE pass on the GPU.  The coefficients are generated on the device from `seed`
// (32 bytes of OS entropy; ChaCha20 keystream) -- the reference draws Fr::rand(thread_rng()) per element.  `enc` is the encoding
// the points have in the file (compressed chunks are decompressed on the device); flags: P2B_CHECK_INPUT, P2B_REJECT_INFINITY.
struct PointPairG1 { std::array<uint8_t, 64> s, sx; };
struct PointPairG2 { std::array<uint8_t, 128> s, sx; };
inline PointPairG1 merge_pairs_g1(const Context &ctx, const uint8_t *v1, const uint8_t *v2, size_t n, const uint8_t seed[32],
                                  uint32_t scalar_bits = 253, int enc = P2B_ENC_UNCOMPRESSED, int flags = 0) {
    PointPairG1 r;
    ctx.check(p2b_g1_msm_pair(ctx.get(), v1, v2, nullptr, n, seed, scalar_bits, enc, flags, r.s.data(), r.sx.data()));
    return r;
}
inline PointPairG1 power_pairs_g1(const Context &ctx, const uint8_t *v, size_t n_points, const uint8_t seed[32], uint32_t scalar_bits = 253,
                                  int enc = P2B_ENC_UNCOMPRESSED, int flags = 0) {
    PointPairG1 r;
    ctx.check(p2b_g1_power_pairs(ctx.get(), v, n_points, nullptr, seed, scalar_bits, enc, flags, r.s.data(), r.sx.data()));
    return r;
}
inline PointPairG2 power_pairs_g2(const Context &ctx, const uint8_t *v, size_t n_points, const uint8_t seed[32], uint32_t scalar_bits = 253,
                                  int enc = P2B_ENC_UNCOMPRESSED, int flags = 0) {
    PointPairG2 r;
    ctx.check(p2b_g2_power_pairs(ctx.get(), v, n_points, nullptr, seed, scalar_bits, enc, flags, r.s.data(), r.sx.data()));
    return r;
}

// EvaluationDomain over Fr (domain.rs:52-205): coefficients padded with zeros to the next power of two.
class EvaluationDomain {
   public:
    std::vector<uint8_t> coeffs;   // m x 32 bytes, big-endian canonical
    uint32_t exp = 0;

    static EvaluationDomain from_coeffs(const std::vector<Scalar> &c) {
        if (c.size() > (((size_t)1 << 28) - 1)) throw std::length_error("PolynomialDegreeTooLarge");   // SynthesisError, domain.rs:64-78
        EvaluationDomain d;
        size_t m = 1;
        while (m < c.size()) { m *= 2; d.exp++; }
        d.coeffs.assign(m * 32, 0);
        for (size_t i = 0; i < c.size(); i++) memcpy(&d.coeffs[32 * i], c[i].data(), 32);
        return d;
    }
    void fft(const Context &ctx) { ctx.check(p2b_fr_fft(ctx.get(), coeffs.data(), exp, 0, 0)); }
    void ifft(const Context &ctx) { ctx.check(p2b_fr_fft(ctx.get(), coeffs.data(), exp, 1, 0)); }
    void coset_fft(const Context &ctx) { ctx.check(p2b_fr_fft(ctx.get(), coeffs.data(), exp, 0, 1)); }
    void icoset_fft(const Context &ctx) { ctx.check(p2b_fr_fft(ctx.get(), coeffs.data(), exp, 1, 1)); }
};

}  // namespace bellman
}  // namespace p2b
