/* p2b.h -- C ABI of the B200 compute core for the phase2-bn254 contribution hot path.
 *
 * The reference (kobigurk/phase2-bn254, Rust) has no FFI for this path: the hot loop is a nested
 * `fn batch_exp` inside each entry point.  This header is the boundary a Rust shim binds with
 * `extern "C"` (INTEGRATION.md shows the shim); every entry point names the reference code it replaces.
 *
 *   level 1  p2b_g{1,2}_batch_mul          -> body of batch_exp, phase2/src/parameters.rs:424-470
 *            p2b_g{1,2}_batch_mul_powers   -> tau-power generation + batch_exp,
 *                                             powersoftau/src/batched_accumulator.rs:1130-1181,1201-1229
 *   level 2  p2b_pot_transform             -> chunk loops of BatchedAccumulator::transform,
 *                                             powersoftau/src/batched_accumulator.rs:1187-1289
 *            p2b_phase2_contribute         -> MPCParameters::contribute on the serialized parameters,
 *                                             phase2/src/parameters.rs:414-522 (+ read/write 663-703)
 *   benches  p2b_g{1,2}_msm                -> multiexp / dense_multiexp, bellman/src/multiexp.rs:330-475
 *            p2b_fr_fft                    -> EvaluationDomain::{fft,ifft,coset_fft,icoset_fft},
 *                                             bellman/src/domain.rs:154-205
 *
 * Conventions
 *   - All buffers are caller-owned; the library borrows them for the duration of the call.
 *   - Plain entry points take HOST pointers (pageable or pinned) and do their own H2D/D2H.
 *     `_dev` entry points take DEVICE pointers on the ctx's GPU and run on p2b_stream(ctx).
 *   - Points use the reference's wire encodings (pairing/src/bn256/ec.rs:763-946,1136-1344):
 *     P2B_ENC_UNCOMPRESSED (G1 64 B x||y, G2 128 B x.c1||x.c0||y.c1||y.c0, big-endian),
 *     P2B_ENC_COMPRESSED (32 / 64 B, bit 7 of byte 0 = y is the larger root), bit 6 of byte 0 = infinity;
 *     P2B_ENC_RAW_MONT_LE = RawEncodable::into_raw_uncompressed_le (ec.rs:653-706; G1: 64 B of
 *     Montgomery limbs, little-endian, all-zero = infinity; G2: same layout over x.c0,x.c1,y.c0,y.c1).
 *   - Scalars are 32-byte big-endian canonical values (< r) = `into_repr().write_be()`.
 *   - Calls are blocking; one ctx is single-caller (not re-entrant); one ctx per GPU.
 *   - Return value: P2B_OK or a P2B_E* code; p2b_last_error() / p2b_error_detail() describe it.
 */
#ifndef P2B_H
#define P2B_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

typedef struct p2b_ctx p2b_ctx;

enum {
    P2B_OK = 0,
    P2B_EARG = 1,          /* size / argument mismatch (the bins' length panics, compute_constrained.rs:96-102) */
    P2B_EDECODE = 2,       /* GroupDecodingError (pairing/src/lib.rs:280-291); sub-code in p2b_error_detail */
    P2B_EINFINITY_IN = 3,  /* DeserializationError::PointAtInfinity (batched_accumulator.rs:987-991) */
    P2B_EINFINITY_OUT = 4, /* "your contribution happened to produce a point at infinity" (:1176-1179) */
    P2B_ECUDA = 5          /* CUDA failure */
};
enum { /* decode sub-codes */
    P2B_DEC_NOT_ON_CURVE = 1,
    P2B_DEC_COORDINATE = 2,
    P2B_DEC_UNEXPECTED_INFORMATION = 3,
    P2B_DEC_UNEXPECTED_COMPRESSION_MODE = 4
};
enum { P2B_ENC_UNCOMPRESSED = 0, P2B_ENC_COMPRESSED = 1, P2B_ENC_RAW_MONT_LE = 2 };
enum { /* flags */
    P2B_CHECK_INPUT = 1,     /* CheckForCorrectness::Yes: is_on_curve on every decoded point */
    P2B_REJECT_INFINITY = 2, /* infinity in the input or the output is an error (phase-1 semantics) */
    P2B_G2_SUBGROUP = 4,     /* the caller vouches that every G2 input lies in the order-r subgroup (true for any point
                                produced by scalar multiplication of the generator): enables the endomorphism-split G2
                                path (~1.4x faster).  Without it G2 results are exact for EVERY on-curve point, like the
                                reference, which decodes G2 without a subgroup check (pairing/src/bn256/ec.rs:1145-1213). */
    P2B_G2_EXACT = 8         /* never use the endomorphism split for G2.  Without P2B_G2_SUBGROUP and P2B_G2_EXACT a G2 batch of
                                >= 2^17 points is first PROVEN to lie in the order-r subgroup by the library itself (random
                                linear combinations of the whole batch, 8 independent 14..16-bit window sums W_j = sum rho_ij P_i
                                tested for [r] W_j = O; every on-curve point outside the subgroup is caught with probability
                                >= 1 - 2^-104 over the library's CSPRNG coefficients, an off-curve point always) and then takes
                                the split path; a batch that fails the test takes the exact path.  Results are the exact [k]P
                                either way. */
};

/* ---- context ---- */
int p2b_init(int device, p2b_ctx **out);
void p2b_destroy(p2b_ctx *ctx);
const char *p2b_last_error(p2b_ctx *ctx);
/* index of the first failing element and the decode sub-code of the last error */
void p2b_error_detail(p2b_ctx *ctx, uint64_t *index, int *sub);
/* cudaStream_t the ctx launches on (for event timing by the caller) */
void *p2b_stream(p2b_ctx *ctx);
/* number of kernels this ctx has launched so far */
uint64_t p2b_launch_count(p2b_ctx *ctx);
const char *p2b_version(void);
/* G2 subgroup probe (see P2B_G2_EXACT): number of batches probed so far by this ctx, and the verdict of the last one
 * (0 = proven in the subgroup, split path taken; 1 = exact path taken; -1 = no probe yet).  Waits for the ctx's stream. */
int p2b_g2_probe_stats(p2b_ctx *ctx, uint64_t *probes, int *last_verdict);
/* Profiling: when enabled, the ctx brackets its dominant kernels with CUDA events on p2b_stream(ctx) (the stream they
 * are launched on).  p2b_profile_read waits for the stream and returns the summed device time and the number of kernel
 * launches recorded in `slot` since p2b_profile_enable(ctx, 1) was last called (which also resets the counters). */
enum {
    P2B_PROF_BATCH_MUL = 0,      /* k_batch_mul: decode + scalar + [k]P */
    P2B_PROF_NORMALIZE = 1,      /* k_normalize: batched inversion + encode */
    P2B_PROF_MSM_SORT = 2,       /* prepare + histogram + scan + scatter */
    P2B_PROF_MSM_ACCUMULATE = 3, /* bucket accumulation */
    P2B_PROF_MSM_REDUCE = 4,     /* bucket / window reduction + final */
    P2B_PROF_FFT_PASS = 5,       /* radix-256 FFT passes */
    P2B_PROF_SLOTS = 6
};
/* Caller buffers may be pageable (the reference's callers pass memory maps of the challenge / response files,
 * powersoftau/src/bin/compute_constrained.rs:83-132): such buffers are staged through pinned rings by host threads
 * (csrc/hostio.cu); page-locked buffers are copied directly.  Bytes that took the staged path so far: */
void p2b_io_stats(p2b_ctx *ctx, uint64_t *staged_h2d_bytes, uint64_t *staged_d2h_bytes);
int p2b_profile_enable(p2b_ctx *ctx, int on);
int p2b_profile_read(p2b_ctx *ctx, int slot, double *total_ms, uint64_t *kernels);

/* ---- level 1: batch_exp ---- */
/* out[i] = [s_i] in[i];  n_scalars == n (one per point) or 1 (broadcast, phase-2 shape). */
int p2b_g1_batch_mul(p2b_ctx *ctx, const uint8_t *in, uint8_t *out, size_t n, const uint8_t *scalars_be32,
                     size_t n_scalars, int in_enc, int out_enc, int flags);
int p2b_g2_batch_mul(p2b_ctx *ctx, const uint8_t *in, uint8_t *out, size_t n, const uint8_t *scalars_be32,
                     size_t n_scalars, int in_enc, int out_enc, int flags);
/* out[i] = [tau^(start_index + i) * coeff] in[i]  (coeff may be NULL = 1). */
int p2b_g1_batch_mul_powers(p2b_ctx *ctx, const uint8_t *in, uint8_t *out, size_t n, const uint8_t tau_be[32],
                            const uint8_t *coeff_be_or_null, uint64_t start_index, int in_enc, int out_enc, int flags);
int p2b_g2_batch_mul_powers(p2b_ctx *ctx, const uint8_t *in, uint8_t *out, size_t n, const uint8_t tau_be[32],
                            const uint8_t *coeff_be_or_null, uint64_t start_index, int in_enc, int out_enc, int flags);
/* device-pointer variants (inputs/outputs resident in HBM; async on p2b_stream until p2b_sync) */
int p2b_g1_batch_mul_dev(p2b_ctx *ctx, const void *d_in, void *d_out, size_t n, const uint8_t *scalars_be32_host,
                         size_t n_scalars, int in_enc, int out_enc, int flags);
int p2b_g2_batch_mul_dev(p2b_ctx *ctx, const void *d_in, void *d_out, size_t n, const uint8_t *scalars_be32_host,
                         size_t n_scalars, int in_enc, int out_enc, int flags);
int p2b_g1_batch_mul_powers_dev(p2b_ctx *ctx, const void *d_in, void *d_out, size_t n, const uint8_t tau_be[32],
                                const uint8_t *coeff_be_or_null, uint64_t start_index, int in_enc, int out_enc,
                                int flags);
int p2b_g2_batch_mul_powers_dev(p2b_ctx *ctx, const void *d_in, void *d_out, size_t n, const uint8_t tau_be[32],
                                const uint8_t *coeff_be_or_null, uint64_t start_index, int in_enc, int out_enc,
                                int flags);
/* waits for the stream and returns the status of the work queued by _dev calls since the last sync */
int p2b_sync(p2b_ctx *ctx);

/* ---- level 2: whole entry points on the reference's file formats ---- */
/* powersoftau geometry (powersoftau/src/parameters.rs:72-120): size of the accumulator part of a file
 * = hash(64) + all points; the response file additionally carries the 768-byte public key. */
uint64_t p2b_pot_accumulator_size(uint32_t size_log2, int compressed);
/* Writes the accumulator region [64, accumulator_size(out)) of `response` exactly as write_chunk would.  Bytes [0,64)
 * (hash of the challenge) and the trailing public key stay with the caller (compute_constrained.rs:155-161,207-209).
 * [shard_index, shard_count): this process transforms only its contiguous share of every section (one process per
 * GPU; shards write disjoint byte ranges of `response`); pass 0, 1 for the whole file.
 * check_input: 0 / 1 = CheckForCorrectness::{No, Yes}; may be OR-ed with P2B_G2_SUBGROUP / P2B_G2_EXACT (see the flags). */
int p2b_pot_transform(p2b_ctx *ctx, const uint8_t *challenge, uint64_t challenge_len, uint8_t *response,
                      uint64_t response_len, uint32_t size_log2, uint32_t batch_size, int in_compressed,
                      int out_compressed, int check_input, const uint8_t tau_be[32], const uint8_t alpha_be[32],
                      const uint8_t beta_be[32], uint32_t shard_index, uint32_t shard_count);
/* BatchedAccumulator::decompress (powersoftau/src/batched_accumulator.rs:543-618): compressed response -> uncompressed
 * accumulator of the next challenge.  Writes [64, accumulator_size(uncompressed)) of `challenge`; the 64-byte hash prefix
 * stays with the caller (verify_transform_constrained.rs:207-229).  Any decoded point at infinity is an error, as in
 * read_points_chunk (:987-991); check_input = CheckForCorrectness. */
int p2b_pot_decompress(p2b_ctx *ctx, const uint8_t *response, uint64_t response_len, uint8_t *challenge,
                       uint64_t challenge_len, uint32_t size_log2, int check_input, uint32_t shard_index,
                       uint32_t shard_count);
/* Bulk point codec (no scalar multiplication): out[i] = in[i] re-encoded.  Decompression (square roots), checked
 * deserialisation (P2B_CHECK_INPUT = is_on_curve, ec.rs:133-148; the point reads of Parameters::read,
 * bellman/src/groth16/mod.rs:287-383) and compression.  Same error reporting as the batch_mul entry points.
 * out == NULL: validation only (every point is decoded and checked on the device, nothing is copied back) -- the verifier's
 * read of a chunk it only needs to be well-formed (batched_accumulator.rs:398-410). */
int p2b_g1_recode(p2b_ctx *ctx, const uint8_t *in, uint8_t *out, size_t n, int in_enc, int out_enc, int flags);
int p2b_g2_recode(p2b_ctx *ctx, const uint8_t *in, uint8_t *out, size_t n, int in_enc, int out_enc, int flags);
/* MPCParameters::contribute over the serialized parameters (`MPCParameters::write` format).  The RNG-derived values
 * are inputs: delta (Fr), s (G1 uncompressed, = G1::rand) and r (G2 uncompressed, = hash_to_g2(transcript) -- the
 * caller computes it from p2b_phase2_transcript).  params_out must hold params_len + 384 bytes.  Returns the
 * 64-byte contribution hash (Blake2b-512 of the new public key) in hash_out. */
int p2b_phase2_transcript(p2b_ctx *ctx, const uint8_t *params, uint64_t params_len, const uint8_t delta_be[32],
                          const uint8_t s_g1[64], uint8_t transcript_out[64]);
int p2b_phase2_contribute(p2b_ctx *ctx, const uint8_t *params, uint64_t params_len, uint8_t *params_out,
                          uint64_t params_out_len, const uint8_t delta_be[32], const uint8_t s_g1[64],
                          const uint8_t r_g2[128], uint8_t hash_out[64]);

/* The same with H and L split into `shard_count` contiguous ranges: rank `shard_index` rewrites only its range of both vectors
 * in params_out (shared storage, e.g. the output file mapped by every rank; no collective, disjoint bytes -- the static range
 * split of the reference's thread chunks, parameters.rs:430-438, applied to GPUs); shard 0 also writes every other byte.  All
 * shards return the same hash. */
int p2b_phase2_contribute_sharded(p2b_ctx *ctx, const uint8_t *params, uint64_t params_len, uint8_t *params_out,
                                  uint64_t params_out_len, const uint8_t delta_be[32], const uint8_t s_g1[64],
                                  const uint8_t r_g2[128], uint8_t hash_out[64], uint32_t shard_index, uint32_t shard_count);

/* ---- Pippenger MSM ---- */
/* out = sum scalars[i] * points[i]; points uncompressed wire, out uncompressed wire (64 / 128 B). */
int p2b_g1_msm(p2b_ctx *ctx, const uint8_t *points, const uint8_t *scalars_be32, size_t n, uint8_t *out);
int p2b_g2_msm(p2b_ctx *ctx, const uint8_t *points, const uint8_t *scalars_be32, size_t n, uint8_t *out);
int p2b_g1_msm_dev(p2b_ctx *ctx, const void *d_points, const void *d_scalars_be32, size_t n, uint8_t *out_host);
int p2b_g2_msm_dev(p2b_ctx *ctx, const void *d_points, const void *d_scalars_be32, size_t n, uint8_t *out_host);
/* The verifier's random linear combinations in ONE pass (next row):
 *   merge_pairs(v1, v2)  = (sum rho_i v1_i, sum rho_i v2_i)      powersoftau/src/utils.rs:112-130, phase2/src/utils.rs:59-105
 *   power_pairs(v)       = merge_pairs(v[..n-1], v[1..])          powersoftau/src/utils.rs:133-135
 * One upload, one decode, one counting sort of the shared scalars and one walk of the sorted entries feed TWO bucket sets
 * (every entry costs two mixed adds; for power_pairs both points of an entry are adjacent in memory).
 *   scalars_be32      n x 32 B big-endian canonical coefficients, or NULL: the coefficients are then generated ON THE DEVICE,
 *                     coefficient i = bytes [32 i, 32 i + 32) of the ChaCha20 keystream (RFC 7539 block function; key = seed,
 *                     64-bit block counter in state words 12-13, words 14-15 zero) read as a big-endian integer and cleared
 *                     above scalar_bits bits -- the reference draws `Fr::rand(thread_rng())` per element; pass 32 bytes of
 *                     OS entropy as the seed.  p2b_random_scalars returns the same values for tests and audits.
 *   scalar_bits       upper bound promised for every coefficient (64..253 when generated, 0 = any canonical scalar when given);
 *                     the windows only cover that many bits (128-bit coefficients halve the work, soundness error 2^-128).
 *   in_enc            P2B_ENC_UNCOMPRESSED or P2B_ENC_COMPRESSED (decompressed on the device, as read_chunk does on the host)
 *   flags             P2B_CHECK_INPUT (is_on_curve, CheckForCorrectness::Yes), P2B_REJECT_INFINITY (read_chunk's rule)
 * p2b_*_power_pairs reads n_points points and combines n_points - 1 terms. */
int p2b_g1_msm_pair(p2b_ctx *ctx, const uint8_t *points_a, const uint8_t *points_b, const uint8_t *scalars_be32, size_t n,
                    const uint8_t seed[32], uint32_t scalar_bits, int in_enc, int flags, uint8_t out_a[64], uint8_t out_b[64]);
int p2b_g2_msm_pair(p2b_ctx *ctx, const uint8_t *points_a, const uint8_t *points_b, const uint8_t *scalars_be32, size_t n,
                    const uint8_t seed[32], uint32_t scalar_bits, int in_enc, int flags, uint8_t out_a[128], uint8_t out_b[128]);
int p2b_g1_power_pairs(p2b_ctx *ctx, const uint8_t *points, size_t n_points, const uint8_t *scalars_be32, const uint8_t seed[32],
                       uint32_t scalar_bits, int in_enc, int flags, uint8_t out_a[64], uint8_t out_b[64]);
int p2b_g2_power_pairs(p2b_ctx *ctx, const uint8_t *points, size_t n_points, const uint8_t *scalars_be32, const uint8_t seed[32],
                       uint32_t scalar_bits, int in_enc, int flags, uint8_t out_a[128], uint8_t out_b[128]);
/* the device-generated coefficients [first_index, first_index + n) of `seed`, as 32-byte big-endian values (host buffer) */
int p2b_random_scalars(p2b_ctx *ctx, const uint8_t seed[32], uint64_t first_index, size_t n, uint32_t scalar_bits, uint8_t *out);
/* multi-GPU: the MSM shards by point range; each rank's result is an ordinary (affine, uncompressed) point, the
 * ranks exchange the 64 / 128-byte results (NCCL all-gather by the caller) and every rank adds them locally:
 * out = sum of `count` uncompressed points (infinity allowed). */
int p2b_g1_sum_points(p2b_ctx *ctx, const uint8_t *points, size_t count, uint8_t out[64]);
int p2b_g2_sum_points(p2b_ctx *ctx, const uint8_t *points, size_t count, uint8_t out[128]);

/* ---- self-test of the device multipliers (used by tests/test_gpu_field_selftest.py) ----
 * out[i] = op(a[i], b[i], c[i], d[i]) on RAW 8 x u32 little-endian limbs (no conversions: Montgomery form is the caller's business).
 * op: 0 a*b/R, 1 a^2/R (dedicated squaring), 2 (a*b + c*d)/R (the fused two-product multiplication), 3 reduce(wide product of a, b),
 * 4 a + b, 5 a - b (mod p); field: 0 Fq, 1 Fr.  Operands must be < p. */
int p2b_selftest_field(p2b_ctx *ctx, int field, int op, const uint8_t *a, const uint8_t *b, const uint8_t *c, const uint8_t *d,
                       size_t n, uint8_t *out);

/* ---- Fr radix-2 FFT ---- */
/* In-place, natural order in and out, 2^log_n scalars of 32 BE bytes; inverse => ifft (x m^-1);
 * coset => coset_fft / icoset_fft with the multiplicative generator 7. */
int p2b_fr_fft(p2b_ctx *ctx, uint8_t *data, uint32_t log_n, int inverse, int coset);
int p2b_fr_fft_dev(p2b_ctx *ctx, void *d_data, uint32_t log_n, int inverse, int coset);

/* ---- group-element FFT and prepare_phase2 ---- */
/* EvaluationDomain<E, Point<G>>::{fft, ifft} (bellman/src/domain.rs:154-174, group.rs:30-50): radix-2 transform over
 * 2^log_d POINTS, natural order in and out; inverse => omega^-1 and x d^-1.  Infinity is a valid element. */
int p2b_g1_group_fft(p2b_ctx *ctx, const uint8_t *in, uint8_t *out, uint32_t log_d, int inverse, int in_enc, int out_enc,
                     int flags);
int p2b_g2_group_fft(p2b_ctx *ctx, const uint8_t *in, uint8_t *out, uint32_t log_d, int inverse, int in_enc, int out_enc,
                     int flags);
/* Multi-GPU group FFT (dist.sharded_group_fft): a transform of 2^total_log_d points block-distributed over R = 2^k GPUs is k
 * rank-crossing decimation-in-frequency stages -- each rank computes half of the butterflies it shares with its partner,
 *     out_sum[i] = a[i] + b[i],   out_diff[i] = [w^(start + i)] (a[i] - b[i])      (n = 2^log_n pairs; w = omega^(2^stage))
 * and exchanges the halves (NCCL send / recv of 64 / 128 B points by the caller: about 1 % of a stage's compute time) -- followed
 * by an ordinary transform of the rank's 2^(total_log_d - k) points whose inverse scaling is that of the whole domain
 * (p2b_*_group_fft_scaled).  Rank r then holds X[R j + bitrev_k(r)].  w == NULL in p2b_*_gfft_stage skips the multiplication
 * (plain sums / differences: the H query tau^(i+d) G - tau^i G of prepare_phase2.rs:132-148); either output may be NULL.
 * p2b_fr_root_of_unity: omega_d (inverse: omega_d^-1), bellman/src/domain.rs:52-99. */
int p2b_g1_group_fft_scaled(p2b_ctx *ctx, const uint8_t *in, uint8_t *out, uint32_t log_d, int inverse, int in_enc, int out_enc,
                            int flags, uint32_t total_log_d);
int p2b_g2_group_fft_scaled(p2b_ctx *ctx, const uint8_t *in, uint8_t *out, uint32_t log_d, int inverse, int in_enc, int out_enc,
                            int flags, uint32_t total_log_d);
int p2b_g1_gfft_stage(p2b_ctx *ctx, const uint8_t *a, const uint8_t *b, uint32_t log_n, const uint8_t *w_be_or_null, uint64_t start,
                      int in_enc, int out_enc, int flags, uint8_t *out_sum, uint8_t *out_diff);
int p2b_g2_gfft_stage(p2b_ctx *ctx, const uint8_t *a, const uint8_t *b, uint32_t log_n, const uint8_t *w_be_or_null, uint64_t start,
                      int in_enc, int out_enc, int flags, uint8_t *out_sum, uint8_t *out_diff);
int p2b_fr_root_of_unity(uint32_t log_d, int inverse, uint8_t out_be[32]);
/* One iteration of powersoftau/src/bin/prepare_phase2.rs:62-241: from an accumulator (challenge layout, uncompressed, or
 * response layout, compressed; the 64-byte hash prefix included) to the image of the file phase1radix2m{m}:
 * alpha_g1 | beta_g1 | beta_g2 | coeffs_g1[d] | coeffs_g2[d] | alpha_coeffs_g1[d] | beta_coeffs_g1[d] | h[d-1], all
 * uncompressed, d = 2^m (reader: phase2/src/parameters.rs:182-217).  p2b_pot_radix_file_size(m) = 192 + 384 d bytes.
 * check_input = CheckForCorrectness of the deserialisation (the binary uses Yes); flags: 0, P2B_G2_SUBGROUP or P2B_G2_EXACT. */
uint64_t p2b_pot_radix_file_size(uint32_t m);
int p2b_pot_prepare_phase2(p2b_ctx *ctx, const uint8_t *accumulator, uint64_t accumulator_len, uint32_t size_log2,
                           int compressed_input, int check_input, uint32_t m, uint8_t *out, uint64_t out_len, int flags);

/* ---- sparse linear maps over group elements (the QAP evaluation of MPCParameters::new) ----
 * out[i] = sum_{j in [row_offsets[i], row_offsets[i+1])} coeffs[j] * bases[cols[j]]   (phase2/src/parameters.rs:225-300:
 * a_g1, b_g1, b_g2 and ext = ic / l over the Lagrange-basis points of phase1radix2m{m}; rows = variables, entries =
 * (coefficient, constraint index) pairs of the A / B / C matrices).  CSR layout: row_offsets has n_rows + 1 entries
 * starting at 0; bases are uncompressed points (decoded checked); coeffs are 32-byte big-endian canonical scalars;
 * out receives n_rows uncompressed points (empty or cancelling rows give the point at infinity, flag 0x40). */
int p2b_g1_sparse_mul(p2b_ctx *ctx, const uint8_t *bases, size_t n_bases, const uint64_t *row_offsets,
                      const uint32_t *cols, const uint8_t *coeffs_be32, size_t n_rows, uint8_t *out);
int p2b_g2_sparse_mul(p2b_ctx *ctx, const uint8_t *bases, size_t n_bases, const uint64_t *row_offsets,
                      const uint32_t *cols, const uint8_t *coeffs_be32, size_t n_rows, uint8_t *out);

/* ---- verifier host side: pairings, hash_to_g2, key-generation RNG (CPU code, no ctx, no GPU needed) ----
 * The reference keeps these on the CPU too: a verification does <= 20 pairings (same_ratio,
 * powersoftau/src/utils.rs:151-159, phase2/src/utils.rs:48-57) over pairs that merge_pairs / power_pairs
 * (two p2b_g*_msm calls) condense from millions of points.  Points are uncompressed wire encodings, decoded
 * CHECKED (on curve); P2B_EDECODE when one does not decode. */
/* *is_one = (prod_i e(g1_points[i], g2_points[i]) == 1); pairs containing the point at infinity contribute 1 */
int p2b_pairing_check(const uint8_t *g1_points, const uint8_t *g2_points, size_t n, int *is_one);
/* same_ratio((g1_a, g1_b), (g2_a, g2_b)): e(g1_a, g2_b) == e(g1_b, g2_a); false if any point is infinity */
int p2b_same_ratio(const uint8_t g1_a[64], const uint8_t g1_b[64], const uint8_t g2_a[128], const uint8_t g2_b[128],
                   int *same);
/* hash_to_g2 (powersoftau/src/utils.rs:31-45, phase2/src/utils.rs:111-122): ChaChaRng::from_seed(first 32 digest
 * bytes as 8 big-endian u32).gen::<G2>() -- restated from rand 0.4.6 / ff_derive, which are not vendored in the
 * reference tree; no reference vector exists for it in-tree (parity unpinned for this function). */
int p2b_hash_to_g2(const uint8_t digest[32], uint8_t out[128]);
/* ChaChaRng (rand 0.4.6) in a caller-owned state block, and the reference's samplers over it:
 * Fr::rand (canonical value, 32 B big-endian), G1::rand / G2::rand (pairing/src/bn256/ec.rs:711-726,1091-1106). */
#define P2B_RNG_STATE_BYTES 136
int p2b_rng_seed(uint8_t state[P2B_RNG_STATE_BYTES], const uint32_t seed[8]);
int p2b_rng_u32(uint8_t state[P2B_RNG_STATE_BYTES], uint32_t *out);
int p2b_rng_fr(uint8_t state[P2B_RNG_STATE_BYTES], uint8_t out_be32[32]);
int p2b_rng_g1(uint8_t state[P2B_RNG_STATE_BYTES], uint8_t out[64]);
int p2b_rng_g2(uint8_t state[P2B_RNG_STATE_BYTES], uint8_t out[128]);
/* one scalar multiplication on the host (keypair.rs:64-84: g1_s.mul(x), g2_s.mul(x); CurveAffine::mul_bits) */
int p2b_host_g1_mul(const uint8_t point[64], const uint8_t scalar_be32[32], uint8_t out[64]);
int p2b_host_g2_mul(const uint8_t point[128], const uint8_t scalar_be32[32], uint8_t out[128]);
/* test hook: xi^((q-1)/6), its square and cube (Montgomery limbs) = FROBENIUS_COEFF_FQ12_C1[1],
 * FROBENIUS_COEFF_FQ6_C1[1], XI_TO_Q_MINUS_1_OVER_2 of pairing/src/bn256/fq.rs */
int p2b_pairing_constants(uint8_t out[192]);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif
