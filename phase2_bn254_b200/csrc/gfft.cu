// gfft.cu -- radix-2 FFT / iFFT over GROUP elements (G1 / G2 points) and the prepare_phase2 step built on it.
//
// GPU replacement for bellman's EvaluationDomain<E, Point<G>>::{fft, ifft} (bellman/src/domain.rs:154-174,274-317 with the
// Group impl of bellman/src/group.rs:30-50: a butterfly multiplies a POINT by the twiddle, i.e. a full 254-bit scalar
// multiplication) as used by powersoftau/src/bin/prepare_phase2.rs:62-241 to turn the powers of tau into Lagrange
// coefficients, plus the H query (tau^(i+d) - tau^i) and the phase1radix2m{m} file image.  SURVEY.md 8f rank 1.
//
// Schedule: decimation in frequency, natural order in, bit-reversed out, then one permutation.  Per stage
//   k_gbutterfly   (a, b) -> a + b and a - b as Jacobian points (mixed adds on affine inputs)
//   k_normalize    batched inversion -> affine (Montgomery, raw) again          [batch_mul_impl.cuh]
//   k_gtwiddles    w^(pos << s) as 32-byte scalars for the (a - b) half
//   k_batch_mul    (a - b) * twiddle: the batch_exp kernel, per-point scalars   [launch_batch_mul]
//   k_gscatter     results back to their positions
// The inverse transform multiplies by d^-1 with one more broadcast batch_exp fused with the final encode.  The d/2 log d
// scalar multiplications dominate (>= 97 % of the time); everything else is plumbing around the hot kernel.
#include <cstring>
#include "batch_mul_impl.cuh"

namespace p2b {

void host_domain_constants(uint32_t log_n, int inverse, uint32_t omega_mont[8], uint32_t ninv_canon[8]);   // fft.cu

template <class F> __device__ __forceinline__ void load_raw_point(const uint32_t *base, size_t i, Aff<F> &a, bool &inf) {
    constexpr int WU = Wire<F>::WORDS_UNCOMPRESSED;
    uint32_t w[WU];
    load_words<WU>(w, base + i * WU);
    point_decode<F>(a, inf, w, ENC_RAW_MONT_LE, false);
}
// p + q for affine operands either of which may be infinity
template <class F> __device__ __forceinline__ Jac<F> aff_add(const Aff<F> &p, bool p_inf, const Aff<F> &q, bool q_inf) {
    Jac<F> r = jac_madd(p_inf ? jac_infinity<F>() : jac_from_aff(p), q);
    if (q_inf) r = p_inf ? jac_infinity<F>() : jac_from_aff(p);
    return r;
}

// stage s of a DIF transform over d = 2^log_d points: butterfly b pairs positions i0, i1 = i0 + half
__device__ __forceinline__ void butterfly_index(uint32_t b, uint32_t log_d, uint32_t s, size_t &i0, size_t &i1, uint32_t &pos) {
    const uint32_t lh = log_d - 1 - s;
    pos = b & ((1u << lh) - 1u);
    i0 = ((size_t)(b >> lh) << (lh + 1)) + pos;
    i1 = i0 + ((size_t)1 << lh);
}
template <class F> __global__ void __launch_bounds__(128) k_gbutterfly(const uint32_t *X, uint32_t *jx, uint32_t *jy, uint32_t *jz,
                                                                       uint32_t log_d, uint32_t s) {
    const size_t nb = (size_t)1 << (log_d - 1);
    for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += (size_t)gridDim.x * blockDim.x) {
        size_t i0, i1;
        uint32_t pos;
        butterfly_index((uint32_t)b, log_d, s, i0, i1, pos);
        Aff<F> p, q;
        bool pi, qi;
        load_raw_point<F>(X, i0, p, pi);
        load_raw_point<F>(X, i1, q, qi);
        Jac<F> sum = aff_add(p, pi, q, qi);
        q.y = neg(q.y);
        Jac<F> dif = aff_add(p, pi, q, qi);
        store_elem<F>(jx, b, sum.x); store_elem<F>(jy, b, sum.y); store_elem<F>(jz, b, sum.z);
        store_elem<F>(jx, nb + b, dif.x); store_elem<F>(jy, nb + b, dif.y); store_elem<F>(jz, nb + b, dif.z);
    }
}
// twiddles of stage s as big-endian canonical scalars: S[b] = omega^(pos << s)
static __global__ void __launch_bounds__(128) k_gtwiddles(uint32_t *S, Fr omega, uint32_t log_d, uint32_t s) {
    const size_t nb = (size_t)1 << (log_d - 1);
    for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += (size_t)gridDim.x * blockDim.x) {
        size_t i0, i1;
        uint32_t pos;
        butterfly_index((uint32_t)b, log_d, s, i0, i1, pos);
        Fr w = from_mont(pow_u64(omega, (uint64_t)pos << s));
        uint32_t o[8];
        limbs_to_be_words(w, o);
        store_words<8>(S + b * 8, o);
    }
}
// X[i0] = N[b] (sums), X[i1] = M[b] (twiddled differences)
template <class F> __global__ void __launch_bounds__(128) k_gscatter(uint32_t *X, const uint32_t *N, const uint32_t *M, uint32_t log_d, uint32_t s) {
    constexpr int WU = Wire<F>::WORDS_UNCOMPRESSED;
    const size_t nb = (size_t)1 << (log_d - 1);
    for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += (size_t)gridDim.x * blockDim.x) {
        size_t i0, i1;
        uint32_t pos;
        butterfly_index((uint32_t)b, log_d, s, i0, i1, pos);
        uint32_t w[WU];
        load_words<WU>(w, N + b * WU);
        store_words<WU>(X + i0 * WU, w);
        load_words<WU>(w, M + b * WU);
        store_words<WU>(X + i1 * WU, w);
    }
}
template <class F> __global__ void __launch_bounds__(128) k_gbitrev(const uint32_t *X, uint32_t *T, uint32_t log_d) {
    constexpr int WU = Wire<F>::WORDS_UNCOMPRESSED;
    const size_t d = (size_t)1 << log_d;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < d; i += (size_t)gridDim.x * blockDim.x) {
        const size_t j = log_d ? (size_t)(__brev((uint32_t)i) >> (32 - log_d)) : 0;
        uint32_t w[WU];
        load_words<WU>(w, X + j * WU);
        store_words<WU>(T + i * WU, w);
    }
}
// H query of prepare_phase2.rs:132-148: h[i] = P[i + d] - P[i]
template <class F> __global__ void __launch_bounds__(128) k_gdiff(const uint32_t *P, uint32_t *jx, uint32_t *jy, uint32_t *jz, size_t d, size_t count) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        Aff<F> hi, lo;
        bool hinf, linf;
        load_raw_point<F>(P, i + d, hi, hinf);
        load_raw_point<F>(P, i, lo, linf);
        lo.y = neg(lo.y);
        Jac<F> r = aff_add(hi, hinf, lo, linf);
        store_elem<F>(jx, i, r.x); store_elem<F>(jy, i, r.y); store_elem<F>(jz, i, r.z);
    }
}

static int grid_for(Ctx *c, size_t n) {
    size_t g = (n + 127) / 128;
    if (g > (size_t)c->sm_count * 12) g = (size_t)c->sm_count * 12;
    return g ? (int)g : 1;
}

template <class F> static int normalize_to(Ctx *c, const uint32_t *jx, const uint32_t *jy, const uint32_t *jz, uint32_t *prefix, void *out,
                                           size_t n, int out_enc) {
    if (!n) return P2B_OK;
    size_t threads = (n + 31) / 32;
    if (threads < (size_t)c->sm_count * 128) threads = n < (size_t)c->sm_count * 128 ? n : (size_t)c->sm_count * 128;
    NormalizeParams np{jx, jy, jz, prefix, (uint32_t *)out, n, out_enc, 0, c->d_err, 0};
    k_normalize<F><<<(int)((threads + 127) / 128), 128, 0, c->stream>>>(np);
    c->launches++;
    return P2B_OK;
}

// Work area carved out of c->gfft for a transform of d points
template <class F> struct GfftArea {
    uint32_t *X, *N, *M, *S, *jx, *jy, *jz, *prefix;
    static size_t bytes(size_t d) {
        constexpr size_t P = Wire<F>::WORDS_UNCOMPRESSED * 4, E = FieldTraits<F>::WORDS * 4;
        return d * P * 2 + (d / 2 + 1) * P + (d / 2 + 1) * 32 + 4 * d * E + 4096;
    }
    void carve(void *base, size_t d) {
        constexpr size_t PW = Wire<F>::WORDS_UNCOMPRESSED, EW = FieldTraits<F>::WORDS;
        uint32_t *p = (uint32_t *)base;
        X = p; p += d * PW;
        N = p; p += d * PW;
        M = p; p += (d / 2 + 1) * PW;
        S = p; p += (d / 2 + 1) * 8;
        jx = p; p += d * EW;
        jy = p; p += d * EW;
        jz = p; p += d * EW;
        prefix = p;
    }
};

// In: a.X holds d = 2^log_d affine points (raw Montgomery).  Out: natural-order transform written to d_out in out_enc.
// total_log_d (>= log_d): the inverse transform scales by 2^-total_log_d -- the block-local part of a transform of 2^total_log_d
// points sharded over 2^(total_log_d - log_d) GPUs (dist.sharded_group_fft); 0 = log_d.
template <class F> static int group_fft_dev(Ctx *c, GfftArea<F> &a, uint32_t log_d, int inverse, int g2, void *d_out, int out_enc, int flags,
                                            uint32_t total_log_d = 0) {
    const size_t d = (size_t)1 << log_d;
    uint32_t omega[8], ninv[8];
    host_domain_constants(log_d, inverse, omega, ninv);
    if (total_log_d > log_d) {
        uint32_t unused[8];
        host_domain_constants(total_log_d, inverse, unused, ninv);
    }
    Fr w;
    memcpy(w.l, omega, 32);
    int rc;
    for (uint32_t s = 0; s < log_d; s++) {
        const size_t nb = d / 2;
        k_gbutterfly<F><<<grid_for(c, nb), 128, 0, c->stream>>>(a.X, a.jx, a.jy, a.jz, log_d, s);
        c->launches++;
        if ((rc = normalize_to<F>(c, a.jx, a.jy, a.jz, a.prefix, a.N, d, ENC_RAW_MONT_LE))) return rc;
        const uint32_t *second = a.N + nb * Wire<F>::WORDS_UNCOMPRESSED;
        if (s + 1 < log_d) {       // the last stage only has the twiddle w^0 = 1
            k_gtwiddles<<<grid_for(c, nb), 128, 0, c->stream>>>(a.S, w, log_d, s);
            c->launches++;
            ScalarSpec sc;
            memset(&sc, 0, sizeof sc);
            sc.mode = 0;
            sc.d_scalars = a.S;
            if ((rc = launch_batch_mul(c, g2, second, a.M, nb, sc, ENC_RAW_MONT_LE, ENC_RAW_MONT_LE, flags & (P2B_G2_SUBGROUP | P2B_G2_EXACT), 0))) return rc;
            second = a.M;
        }
        k_gscatter<F><<<grid_for(c, nb), 128, 0, c->stream>>>(a.X, a.N, second, log_d, s);
        c->launches++;
    }
    k_gbitrev<F><<<grid_for(c, d), 128, 0, c->stream>>>(a.X, a.N, log_d);
    c->launches++;
    ScalarSpec sc;
    memset(&sc, 0, sizeof sc);
    if (inverse) {                 // x d^-1 (domain.rs:163-173), fused with the final encode
        sc.mode = 1;
        memcpy(sc.k, ninv, 32);
    } else sc.mode = 3;
    if ((rc = launch_batch_mul(c, g2, a.N, d_out, d, sc, ENC_RAW_MONT_LE, out_enc, flags & (P2B_G2_SUBGROUP | P2B_G2_EXACT), 0))) return rc;
    P2B_CUDA(c, cudaGetLastError());
    return P2B_OK;
}

// wire points (device) -> a.X as raw Montgomery affine; decode errors are reported with err_base
static int to_raw(Ctx *c, int g2, const void *d_wire, void *d_raw, size_t n, int in_enc, int flags, uint64_t err_base) {
    ScalarSpec sc;
    memset(&sc, 0, sizeof sc);
    sc.mode = 3;
    return launch_batch_mul(c, g2, d_wire, d_raw, n, sc, in_enc, ENC_RAW_MONT_LE, flags & (P2B_CHECK_INPUT | P2B_REJECT_INFINITY), err_base);
}

template <class F> static int group_fft_host(Ctx *c, int g2, const uint8_t *in, uint8_t *out, uint32_t log_d, int inverse, int in_enc, int out_enc,
                                             int flags, uint32_t total_log_d = 0) {
    if (!in || !out) return ctx_fail(c, P2B_EARG, "null buffer");
    if (log_d > 28 || total_log_d > 28 || (total_log_d && total_log_d < log_d))
        return ctx_fail(c, P2B_EARG, "group fft: log_d <= total_log_d <= 28 (Fr::S)");
    if (in_enc < 0 || in_enc > 2 || out_enc < 0 || out_enc > 2) return ctx_fail(c, P2B_EARG, "bad encoding");
    P2B_CUDA(c, cudaSetDevice(c->device));
    c->last_error.clear();
    P2B_CUDA(c, cudaMemsetAsync(c->d_err, 0xff, sizeof(unsigned long long), c->stream));
    const size_t d = (size_t)1 << log_d, isz = enc_size(g2, in_enc), osz = enc_size(g2, out_enc);
    int rc;
    if ((rc = dev_reserve(c, c->gfft, GfftArea<F>::bytes(d)))) return rc;
    if ((rc = dev_reserve(c, c->stage_in[0], d * isz))) return rc;
    if ((rc = dev_reserve(c, c->stage_out[0], d * osz))) return rc;
    GfftArea<F> a;
    a.carve(c->gfft.p, d);
    if ((rc = io_h2d(c, c->stage_in[0].p, in, d * isz, c->stream))) return rc;
    if ((rc = to_raw(c, g2, c->stage_in[0].p, a.X, d, in_enc, flags, 0))) return rc;
    if ((rc = group_fft_dev<F>(c, a, log_d, inverse, g2, c->stage_out[0].p, out_enc, flags, total_log_d))) return rc;
    if ((rc = io_d2h(c, out, c->stage_out[0].p, d * osz, c->stream))) return rc;
    return ctx_collect_error(c);
}

// One rank-crossing stage of a transform sharded over several GPUs: out_sum[i] = a[i] + b[i], out_diff[i] = [w^(start + i)] (a[i] - b[i])
// for n = 2^log_n point pairs (w == NULL: plain differences, the H query of prepare_phase2.rs:132-148).  The same kernels as one
// stage of group_fft_dev: butterfly -> batched normalisation -> batch_exp with the powers of w generated on the device.
template <class F> static int gfft_stage_host(Ctx *c, int g2, const uint8_t *pa, const uint8_t *pb, uint32_t log_n, const uint8_t *w_be,
                                              uint64_t start, int in_enc, int out_enc, int flags, uint8_t *out_sum, uint8_t *out_diff) {
    constexpr int WU = Wire<F>::WORDS_UNCOMPRESSED;
    if (!pa || !pb || (!out_sum && !out_diff)) return ctx_fail(c, P2B_EARG, "null buffer");
    if (log_n > 27) return ctx_fail(c, P2B_EARG, "gfft stage: log_n must be <= 27");
    if (in_enc < 0 || in_enc > 2 || out_enc < 0 || out_enc > 2) return ctx_fail(c, P2B_EARG, "bad encoding");
    P2B_CUDA(c, cudaSetDevice(c->device));
    c->last_error.clear();
    P2B_CUDA(c, cudaMemsetAsync(c->d_err, 0xff, sizeof(unsigned long long), c->stream));
    const size_t n = (size_t)1 << log_n, d = 2 * n, isz = enc_size(g2, in_enc), osz = enc_size(g2, out_enc);
    int rc;
    if ((rc = dev_reserve(c, c->gfft, GfftArea<F>::bytes(d)))) return rc;
    if ((rc = dev_reserve(c, c->stage_in[0], d * isz))) return rc;
    if ((rc = dev_reserve(c, c->stage_out[0], d * osz))) return rc;
    GfftArea<F> a;
    a.carve(c->gfft.p, d);
    char *st = (char *)c->stage_in[0].p, *so = (char *)c->stage_out[0].p;
    if ((rc = io_h2d(c, st, pa, n * isz, c->stream))) return rc;
    if ((rc = io_h2d(c, st + n * isz, pb, n * isz, c->stream))) return rc;
    if ((rc = to_raw(c, g2, st, a.X, d, in_enc, flags, 0))) return rc;
    k_gbutterfly<F><<<grid_for(c, n), 128, 0, c->stream>>>(a.X, a.jx, a.jy, a.jz, log_n + 1, 0);     // pairs (i, i + n)
    c->launches++;
    if ((rc = normalize_to<F>(c, a.jx, a.jy, a.jz, a.prefix, a.N, d, ENC_RAW_MONT_LE))) return rc;
    ScalarSpec sc;
    memset(&sc, 0, sizeof sc);
    sc.mode = 3;
    if (out_sum) {
        if ((rc = launch_batch_mul(c, g2, a.N, so, n, sc, ENC_RAW_MONT_LE, out_enc, 0, 0))) return rc;
        if ((rc = io_d2h(c, out_sum, so, n * osz, c->stream))) return rc;
    }
    if (out_diff) {
        if (w_be) {
            memset(&sc, 0, sizeof sc);
            sc.mode = 2;
            sc.start = start;
            if (!read_scalar_be(w_be, sc.tau)) return ctx_fail(c, P2B_EARG, "twiddle base not canonical");
            sc.coeff[0] = 1;
        }
        if ((rc = launch_batch_mul(c, g2, a.N + n * WU, so + n * osz, n, sc, ENC_RAW_MONT_LE, out_enc, flags & (P2B_G2_SUBGROUP | P2B_G2_EXACT), 0))) return rc;
        if ((rc = io_d2h(c, out_diff, so + n * osz, n * osz, c->stream))) return rc;
    }
    return ctx_collect_error(c);
}

// ------------------------------------------------------------------------------------------------- prepare_phase2
static uint64_t radix_file_size(uint32_t m) { return 192 + 384 * ((uint64_t)1 << m); }
static uint64_t pot_acc_size(uint32_t size_log2, int compressed) {
    uint64_t p = 1ull << size_log2, pg1 = 2 * p - 1, s1 = compressed ? 32 : 64, s2 = compressed ? 64 : 128;
    return pg1 * s1 + p * s2 + 2 * p * s1 + s2 + 64;
}

// One degree d = 2^m of prepare_phase2.rs:62-241: the image of the file phase1radix2m{m}.
static int prepare_phase2(Ctx *c, const uint8_t *acc, uint64_t acc_len, uint32_t size_log2, int compressed, int check, uint32_t m, uint8_t *out,
                          uint64_t out_len, int flags) {
    if (!acc || !out) return ctx_fail(c, P2B_EARG, "null buffer");
    if (size_log2 == 0 || size_log2 > 28 || m > size_log2) return ctx_fail(c, P2B_EARG, "m must be <= size_log2 <= 28");
    if (acc_len < pot_acc_size(size_log2, compressed)) return ctx_fail(c, P2B_EARG, "accumulator buffer too small");
    if (out_len < radix_file_size(m)) return ctx_fail(c, P2B_EARG, "output buffer too small");
    P2B_CUDA(c, cudaSetDevice(c->device));
    c->last_error.clear();
    P2B_CUDA(c, cudaMemsetAsync(c->d_err, 0xff, sizeof(unsigned long long), c->stream));
    const uint64_t powers = 1ull << size_log2, powers_g1 = 2 * powers - 1, d = 1ull << m;
    const uint64_t g1 = compressed ? 32 : 64, g2s = compressed ? 64 : 128;
    const int enc = compressed ? P2B_ENC_COMPRESSED : P2B_ENC_UNCOMPRESSED;
    const int dflags = (check ? P2B_CHECK_INPUT : 0) | P2B_REJECT_INFINITY;      // deserialize(): checked, infinity rejected
    const uint64_t off_tau_g1 = 64, off_tau_g2 = off_tau_g1 + powers_g1 * g1, off_alpha = off_tau_g2 + powers * g2s,
                   off_beta = off_alpha + powers * g1, off_beta_g2 = off_beta + powers * g1;
    int rc;
    const size_t area = GfftArea<Fq2>::bytes(d) > GfftArea<Fq>::bytes(2 * d) ? GfftArea<Fq2>::bytes(d) : GfftArea<Fq>::bytes(2 * d);
    if ((rc = dev_reserve(c, c->gfft, area))) return rc;
    if ((rc = dev_reserve(c, c->stage_in[0], 2 * d * 128))) return rc;
    if ((rc = dev_reserve(c, c->stage_out[0], radix_file_size(m)))) return rc;
    uint8_t *img = (uint8_t *)c->stage_out[0].p;
    char *stg = (char *)c->stage_in[0].p;
    // header: alpha_g1[0], beta_g1[0], beta_g2 (uncompressed)
    {
        P2B_CUDA(c, cudaMemcpyAsync(stg, acc + off_alpha, g1, cudaMemcpyHostToDevice, c->stream));
        P2B_CUDA(c, cudaMemcpyAsync(stg + 128, acc + off_beta, g1, cudaMemcpyHostToDevice, c->stream));
        P2B_CUDA(c, cudaMemcpyAsync(stg + 256, acc + off_beta_g2, g2s, cudaMemcpyHostToDevice, c->stream));
        ScalarSpec sc;
        memset(&sc, 0, sizeof sc);
        sc.mode = 3;
        if ((rc = launch_batch_mul(c, 0, stg, img, 1, sc, enc, P2B_ENC_UNCOMPRESSED, dflags, 0))) return rc;
        if ((rc = launch_batch_mul(c, 0, stg + 128, img + 64, 1, sc, enc, P2B_ENC_UNCOMPRESSED, dflags, 0))) return rc;
        if ((rc = launch_batch_mul(c, 1, stg + 256, img + 128, 1, sc, enc, P2B_ENC_UNCOMPRESSED, dflags, 0))) return rc;
        if ((rc = ctx_collect_error(c))) return rc;
    }
    // Lagrange coefficients: ifft of the first d powers of each vector (prepare_phase2.rs:68-105)
    uint64_t o = 256;
    const struct { int g2; uint64_t off; } vecs[4] = {{0, off_tau_g1}, {1, off_tau_g2}, {0, off_alpha}, {0, off_beta}};
    for (int v = 0; v < 4; v++) {
        const uint64_t isz = vecs[v].g2 ? g2s : g1, osz = vecs[v].g2 ? 128 : 64;
        if ((rc = io_h2d(c, stg, acc + vecs[v].off, d * isz, c->stream))) return rc;
        if (vecs[v].g2) {
            GfftArea<Fq2> a;
            a.carve(c->gfft.p, d);
            if ((rc = to_raw(c, 1, stg, a.X, d, enc, dflags, 0))) return rc;
            if ((rc = group_fft_dev<Fq2>(c, a, m, 1, 1, img + o, P2B_ENC_UNCOMPRESSED, flags))) return rc;
        } else {
            GfftArea<Fq> a;
            a.carve(c->gfft.p, d);
            if ((rc = to_raw(c, 0, stg, a.X, d, enc, dflags, 0))) return rc;
            if ((rc = group_fft_dev<Fq>(c, a, m, 1, 0, img + o, P2B_ENC_UNCOMPRESSED, flags))) return rc;
        }
        if ((rc = ctx_collect_error(c))) return rc;      // per vector: the failing index is vector-relative
        o += d * osz;
    }
    // H query: tau^(i+d) G - tau^i G for i < d - 1 (prepare_phase2.rs:132-148)
    if (d > 1) {
        const uint64_t np = 2 * d - 1;
        GfftArea<Fq> a;
        a.carve(c->gfft.p, 2 * d);
        if ((rc = io_h2d(c, stg, acc + off_tau_g1, np * g1, c->stream))) return rc;
        if ((rc = to_raw(c, 0, stg, a.X, np, enc, dflags, 0))) return rc;
        k_gdiff<Fq><<<grid_for(c, d - 1), 128, 0, c->stream>>>(a.X, a.jx, a.jy, a.jz, d, d - 1);
        c->launches++;
        if ((rc = normalize_to<Fq>(c, a.jx, a.jy, a.jz, a.prefix, img + o, d - 1, ENC_UNCOMPRESSED))) return rc;
    }
    if ((rc = io_d2h(c, out, img, radix_file_size(m), c->stream))) return rc;
    return ctx_collect_error(c);
}

}  // namespace p2b

using namespace p2b;
extern "C" {
int p2b_g1_group_fft(p2b_ctx *h, const uint8_t *in, uint8_t *out, uint32_t log_d, int inverse, int in_enc, int out_enc, int flags) { P2B_RANGE("p2b_g1_group_fft");
    return h ? group_fft_host<Fq>(&h->c, 0, in, out, log_d, inverse, in_enc, out_enc, flags) : P2B_EARG;
}
int p2b_g2_group_fft(p2b_ctx *h, const uint8_t *in, uint8_t *out, uint32_t log_d, int inverse, int in_enc, int out_enc, int flags) { P2B_RANGE("p2b_g2_group_fft");
    return h ? group_fft_host<Fq2>(&h->c, 1, in, out, log_d, inverse, in_enc, out_enc, flags) : P2B_EARG;
}
int p2b_g1_group_fft_scaled(p2b_ctx *h, const uint8_t *in, uint8_t *out, uint32_t log_d, int inverse, int in_enc, int out_enc, int flags,
                            uint32_t total_log_d) { P2B_RANGE("p2b_g1_group_fft_scaled");
    return h ? group_fft_host<Fq>(&h->c, 0, in, out, log_d, inverse, in_enc, out_enc, flags, total_log_d) : P2B_EARG;
}
int p2b_g2_group_fft_scaled(p2b_ctx *h, const uint8_t *in, uint8_t *out, uint32_t log_d, int inverse, int in_enc, int out_enc, int flags,
                            uint32_t total_log_d) { P2B_RANGE("p2b_g2_group_fft_scaled");
    return h ? group_fft_host<Fq2>(&h->c, 1, in, out, log_d, inverse, in_enc, out_enc, flags, total_log_d) : P2B_EARG;
}
int p2b_g1_gfft_stage(p2b_ctx *h, const uint8_t *a, const uint8_t *b, uint32_t log_n, const uint8_t *w_be, uint64_t start, int in_enc,
                      int out_enc, int flags, uint8_t *out_sum, uint8_t *out_diff) { P2B_RANGE("p2b_g1_gfft_stage");
    return h ? gfft_stage_host<Fq>(&h->c, 0, a, b, log_n, w_be, start, in_enc, out_enc, flags, out_sum, out_diff) : P2B_EARG;
}
int p2b_g2_gfft_stage(p2b_ctx *h, const uint8_t *a, const uint8_t *b, uint32_t log_n, const uint8_t *w_be, uint64_t start, int in_enc,
                      int out_enc, int flags, uint8_t *out_sum, uint8_t *out_diff) { P2B_RANGE("p2b_g2_gfft_stage");
    return h ? gfft_stage_host<Fq2>(&h->c, 1, a, b, log_n, w_be, start, in_enc, out_enc, flags, out_sum, out_diff) : P2B_EARG;
}
/* omega_d^(+-1) (32-byte big-endian canonical) of the 2^log_d domain: the twiddle base of a sharded transform's rank-crossing stages */
int p2b_fr_root_of_unity(uint32_t log_d, int inverse, uint8_t out_be[32]) {
    if (log_d > 28 || !out_be) return P2B_EARG;
    uint32_t omega[8], ninv[8];
    host_domain_constants(log_d, inverse, omega, ninv);
    Fr w;
    memcpy(w.l, omega, 32);
    w = from_mont(w);
    for (int i = 0; i < 8; i++) { uint32_t v = w.l[7 - i]; out_be[4 * i] = v >> 24; out_be[4 * i + 1] = v >> 16; out_be[4 * i + 2] = v >> 8; out_be[4 * i + 3] = v; }
    return P2B_OK;
}
uint64_t p2b_pot_radix_file_size(uint32_t m) { return radix_file_size(m); }
int p2b_pot_prepare_phase2(p2b_ctx *h, const uint8_t *accumulator, uint64_t accumulator_len, uint32_t size_log2, int compressed_input,
                           int check_input, uint32_t m, uint8_t *out, uint64_t out_len, int flags) { P2B_RANGE("p2b_pot_prepare_phase2");
    return h ? prepare_phase2(&h->c, accumulator, accumulator_len, size_log2, compressed_input, check_input, m, out, out_len, flags)
             : P2B_EARG;
}
}
