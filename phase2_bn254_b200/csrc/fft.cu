// fft.cu -- K6: radix-2 FFT / iFFT (and the coset variants) over the BN254 scalar field Fr on sm_100a.
//
// GPU replacement for bellman's EvaluationDomain::{fft, ifft, coset_fft, icoset_fft} over Scalar<E>
// (bellman/src/domain.rs:154-205) and the best_fft / serial_fft / parallel_fft kernels under them
// (domain.rs:263-376).  Natural order in, natural order out, in place from the caller's point of view;
// every output is the canonical residue of a uniquely defined field element, so the schedule is free.
//
// Schedule: Stockham autosort, ceil(log n / 8) passes over HBM, each pass a radix-2^r (r <= 8) step done in
// shared memory:
//   pass with Ns = product of the earlier radices, for column j in [0, n / R):
//       v[t]  = in[j + t * n/R] * w^(t * (j mod Ns)),  w = omega_(Ns*R)            (twiddle at load)
//       V     = DFT_R(v)                                                           (r DIF stages in smem)
//       out[(j div Ns) * Ns * R + (j mod Ns) + q * Ns] = V[q]                      (autosort scatter)
// A block owns a tile of J = 2048 / R consecutive columns so that every global access is a run of J (first
// pass: R) consecutive 32-byte elements.  The first pass decodes the wire form (big-endian words; no Montgomery conversion
// is needed, see the load step), the last pass encodes it again and applies the 1/n of the inverse transform, so the data
// makes exactly `passes` round trips through HBM.  Algorithmic traffic: 64 B per element; actual: 64 B x passes.
// Twiddles: w^e for e < n comes from a two-level table (2^ceil(log n / 2) + 2^floor(log n / 2) entries, L2
// resident); the in-tile twiddles omega_256^t sit in shared memory.
#include <cstdio>
#include <cstring>
#include "fp.cuh"
#include "p2b_internal.h"

namespace p2b {

static constexpr int FFT_BLOCK = 256;
static constexpr int FFT_TILE_LOG = 11;   // elements per tile (64 KB of shared memory + padding)
static constexpr int FFT_RMAX = 8;        // stages per pass
#ifndef FFT_REG_STAGES
#define FFT_REG_STAGES 3                  // last stages of a pass done in registers (3: radix-8 groups, 2: radix-4)
#endif
static constexpr uint32_t FFT_DIRECT_LOG = 20;   // largest per-pass twiddle table kept as a direct table (2^20 x 32 B = 32 MB, L2 resident)
#ifndef FFT_MIN_BLOCKS
#define FFT_MIN_BLOCKS 3
#endif

struct FftPass {
    const uint32_t *in;
    uint32_t *out;
    uint32_t log_n, r, log_ns, log_te;    // tile has 2^log_te elements = 2^(log_te - r) columns x 2^r
    uint32_t lb;                          // bits of the low twiddle table
    const Fr *tlo, *thi, *twr;
    const Fr *tdir;                       // direct table omega_(Ns R)^e, e < Ns R (null: use the two-level table)
    uint32_t scale[8];                    // LAST, inverse transform: n^-1 in Montgomery form
    uint32_t do_scale;
    unsigned long long *err;
};

__device__ __forceinline__ uint32_t fft_phys(uint32_t slot) { return slot + (slot >> 5); }

template <bool FIRST, bool LAST> __global__ void __launch_bounds__(FFT_BLOCK, FFT_MIN_BLOCKS) k_fft_pass(FftPass p) {
    extern __shared__ __align__(16) uint32_t sm[];
    const uint32_t te = 1u << p.log_te, plane = te + (te >> 5) + 1;
    uint32_t *u = sm;                      // 8 planes of `plane` words
    uint32_t *tw = sm + 8 * plane;         // 8 planes of 128 words: omega_256^t
    const uint32_t tid = threadIdx.x, r = p.r, R = 1u << r, log_j = p.log_te - r, J = 1u << log_j;
    for (uint32_t i = tid; i < 128 * 8; i += FFT_BLOCK) tw[(i & 7) * 128 + (i >> 3)] = p.twr[i >> 3].l[i & 7];
    const size_t jbase = (size_t)blockIdx.x << log_j;
    const uint32_t ns_mask = (1u << p.log_ns) - 1u;
    // ---- load (+ decode / twiddle) ----
    for (uint32_t e = tid; e < te; e += FFT_BLOCK) {
        const uint32_t jj = e & (J - 1), t = e >> log_j;
        const size_t j = jbase + jj, idx = j + ((size_t)t << (p.log_n - r));
        const uint4 *src = reinterpret_cast<const uint4 *>(p.in + idx * 8);
        uint4 a = __ldg(src), b = __ldg(src + 1);
        Fr v;
        if (FIRST) {
            uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            v = limbs_from_be_words<FrP>(w);
            if (!is_canonical(v)) atomicMin(p.err, (unsigned long long)(((uint64_t)idx << 8) | (P2B_EARG << 4)));
            // No conversion to Montgomery form: the canonical words are read AS a Montgomery representation, i.e. of the value
            // a / R.  The transform is linear and every twiddle is a true Montgomery constant, so the words after the last
            // pass represent DFT(a) / R -- which are the canonical words of DFT(a).  Saves two multiplications per element.
        } else {
            v.l[0] = a.x; v.l[1] = a.y; v.l[2] = a.z; v.l[3] = a.w; v.l[4] = b.x; v.l[5] = b.y; v.l[6] = b.z; v.l[7] = b.w;
            if (p.log_ns) {
                const uint32_t k = (uint32_t)j & ns_mask;
                const uint32_t ex = (t * k) << (p.log_n - p.log_ns - r);
                // omega_(Ns R)^(t k): straight from the pass's own table when it is small enough to stay in L2 (one
                // multiplication per element), else combined from the two-level table of omega_n (two multiplications)
                Fr w = p.tdir ? p.tdir[t * k] : mul(p.tlo[ex & ((1u << p.lb) - 1u)], p.thi[ex >> p.lb]);
                v = mul(v, w);
            }
        }
        const uint32_t ph = fft_phys(t * J + jj);
#pragma unroll
        for (int w = 0; w < 8; w++) u[w * plane + ph] = v.l[w];
    }
    __syncthreads();
    // ---- r decimation-in-frequency stages: u[bitrev(q)] = V[q] ----
    // The last FFT_REG_STAGES stages (butterfly spans 4, 2, 1) run in registers, one radix-8 (radix-4) group per thread: their
    // twiddles are omega_8^p only, and the trivial ones (p = 0) are known at compile time: 5 multiplications per 8 elements
    // instead of 8 (1 per 4 instead of 2), and two (one) shared-memory round trips less.
    const uint32_t smem_stages = r >= FFT_REG_STAGES ? r - FFT_REG_STAGES : r;
    for (uint32_t s = 0; s < smem_stages; s++) {
        const uint32_t lh = r - 1 - s, half = 1u << lh;
        for (uint32_t bi = tid; bi < (te >> 1); bi += FFT_BLOCK) {
            const uint32_t jj = bi & (J - 1), b = bi >> log_j;
            const uint32_t pos = b & (half - 1), i0 = ((b >> lh) << (lh + 1)) + pos, i1 = i0 + half;
            const uint32_t p0 = fft_phys(i0 * J + jj), p1 = fft_phys(i1 * J + jj);
            Fr x, y;
#pragma unroll
            for (int w = 0; w < 8; w++) { x.l[w] = u[w * plane + p0]; y.l[w] = u[w * plane + p1]; }
            Fr sum = add(x, y), d = sub(x, y);
            if (half > 1) {
                const uint32_t ti = (pos << s) << (FFT_RMAX - r);
                Fr wv;
#pragma unroll
                for (int w = 0; w < 8; w++) wv.l[w] = tw[w * 128 + ti];
                d = mul(d, wv);
            }
#pragma unroll
            for (int w = 0; w < 8; w++) { u[w * plane + p0] = sum.l[w]; u[w * plane + p1] = d.l[w]; }
        }
        __syncthreads();
    }
    if (r >= FFT_REG_STAGES) {
        constexpr int RG = 1 << FFT_REG_STAGES;                        // 8 (or 4) elements per thread
        Fr w8[3];                                                      // omega_8^1, ^2, ^3 = omega_256^32, ^64, ^96
#pragma unroll
        for (int k = 0; k < 3; k++)
#pragma unroll
            for (int w = 0; w < 8; w++) w8[k].l[w] = tw[w * 128 + 32 * (k + 1)];
        const uint32_t log_g = r - FFT_REG_STAGES;                     // register groups per column
        for (uint32_t gi = tid; gi < (te >> FFT_REG_STAGES); gi += FFT_BLOCK) {
            const uint32_t g = gi & ((1u << log_g) - 1u), jj = gi >> log_g;
            Fr x[RG];
#pragma unroll
            for (int k = 0; k < RG; k++) {
                const uint32_t ph = fft_phys((g * RG + k) * J + jj);
#pragma unroll
                for (int w = 0; w < 8; w++) x[k].l[w] = u[w * plane + ph];
            }
            if (RG == 8) {
#pragma unroll
                for (int q = 0; q < 4; q++) {                          // span 4: twiddle omega_8^q
                    Fr sum = add(x[q], x[q + 4]), d = sub(x[q], x[q + 4]);
                    x[q] = sum;
                    x[q + 4] = q ? mul(d, w8[q - 1]) : d;
                }
            }
#pragma unroll
            for (int base = 0; base < RG; base += 4)
#pragma unroll
                for (int q = 0; q < 2; q++) {                          // span 2: twiddle omega_4^q
                    Fr sum = add(x[base + q], x[base + q + 2]), d = sub(x[base + q], x[base + q + 2]);
                    x[base + q] = sum;
                    x[base + q + 2] = q ? mul(d, w8[1]) : d;
                }
#pragma unroll
            for (int base = 0; base < RG; base += 2) {                 // span 1: no twiddle
                Fr sum = add(x[base], x[base + 1]), d = sub(x[base], x[base + 1]);
                x[base] = sum;
                x[base + 1] = d;
            }
#pragma unroll
            for (int k = 0; k < RG; k++) {
                const uint32_t ph = fft_phys((g * RG + k) * J + jj);
#pragma unroll
                for (int w = 0; w < 8; w++) u[w * plane + ph] = x[k].l[w];
            }
        }
        __syncthreads();
    }
    // ---- autosort scatter (+ encode) ----
    Fr scale;
#pragma unroll
    for (int w = 0; w < 8; w++) scale.l[w] = p.scale[w];
    for (uint32_t e = tid; e < te; e += FFT_BLOCK) {
        uint32_t jj, q;
        if (p.log_ns == 0) { q = e & (R - 1); jj = e >> r; }          // first pass: runs of R consecutive outputs
        else { jj = e & (J - 1); q = e >> log_j; }                    // later passes: runs of J consecutive outputs
        const uint32_t qr = r ? (__brev(q) >> (32 - r)) : 0u;
        const uint32_t ph = fft_phys(qr * J + jj);
        Fr v;
#pragma unroll
        for (int w = 0; w < 8; w++) v.l[w] = u[w * plane + ph];
        const size_t j = jbase + jj, k = j & ns_mask;
        const size_t idx = (((j >> p.log_ns) << (p.log_ns + r)) | k) + ((size_t)q << p.log_ns);
        uint32_t o[8];
        if (LAST) {
            if (p.do_scale) v = mul(v, scale);                        // inverse transform: x n^-1 (Montgomery constant)
            limbs_to_be_words(v, o);
        } else {
#pragma unroll
            for (int w = 0; w < 8; w++) o[w] = v.l[w];
        }
        uint4 *dst = reinterpret_cast<uint4 *>(p.out + idx * 8);
        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
        dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
    }
}

// tables: tlo[i] = base^i (i < 2^lb), thi[i] = base^(i << lb) (i < 2^hb), optional twr[i] = root256^i (i < 128).
// hi_raw: store thi as canonical limbs instead of Montgomery (so that mont_mul(tlo, thi) is canonical).
static __global__ void k_fft_tables(Fr *tlo, Fr *thi, Fr *twr, Fr base, Fr root256, uint32_t lb, uint32_t hb, int hi_raw) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, nlo = 1u << lb, nhi = 1u << hb;
    if (i < nlo) tlo[i] = pow_u64(base, i);
    else if (i < nlo + nhi) {
        Fr v = pow_u64(base, (uint64_t)(i - nlo) << lb);
        thi[i - nlo] = hi_raw ? from_mont(v) : v;
    } else if (twr && i < nlo + nhi + 128) twr[i - nlo - nhi] = pow_u64(root256, i - nlo - nhi);
}

// dst[e] = base^(e << shift) from the two-level table (thi may be the copy that carries 1/n)
static __global__ void k_fft_direct_table(Fr *dst, const Fr *tlo, const Fr *thi, uint32_t lb, uint32_t shift, uint32_t count) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= count) return;
    const uint32_t ex = e << shift;
    dst[e] = mul(tlo[ex & ((1u << lb) - 1u)], thi[ex >> lb]);
}
static __global__ void k_fft_scale_table(Fr *dst, const Fr *src, uint32_t n, Fr k) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = mul(src[i], k);
}

// data[i] <- data[i] * g^i on the wire form (distribute_powers, domain.rs:176-189)
static __global__ void __launch_bounds__(256) k_fr_distribute_powers(uint32_t *data, size_t n, const Fr *glo, const Fr *ghi_raw,
                                                                      uint32_t lb, unsigned long long *err) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint4 *ptr = reinterpret_cast<uint4 *>(data + i * 8);
        uint4 a = ptr[0], b = ptr[1];
        uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        Fr v = limbs_from_be_words<FrP>(w);
        if (!is_canonical(v)) atomicMin(err, (unsigned long long)(((uint64_t)i << 8) | (P2B_EARG << 4)));
        Fr s = mul(glo[i & ((1u << lb) - 1u)], ghi_raw[i >> lb]);     // canonical g^i
        v = mul(to_mont(v), s);                                       // canonical v * g^i
        limbs_to_be_words(v, w);
        ptr[0] = make_uint4(w[0], w[1], w[2], w[3]);
        ptr[1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
}

// ------------------------------------------------------------------------------------------------- host side
static Fr host_fr_from_u64(uint64_t v) {
    Fr a = fp_zero<FrP>();
    a.l[0] = (uint32_t)v; a.l[1] = (uint32_t)(v >> 32);
    return to_mont(a);
}
// 2^28-th root of unity 7^((r-1)/2^28) (fr.rs:3-6 PrimeFieldGenerator = 7, S = 28), Montgomery form
static Fr host_root_of_unity() {
    uint32_t e[8];
    for (int i = 0; i < 8; i++) e[i] = FrP::p(i);
    e[0] -= 1;                                           // r - 1 (no borrow: r is odd)
    for (int i = 0; i < 8; i++) {                        // >> 28
        uint64_t lo = e[i], hi = i + 1 < 8 ? e[i + 1] : 0;
        e[i] = (uint32_t)(((hi << 32) | lo) >> 28);
    }
    return pow_limbs(host_fr_from_u64(7), e);
}

// omega_(2^log_n) (or its inverse) in Montgomery form and the canonical limbs of n^-1 (1 for the forward transform)
void host_domain_constants(uint32_t log_n, int inverse, uint32_t omega_mont[8], uint32_t ninv_canon[8]) {
    Fr omega = host_root_of_unity();
    for (uint32_t i = log_n; i < 28; i++) omega = sqr(omega);
    Fr scale = fp_zero<FrP>();
    scale.l[0] = 1;
    if (inverse) {
        omega = inv(omega);
        scale = from_mont(inv(host_fr_from_u64((uint64_t)1 << log_n)));
    }
    memcpy(omega_mont, omega.l, 32);
    memcpy(ninv_canon, scale.l, 32);
}

struct FftPlan {
    uint32_t npass, r[4], log_te;
};
static FftPlan fft_plan(uint32_t log_n) {
    FftPlan pl;
    pl.npass = log_n ? (log_n + FFT_RMAX - 1) / FFT_RMAX : 1;
    uint32_t left = log_n;
    for (uint32_t i = 0; i < pl.npass; i++) {
        pl.r[i] = (left + (pl.npass - i) - 1) / (pl.npass - i);
        left -= pl.r[i];
    }
    pl.log_te = log_n < (uint32_t)FFT_TILE_LOG ? log_n : FFT_TILE_LOG;
    return pl;
}

static int fft_tables(Ctx *c, uint32_t log_n, int inverse) {
    // one table set per direction: alternating fft / ifft calls (EvaluationDomain round trips, the prover's pattern) keep both
    DevBuf &tw = c->fft_tw_dir[inverse ? 1 : 0];
    if (tw.p && c->fft_tw_dir_log_n[inverse ? 1 : 0] == log_n) return P2B_OK;
    const uint32_t lb = (log_n + 1) / 2, hb = log_n - lb;
    const size_t nent = ((size_t)1 << lb) + ((size_t)1 << hb) + 128;
    // [omega tables | coset tables | thi x n^-1 | direct tables of the passes whose omega_(Ns R) table has <= 2^FFT_DIRECT_LOG entries]
    const FftPlan plan = fft_plan(log_n);
    size_t ndirect = 0;
    for (uint32_t i = 1, lns = plan.r[0]; i < plan.npass; lns += plan.r[i], i++)
        if (lns + plan.r[i] <= FFT_DIRECT_LOG) ndirect += (size_t)1 << (lns + plan.r[i]);
    int rc = dev_reserve(c, tw, (2 * nent + ((size_t)1 << hb) + ndirect) * sizeof(Fr));
    if (rc) return rc;
    Fr root = host_root_of_unity(), omega = root, root256 = root;
    for (uint32_t i = log_n; i < 28; i++) omega = sqr(omega);
    for (uint32_t i = FFT_RMAX; i < 28; i++) root256 = sqr(root256);
    Fr g = host_fr_from_u64(7);
    if (inverse) { omega = inv(omega); root256 = inv(root256); g = inv(g); }
    Fr *tlo = (Fr *)tw.p, *thi = tlo + ((size_t)1 << lb), *twr = thi + ((size_t)1 << hb);
    Fr *glo = twr + 128, *ghi = glo + ((size_t)1 << lb);
    const uint32_t threads = (uint32_t)nent;
    k_fft_tables<<<(threads + 127) / 128, 128, 0, c->stream>>>(tlo, thi, twr, omega, root256, lb, hb, 0);
    k_fft_tables<<<(threads + 127) / 128, 128, 0, c->stream>>>(glo, ghi, nullptr, g, g, lb, hb, 1);
    // inverse transform: the 1/n is folded into the high twiddle table the LAST pass uses (no separate multiplication)
    Fr ninv = inverse ? inv(host_fr_from_u64((uint64_t)1 << log_n)) : fp_one<FrP>();
    k_fft_scale_table<<<(int)((((size_t)1 << hb) + 127) / 128), 128, 0, c->stream>>>((Fr *)tw.p + 2 * nent, thi, (uint32_t)1 << hb, ninv);
    c->launches += 3;
    {
        Fr *dir = (Fr *)tw.p + 2 * nent + ((size_t)1 << hb);
        for (uint32_t i = 1, lns = plan.r[0]; i < plan.npass; lns += plan.r[i], i++) {
            const uint32_t bits = lns + plan.r[i];
            if (bits > FFT_DIRECT_LOG) continue;
            const bool last = i + 1 == plan.npass;
            k_fft_direct_table<<<(int)((((size_t)1 << bits) + 127) / 128), 128, 0, c->stream>>>(
                dir, tlo, last ? (const Fr *)((Fr *)tw.p + 2 * nent) : (const Fr *)thi, lb, log_n - bits, 1u << bits);
            c->launches++;
            dir += (size_t)1 << bits;
        }
    }
    P2B_CUDA(c, cudaGetLastError());
    c->fft_tw_dir_log_n[inverse ? 1 : 0] = log_n;
    return P2B_OK;
}

template <bool FIRST, bool LAST> static int fft_launch_pass(Ctx *c, const FftPass &p, uint32_t blocks, size_t smem) {
    static bool attr_done[64] = {};          // function attributes are per device
    bool &attr = attr_done[c->device & 63];
    if (!attr) {
        P2B_CUDA(c, cudaFuncSetAttribute(k_fft_pass<FIRST, LAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
        attr = true;
    }
    prof_begin(c, P2B_PROF_FFT_PASS);
    k_fft_pass<FIRST, LAST><<<blocks, FFT_BLOCK, smem, c->stream>>>(p);
    prof_end(c, P2B_PROF_FFT_PASS, 1);
    c->launches++;
    return P2B_OK;
}

// d_data: 2^log_n wire scalars; d_tmp: scratch of the same size.  The result lands in *d_result (d_data or d_tmp).
// With a second scratch buffer a three-pass transform runs data -> tmp -> tmp2 -> data and needs no copy back.
static int fft_run(Ctx *c, void *d_data, void *d_tmp, uint32_t log_n, int inverse, int coset, void **d_result, void *d_tmp2 = nullptr) {
    if (log_n > 28) return ctx_fail(c, P2B_EARG, "fft: log_n must be <= 28 (Fr::S, domain.rs:64-78)");
    int rc = fft_tables(c, log_n, inverse);
    if (rc) return rc;
    DevBuf &tw = c->fft_tw_dir[inverse ? 1 : 0];
    const size_t n = (size_t)1 << log_n;
    const uint32_t lb = (log_n + 1) / 2, hb = log_n - lb;
    Fr *tlo = (Fr *)tw.p, *thi = tlo + ((size_t)1 << lb), *twr = thi + ((size_t)1 << hb);
    Fr *glo = twr + 128, *ghi = glo + ((size_t)1 << lb);
    const Fr *thi_scaled = (const Fr *)tw.p + 2 * (((size_t)1 << lb) + ((size_t)1 << hb) + 128);
    int sgrid = (int)((n + 255) / 256);
    if (sgrid > c->sm_count * 8) sgrid = c->sm_count * 8;
    if (coset && !inverse) {   // coset_fft: distribute_powers(g) then fft (domain.rs:191-195)
        k_fr_distribute_powers<<<sgrid, 256, 0, c->stream>>>((uint32_t *)d_data, n, glo, ghi, lb, c->d_err);
        c->launches++;
    }
    FftPlan pl = fft_plan(log_n);
    FftPass p;
    memset(&p, 0, sizeof p);
    p.log_n = log_n; p.log_te = pl.log_te; p.lb = lb; p.tlo = tlo; p.thi = thi; p.twr = twr; p.err = c->d_err;
    Fr scale = fp_one<FrP>();
    if (inverse) scale = inv(host_fr_from_u64((uint64_t)n));               // n^-1 (domain.rs:163-173)
    memcpy(p.scale, scale.l, 32);
    p.do_scale = inverse ? 1u : 0u;
    const uint32_t te = 1u << pl.log_te;
    const size_t smem = (size_t)(8 * (te + (te >> 5) + 1) + 8 * 128) * 4;
    const uint32_t blocks = (uint32_t)(n >> pl.log_te);
    const Fr *dir_next = thi_scaled + ((size_t)1 << hb);             // direct tables, in pass order (fft_tables)
    void *src = d_data, *dst = d_tmp;
    const bool via_tmp2 = d_tmp2 && pl.npass == 3;
    uint32_t log_ns = 0;
    for (uint32_t i = 0; i < pl.npass; i++) {
        if (via_tmp2) dst = i == 0 ? d_tmp : i == 1 ? d_tmp2 : d_data;
        p.in = (const uint32_t *)src; p.out = (uint32_t *)dst; p.r = pl.r[i]; p.log_ns = log_ns;
        const bool first = i == 0, last = i + 1 == pl.npass;
        if (last && inverse && !first) {                               // the twiddle at load already carries 1/n
            p.thi = thi_scaled;
            p.do_scale = 0;
        }
        p.tdir = nullptr;
        if (!first && log_ns + pl.r[i] <= FFT_DIRECT_LOG) {
            p.tdir = dir_next;
            dir_next += (size_t)1 << (log_ns + pl.r[i]);
        }
        if (first && last) rc = fft_launch_pass<true, true>(c, p, blocks, smem);
        else if (first) rc = fft_launch_pass<true, false>(c, p, blocks, smem);
        else if (last) rc = fft_launch_pass<false, true>(c, p, blocks, smem);
        else rc = fft_launch_pass<false, false>(c, p, blocks, smem);
        if (rc) return rc;
        log_ns += pl.r[i];
        void *t = src; src = dst; dst = t;                             // (dst is overridden above when via_tmp2)
    }
    if (coset && inverse) {    // icoset_fft: ifft then distribute_powers(g^-1) (domain.rs:197-205)
        k_fr_distribute_powers<<<sgrid, 256, 0, c->stream>>>((uint32_t *)src, n, glo, ghi, lb, c->d_err);
        c->launches++;
    }
    P2B_CUDA(c, cudaGetLastError());
    *d_result = src;
    return P2B_OK;
}

int launch_fr_fft(Ctx *c, void *d_data, uint32_t log_n, int inverse, int coset) {
    const size_t bytes = (size_t)32 << log_n;
    const bool three = fft_plan(log_n).npass == 3;
    int rc = dev_reserve(c, c->misc, three ? 2 * bytes : bytes);
    if (rc) return rc;
    void *res = nullptr;
    if ((rc = fft_run(c, d_data, c->misc.p, log_n, inverse, coset, &res, three ? (char *)c->misc.p + bytes : nullptr))) return rc;
    if (res != d_data) P2B_CUDA(c, cudaMemcpyAsync(d_data, res, bytes, cudaMemcpyDeviceToDevice, c->stream));
    return P2B_OK;
}

static int fft_begin(Ctx *c) {
    P2B_CUDA(c, cudaSetDevice(c->device));
    c->last_error.clear();
    P2B_CUDA(c, cudaMemsetAsync(c->d_err, 0xff, sizeof(unsigned long long), c->stream));
    return P2B_OK;
}
static int fft_host(Ctx *c, uint8_t *data, uint32_t log_n, int inverse, int coset) {
    if (!data) return ctx_fail(c, P2B_EARG, "null buffer");
    if (log_n > 28) return ctx_fail(c, P2B_EARG, "fft: log_n must be <= 28 (Fr::S, domain.rs:64-78)");
    int rc = fft_begin(c);
    if (rc) return rc;
    const size_t bytes = (size_t)32 << log_n;
    if ((rc = dev_reserve(c, c->stage_in[0], bytes))) return rc;
    if ((rc = dev_reserve(c, c->misc, bytes))) return rc;
    if ((rc = io_h2d(c, c->stage_in[0].p, data, bytes, c->stream))) return rc;
    void *res = nullptr;
    if ((rc = fft_run(c, c->stage_in[0].p, c->misc.p, log_n, inverse, coset, &res))) return rc;
    if ((rc = io_d2h(c, data, res, bytes, c->stream))) return rc;
    return ctx_collect_error(c);
}
static int fft_dev(Ctx *c, void *d_data, uint32_t log_n, int inverse, int coset) {
    if (!d_data) return ctx_fail(c, P2B_EARG, "null buffer");
    P2B_CUDA(c, cudaSetDevice(c->device));
    return launch_fr_fft(c, d_data, log_n, inverse, coset);
}

}  // namespace p2b

using namespace p2b;
extern "C" {
int p2b_fr_fft(p2b_ctx *h, uint8_t *data, uint32_t log_n, int inverse, int coset) { P2B_RANGE("p2b_fr_fft");
    return h ? fft_host(&h->c, data, log_n, inverse, coset) : P2B_EARG;
}
int p2b_fr_fft_dev(p2b_ctx *h, void *d_data, uint32_t log_n, int inverse, int coset) { P2B_RANGE("p2b_fr_fft_dev");
    return h ? fft_dev(&h->c, d_data, log_n, inverse, coset) : P2B_EARG;
}
}
