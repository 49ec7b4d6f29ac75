// msm_impl.cuh -- K4/K5: Pippenger multi-scalar multiplication sum_i k_i * P_i over BN254 G1 / G2 on sm_100a.
//
// GPU replacement for bellman's multiexp / dense_multiexp (bellman/src/multiexp.rs:53-157,330-475; copy in
// powersoftau/src/utils.rs:190-292).  The result is one group element, so the window width, the bucket layout
// and the reduction order are free (the reference's `c = ceil(ln n)` is a CPU cache heuristic); only the affine
// encoding of the sum is observable.
//
//   k_msm_prepare     wire points -> affine Montgomery (AoS, 64/128 B), once per call
//   k_msm_hist        signed c-bit digits of every scalar; histogram of (window, |digit|) with L2 atomics
//   k_msm_scan_*      exclusive prefix sum of the histogram -> bucket offsets (counting sort, pass 2)
//   k_msm_scatter     digits recomputed; (point index, sign) written to its bucket's slot (counting sort, pass 3)
//   k_msm_accumulate  one thread per bucket: gathers its points (coalescing is impossible by construction; every
//                     gather is a full 64/128 B point), mixed adds in XYZZ coordinates (8M + 2S)
//   k_msm_reduce1     per window, 32-bucket strips: running-sum trick -> (sum, weighted sum) per strip
//   k_msm_reduce2     per window, one block: warp-shuffle suffix scans + shuffle tree reductions combine the strips
//   k_msm_final       Horner over the windows, one inversion, affine wire encoding
//
// Signed digits halve the bucket count: digit d in (-2^(c-1), 2^(c-1)], bucket |d|, sign folded into the point's y.
#pragma once
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include "codec.cuh"
#include "xyzz.cuh"
#include "msm_g2x2.cuh"
#include "p2b_internal.h"

#ifndef P2B_ACC_DEFAULT_VARIANT
#define P2B_ACC_DEFAULT_VARIANT 4
#endif

namespace p2b {

// NOTE: the reduction kernels call the inlined adds from as few call sites as possible (every site is ~3k SASS
// instructions).  An earlier version routed them through __noinline__ wrappers taking pointers to locals; combined
// with the warp shuffles around the calls that produced wrong sums on sm_100a (see tools/reduce2_test.cu), so the
// adds are inlined and the kernels are structured as loops over "items" instead.
template <class F> __device__ __forceinline__ Xyzz<F> xadd(const Xyzz<F> &p, const Xyzz<F> &q) { return xyzz_add(p, q); }
template <class F> __device__ __forceinline__ Xyzz<F> xdbl(const Xyzz<F> &p) { return xyzz_dbl(p); }

// ------------------------------------------------------------------------------------------------- memory helpers
template <int NW> __device__ __forceinline__ void ldw(uint32_t *dst, const uint32_t *src) {
    const uint4 *s = reinterpret_cast<const uint4 *>(src);
#pragma unroll
    for (int i = 0; i < NW / 4; i++) { uint4 v = __ldg(s + i); dst[4 * i] = v.x; dst[4 * i + 1] = v.y; dst[4 * i + 2] = v.z; dst[4 * i + 3] = v.w; }
}
template <int NW> __device__ __forceinline__ void stw(uint32_t *dst, const uint32_t *src) {
    uint4 *d = reinterpret_cast<uint4 *>(dst);
#pragma unroll
    for (int i = 0; i < NW / 4; i++) d[i] = make_uint4(src[4 * i], src[4 * i + 1], src[4 * i + 2], src[4 * i + 3]);
}
template <class F> __device__ __forceinline__ void store_xyzz(uint32_t *base, size_t i, const Xyzz<F> &p) {
    constexpr int W = FieldTraits<F>::WORDS;
    uint32_t w[4 * W];
#pragma unroll
    for (int j = 0; j < W; j++) { w[j] = get_word(p.x, j); w[W + j] = get_word(p.y, j); w[2 * W + j] = get_word(p.zz, j); w[3 * W + j] = get_word(p.zzz, j); }
    stw<4 * W>(base + i * 4 * W, w);
}
template <class F> __device__ __forceinline__ Xyzz<F> load_xyzz(const uint32_t *base, size_t i) {
    constexpr int W = FieldTraits<F>::WORDS;
    uint32_t w[4 * W];
    const uint4 *s = reinterpret_cast<const uint4 *>(base + i * 4 * W);
#pragma unroll
    for (int j = 0; j < W; j++) { uint4 v = s[j]; w[4 * j] = v.x; w[4 * j + 1] = v.y; w[4 * j + 2] = v.z; w[4 * j + 3] = v.w; }
    Xyzz<F> p;
#pragma unroll
    for (int j = 0; j < W; j++) { set_word(p.x, j, w[j]); set_word(p.y, j, w[W + j]); set_word(p.zz, j, w[2 * W + j]); set_word(p.zzz, j, w[3 * W + j]); }
    return p;
}
template <class F> __device__ __forceinline__ Xyzz<F> shfl_xyzz(const Xyzz<F> &p, int src_lane) {
    constexpr int W = FieldTraits<F>::WORDS;
    Xyzz<F> r;
#pragma unroll
    for (int j = 0; j < W; j++) {
        set_word(r.x, j, __shfl_sync(0xffffffffu, get_word(p.x, j), src_lane));
        set_word(r.y, j, __shfl_sync(0xffffffffu, get_word(p.y, j), src_lane));
        set_word(r.zz, j, __shfl_sync(0xffffffffu, get_word(p.zz, j), src_lane));
        set_word(r.zzz, j, __shfl_sync(0xffffffffu, get_word(p.zzz, j), src_lane));
    }
    return r;
}

// ------------------------------------------------------------------------------------------------- parameters
struct MsmGeom {
    uint32_t c;        // window bits of the wide windows; windows [0, nnarrow) are one bit narrower (balanced split of 255 bits)
    uint32_t nnarrow;
    uint32_t nwin;     // number of windows
    uint32_t nb;       // buckets per window = 2^(c-1); bucket ids 1..nb (slot 0 unused)
    uint32_t nbk;      // nb + 1
    uint32_t strip;    // buckets per reduce1 thread
    uint32_t tpw;      // reduce threads per window = nb / strip (multiple of 32, <= 1024)
};
__host__ __device__ __forceinline__ uint32_t win_width(const MsmGeom &g, uint32_t w) { return w < g.nnarrow ? g.c - 1 : g.c; }
__host__ __device__ __forceinline__ uint32_t win_off(const MsmGeom &g, uint32_t w) {
    return w < g.nnarrow ? w * (g.c - 1) : g.nnarrow * (g.c - 1) + (w - g.nnarrow) * g.c;
}
__device__ __forceinline__ void scalar_words(uint32_t k[8], const uint32_t *scalars, size_t i) {
    uint32_t w[8];
    ldw<8>(w, scalars + i * 8);
#pragma unroll
    for (int j = 0; j < 8; j++) k[j] = bswap32(w[7 - j]);
}
// raw c-bit window w of k (c <= 16)
__device__ __forceinline__ uint32_t raw_window(const uint32_t k[8], uint32_t bit, uint32_t c) {
    uint32_t word = bit >> 5, sh = bit & 31;
    uint64_t v = k[word];
    if (word + 1 < 8) v |= (uint64_t)k[word + 1] << 32;
    return (uint32_t)(v >> sh) & ((1u << c) - 1u);
}

static __device__ __forceinline__ void report_err(unsigned long long *err, uint64_t index, int kind, int sub) {
    atomicMin(err, (unsigned long long)((index << 8) | ((uint64_t)kind << 4) | (uint64_t)sub));
}

// ------------------------------------------------------------------------------------------------- kernels
// check bit 2 (subgroup probe): a finite point that is not on the curve raises *route (no error: such a batch takes the exact path)
template <class F> __global__ void __launch_bounds__(256) k_msm_prepare(const uint32_t *wire, uint32_t *aff, size_t n, unsigned long long *err, uint64_t err_base,
                                                                        int in_enc, int check, uint32_t *route = nullptr) {
    constexpr int WU = Wire<F>::WORDS_UNCOMPRESSED;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t w[WU];
        ldw<WU>(w, wire + i * WU);
        Aff<F> a;
        bool inf;
        int rc = point_decode<F>(a, inf, w, in_enc, (check & 1) != 0);     // in_enc: uncompressed wire, or RAW (already decoded)
        if (rc) { report_err(err, err_base + i, P2B_EDECODE, rc); inf = true; }
        else if (inf && (check & 2)) report_err(err, err_base + i, P2B_EINFINITY_IN, 0);
        if ((check & 4) && !rc && !inf && !on_curve(a)) atomicOr(route, 1u);
        uint32_t o[WU];
        point_encode<F>(o, a, inf, ENC_RAW_MONT_LE);   // infinity = all-zero: skipped by the accumulator
        stw<WU>(aff + i * WU, o);
    }
}

// Random scalars generated on the device: scalar i = 32 bytes of the ChaCha20 keystream (RFC 7539 block function, 20
// rounds) under `key`, block counter i / 2 in state words 12-13 (little-endian 64-bit), words 14-15 zero; the block's 64
// bytes are two 32-byte big-endian scalars, cleared above `bits` bits (bits <= 253 < log2 r: always canonical).  Used for
// the verifier's random linear combinations (powersoftau/src/utils.rs:118-124, phase2/src/utils.rs:76-85: `Fr::rand(rng)`
// per element): the key comes from the host's CSPRNG, the 2 GB of coefficients of a 2^26-element check never cross PCIe.
static __device__ __forceinline__ uint32_t rotl32(uint32_t v, int k) { return __funnelshift_l(v, v, k); }
#define P2B_QR(a, b, c, d) a += b; d ^= a; d = rotl32(d, 16); c += d; b ^= c; b = rotl32(b, 12); a += b; d ^= a; d = rotl32(d, 8); c += d; b ^= c; b = rotl32(b, 7);
struct ChaKey { uint32_t k[8]; };
static __global__ void __launch_bounds__(256) k_msm_random_scalars(uint32_t *scalars, size_t n, uint64_t first, ChaKey key, uint32_t bits) {
    const size_t nblk = (n + (first & 1) + 1) / 2;               // blocks touched: scalars first .. first + n - 1
    for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < nblk; b += (size_t)gridDim.x * blockDim.x) {
        const uint64_t ctr = first / 2 + b;
        uint32_t x[16], in[16];
        in[0] = 0x61707865u; in[1] = 0x3320646eu; in[2] = 0x79622d32u; in[3] = 0x6b206574u;
#pragma unroll
        for (int i = 0; i < 8; i++) in[4 + i] = key.k[i];
        in[12] = (uint32_t)ctr; in[13] = (uint32_t)(ctr >> 32); in[14] = 0; in[15] = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = in[i];
#pragma unroll 1
        for (int r = 0; r < 10; r++) {
            P2B_QR(x[0], x[4], x[8], x[12]) P2B_QR(x[1], x[5], x[9], x[13]) P2B_QR(x[2], x[6], x[10], x[14]) P2B_QR(x[3], x[7], x[11], x[15])
            P2B_QR(x[0], x[5], x[10], x[15]) P2B_QR(x[1], x[6], x[11], x[12]) P2B_QR(x[2], x[7], x[8], x[13]) P2B_QR(x[3], x[4], x[9], x[14])
        }
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] += in[i];
        // keystream bytes = little-endian words; scalar = those 32 bytes read as a big-endian integer, top bits cleared
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint64_t idx = 2 * ctr + h;
            if (idx < first || idx >= first + n) continue;
            uint32_t w[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                uint32_t be = bswap32(x[8 * h + j]);             // big-endian value of wire bytes 4j .. 4j+3
                const int hi_bit = 256 - 32 * j;                 // this word holds integer bits [hi_bit - 32, hi_bit)
                if ((int)bits <= hi_bit - 32) be = 0;
                else if ((int)bits < hi_bit) be &= (1u << (bits - (hi_bit - 32))) - 1u;
                w[j] = bswap32(be);                              // back to memory order
            }
            stw<8>(scalars + (size_t)(idx - first) * 8, w);
        }
    }
}
#undef P2B_QR

static __global__ void __launch_bounds__(256) k_msm_hist(const uint32_t *scalars, size_t n, MsmGeom g, uint32_t *hist, unsigned long long *err, uint64_t err_base,
                                                             uint32_t w_lo, uint32_t w_hi) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t k[8];
        scalar_words(k, scalars, i);
        Fr kc;
#pragma unroll
        for (int j = 0; j < 8; j++) kc.l[j] = k[j];
        if (w_lo == 0) {
            bool bad = !is_canonical(kc);
            const uint32_t total = g.nwin * g.c - g.nnarrow;       // bits covered by the windows (255 unless the caller bounded the scalars)
            if (total < 255) {                                     // scalars must stay below 2^(total - 1)
                for (uint32_t b = total - 1; b < 256; b++) bad |= ((k[b >> 5] >> (b & 31)) & 1u) != 0;
            }
            if (bad) report_err(err, err_base + i, P2B_EARG, 0);
        }
        uint32_t carry = 0;
        for (uint32_t w = 0; w < w_hi; w++) {                  // the signed-digit carry needs the windows below w_lo too
            const uint32_t wd = win_width(g, w);
            uint32_t d = raw_window(k, win_off(g, w), wd) + carry;
            carry = d > (1u << (wd - 1));
            uint32_t b = carry ? (1u << wd) - d : d;
            if (b && w >= w_lo) atomicAdd(&hist[w * g.nbk + b], 1u);
        }
    }
}

// Exclusive prefix sum of the histogram -> bucket offsets (also copied into `cursor`), offsets[count] = base + total.
// Three small kernels: per-tile sums (4096 entries per block), a serial scan of the <= 2k tile sums, per-tile rescan.
static constexpr uint32_t SCAN_TILE = 4096;     // 256 threads x 16 entries
static __device__ __forceinline__ uint32_t block_sum_256(uint32_t v, uint32_t *sm) {      // returns the block total to all
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    uint32_t t = 0;
    for (int w = 0; w < 8; w++) t += sm[w];
    __syncthreads();
    return t;
}
static __global__ void __launch_bounds__(256) k_msm_scan_tiles(const uint32_t *hist, uint32_t count, uint32_t *tile_sums) {
    __shared__ uint32_t sm[8];
    const uint32_t lo = blockIdx.x * SCAN_TILE + threadIdx.x * 16;
    uint32_t s = 0;
    for (uint32_t k = 0; k < 16; k++) s += lo + k < count ? hist[lo + k] : 0u;
    s = block_sum_256(s, sm);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = s;
}
static __global__ void k_msm_scan_tops(uint32_t *tile_sums, uint32_t ntiles, uint32_t base) {
    if (threadIdx.x || blockIdx.x) return;
    uint32_t run = base;
    for (uint32_t i = 0; i < ntiles; i++) { const uint32_t v = tile_sums[i]; tile_sums[i] = run; run += v; }
    tile_sums[ntiles] = run;
}
static __global__ void __launch_bounds__(256) k_msm_scan_apply(const uint32_t *hist, uint32_t *offsets, uint32_t *cursor, uint32_t count,
                                                               const uint32_t *tile_sums, uint32_t ntiles) {
    __shared__ uint32_t warp_tot[8];
    const uint32_t lo = blockIdx.x * SCAN_TILE + threadIdx.x * 16, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t v[16], s = 0;
    for (uint32_t k = 0; k < 16; k++) { v[k] = lo + k < count ? hist[lo + k] : 0u; s += v[k]; }
    uint32_t incl = s;                                          // inclusive scan of the thread sums within the warp
    for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if ((int)lane >= d) incl += o; }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    uint32_t run = tile_sums[blockIdx.x] + incl - s;
    for (uint32_t w = 0; w < wid; w++) run += warp_tot[w];
    for (uint32_t k = 0; k < 16; k++) {
        if (lo + k < count) { offsets[lo + k] = run; cursor[lo + k] = run; }
        run += v[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) offsets[count] = tile_sums[ntiles];
}

static __global__ void __launch_bounds__(256) k_msm_scatter(const uint32_t *scalars, size_t n, MsmGeom g, uint32_t *cursor, uint32_t *sorted,
                                                                uint32_t w_lo, uint32_t w_hi) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t k[8];
        scalar_words(k, scalars, i);
        uint32_t carry = 0;
        for (uint32_t w = 0; w < w_hi; w++) {
            const uint32_t wd = win_width(g, w);
            uint32_t d = raw_window(k, win_off(g, w), wd) + carry;
            carry = d > (1u << (wd - 1));
            uint32_t b = carry ? (1u << wd) - d : d;
            if (b && w >= w_lo) {
                uint32_t pos = atomicAdd(&cursor[w * g.nbk + b], 1u);
                sorted[pos] = ((uint32_t)i << 1) | carry;     // carry == 1 <=> negative digit
            }
        }
    }
}

// Buckets are handed to the accumulation threads in order of decreasing size: the 32 lanes of a warp then walk lists of
// (almost) equal length -- with one thread per bucket the warp otherwise waits for its longest list (Poisson spread: +6 %
// at 1024 terms per bucket, +17 % at 128) -- and the longest buckets start first.  Counting sort of the bucket sizes,
// capped at the per-thread segment length.
static __global__ void __launch_bounds__(256) k_msm_size_hist(const uint32_t *offsets, uint32_t slot_cnt, uint32_t cap, uint32_t *size_hist) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < slot_cnt; i += gridDim.x * blockDim.x) {
        const uint32_t sz = offsets[i + 1] - offsets[i];
        atomicAdd(&size_hist[sz < cap ? sz : cap], 1u);
    }
}
// single block: start position of every size class, largest first; also resets the histogram for the next group
static __global__ void __launch_bounds__(1024) k_msm_size_scan(uint32_t *size_hist, uint32_t *size_start, uint32_t nbins) {
    __shared__ uint32_t sums[1024];
    const uint32_t t = threadIdx.x, per = (nbins + 1023) / 1024;
    // thread t owns the descending range of bins [hi_bin - per*(t+1) + 1 .. hi_bin - per*t], hi_bin = nbins - 1
    uint32_t s = 0;
    for (uint32_t k = 0; k < per; k++) {
        const uint32_t r = t * per + k;                        // rank in descending order
        if (r < nbins) s += size_hist[nbins - 1 - r];
    }
    sums[t] = s;
    __syncthreads();
    for (uint32_t d = 1; d < 1024; d <<= 1) {
        uint32_t v = t >= d ? sums[t - d] : 0;
        __syncthreads();
        sums[t] += v;
        __syncthreads();
    }
    uint32_t run = t ? sums[t - 1] : 0;
    for (uint32_t k = 0; k < per; k++) {
        const uint32_t r = t * per + k;
        if (r < nbins) {
            const uint32_t b = nbins - 1 - r, cnt = size_hist[b];
            size_start[b] = run;
            size_hist[b] = 0;
            run += cnt;
        }
    }
}
static __global__ void __launch_bounds__(256) k_msm_size_scatter(const uint32_t *offsets, uint32_t slot_cnt, uint32_t cap, uint32_t *size_start,
                                                                 uint32_t *perm) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < slot_cnt; i += gridDim.x * blockDim.x) {
        const uint32_t sz = offsets[i + 1] - offsets[i];
        perm[atomicAdd(&size_start[sz < cap ? sz : cap], 1u)] = i;
    }
}

// The same two steps with block-private histograms in shared memory: with uniform scalars nearly all buckets of a group have one
// of ~100 sizes, so the global-atomic versions above serialise on a handful of counters (5 ms each per 2^26 MSM).  Here a block
// counts its 4096 slots in shared memory (each thread keeps the rank its atomicAdd returned), reserves a range per non-empty
// size class with ONE global atomicAdd, and writes its slots into those ranges.  nbins * 8 bytes of dynamic shared memory.
static constexpr uint32_t SIZE_TILE = 4096;      // slots per block iteration: 256 threads x 16
static __global__ void __launch_bounds__(256) k_msm_size_hist_smem(const uint32_t *offsets, uint32_t slot_cnt, uint32_t cap, uint32_t *size_hist,
                                                                   uint32_t nbins) {
    extern __shared__ uint32_t sm_bins[];
    for (uint32_t b = threadIdx.x; b < nbins; b += 256) sm_bins[b] = 0;
    __syncthreads();
    for (uint32_t base = blockIdx.x * SIZE_TILE; base < slot_cnt; base += gridDim.x * SIZE_TILE) {
        for (uint32_t k = 0; k < 16; k++) {
            const uint32_t i = base + k * 256 + threadIdx.x;
            if (i < slot_cnt) {
                const uint32_t sz = offsets[i + 1] - offsets[i];
                atomicAdd(&sm_bins[sz < cap ? sz : cap], 1u);
            }
        }
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < nbins; b += 256)
        if (sm_bins[b]) atomicAdd(&size_hist[b], sm_bins[b]);
}
static __global__ void __launch_bounds__(256) k_msm_size_scatter_smem(const uint32_t *offsets, uint32_t slot_cnt, uint32_t cap, uint32_t *size_start,
                                                                      uint32_t *perm, uint32_t nbins) {
    extern __shared__ uint32_t sm_bins[];           // [0, nbins): counts, then the reserved base of each class
    for (uint32_t base = blockIdx.x * SIZE_TILE; base < slot_cnt; base += gridDim.x * SIZE_TILE) {
        for (uint32_t b = threadIdx.x; b < nbins; b += 256) sm_bins[b] = 0;
        __syncthreads();
        uint32_t bin[16], rank[16];
        for (uint32_t k = 0; k < 16; k++) {
            const uint32_t i = base + k * 256 + threadIdx.x;
            bin[k] = 0xffffffffu;
            if (i < slot_cnt) {
                const uint32_t sz = offsets[i + 1] - offsets[i];
                bin[k] = sz < cap ? sz : cap;
                rank[k] = atomicAdd(&sm_bins[bin[k]], 1u);
            }
        }
        __syncthreads();
        for (uint32_t b = threadIdx.x; b < nbins; b += 256)
            if (sm_bins[b]) sm_bins[b] = atomicAdd(&size_start[b], sm_bins[b]);
        __syncthreads();
        for (uint32_t k = 0; k < 16; k++)
            if (bin[k] != 0xffffffffu) perm[sm_bins[bin[k]] + rank[k]] = base + k * 256 + threadIdx.x;
        __syncthreads();
    }
}

// Skewed inputs (many equal digits) would put most terms of a window into one bucket and serialise them on one thread.
// A bucket thread therefore takes at most `seg` entries; the rest is cut into items of `chunk` entries that whole blocks
// reduce (k_msm_heavy) and a last kernel folds into the bucket (k_msm_heavy_combine).  Uniform scalars never overflow
// (seg = 8 x the average load), so the fast path only pays two empty launches.
struct MsmHeavy {
    uint32_t *counters;   // [0] items pushed, [1] heavy buckets pushed
    uint4 *items;         // (bucket slot, first entry, end entry, -)
    uint4 *hbuckets;      // (bucket slot, first item, item count, -)
    uint32_t *partials;   // one XYZZ per item
    uint32_t seg, chunk, cap_items, cap_buckets;
};
template <class F> __device__ __forceinline__ Xyzz<F> madd_fused_if_g1(const Xyzz<F> &acc, const Aff<F> &q) {
    if constexpr (FieldTraits<F>::WORDS == 8) return xyzz_madd_fused(acc, q);
    else return xyzz_madd_sq(acc, q);
}
// A gathered point is 64 B (G1) / 128 B (G2) of a random 128-byte line: by default the miss fills the whole line (124 B of DRAM
// reads per 64-byte G1 gather measured).  The L2::64B prefetch-size qualifier asks for the touched half only.
#ifndef P2B_GATHER_L2_HINT
#define P2B_GATHER_L2_HINT 1
#endif
template <int NW> __device__ __forceinline__ void ldw_gather(uint32_t *dst, const uint32_t *src) {
#if P2B_GATHER_L2_HINT
    if constexpr (NW == 16) {
#pragma unroll
        for (int i = 0; i < NW / 4; i++)
            asm volatile("ld.global.nc.L2::64B.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(dst[4 * i]), "=r"(dst[4 * i + 1]), "=r"(dst[4 * i + 2]), "=r"(dst[4 * i + 3])
                         : "l"(src + 4 * i));
        return;
    }
#endif
    ldw<NW>(dst, src);
}
template <class F> __device__ __forceinline__ void accumulate_entry(Xyzz<F> &acc, const uint32_t *aff, uint32_t ent) {
    constexpr int WU = Wire<F>::WORDS_UNCOMPRESSED, W = FieldTraits<F>::WORDS;
    uint32_t w[WU];
    ldw_gather<WU>(w, aff + (size_t)(ent >> 1) * WU);
    uint32_t any = 0;
#pragma unroll
    for (int j = 0; j < WU; j++) any |= w[j];
    if (!any) return;                                          // point at infinity contributes nothing
    Aff<F> q;
#pragma unroll
    for (int j = 0; j < W; j++) { set_word(q.x, j, w[j]); set_word(q.y, j, w[W + j]); }
    q.y = cneg(q.y, (ent & 1u) != 0);
    acc = madd_fused_if_g1(acc, q);
}
// Bucket accumulation, one thread per bucket.  VARIANT (tuning, selected at run time by P2B_ACC_VARIANT; all give the same sums):
//   0  the round-1 loop: entry -> gather -> mixed add, nothing in flight across iterations
//   1  software pipeline: the point of entry e+1 and the index of entry e+2 are loaded while the mixed add of entry e runs
//      (each gather is a dependent pair of loads -- list entry, then a random 64 / 128 B point from HBM -- about 1.5 us of
//      latency that 3 warps per scheduler do not always cover)
//   2  variant 1 + the two squarings of the mixed add on the dedicated squaring routine
//   3  variant 1 at 4 blocks per SM (<= 128 registers)
//   4  variant 2 + Y3 as one fused two-product multiplication (mont_mul2)
//   5  the plain loop of variant 0 at 3 blocks per SM (<= 168 registers): a G2 tuning variant (P2B_ACC_VARIANT_G2)
template <class F, int VARIANT> struct AccBounds { static constexpr int MIN_BLOCKS = VARIANT == 3 ? 4 : (VARIANT == 0 ? 1 : 3); };
template <class F> __device__ __forceinline__ void load_point_words(uint32_t *w, const uint32_t *aff, uint32_t ent) {
    ldw_gather<Wire<F>::WORDS_UNCOMPRESSED>(w, aff + (size_t)(ent >> 1) * Wire<F>::WORDS_UNCOMPRESSED);
}
template <class F, int SQ> __device__ __forceinline__ void accumulate_words(Xyzz<F> &acc, const uint32_t *w, uint32_t ent) {
    constexpr int WU = Wire<F>::WORDS_UNCOMPRESSED, W = FieldTraits<F>::WORDS;
    uint32_t any = 0;
#pragma unroll
    for (int j = 0; j < WU; j++) any |= w[j];
    if (!any) return;                                          // point at infinity contributes nothing
    Aff<F> q;
#pragma unroll
    for (int j = 0; j < W; j++) { set_word(q.x, j, w[j]); set_word(q.y, j, w[W + j]); }
    q.y = cneg(q.y, (ent & 1u) != 0);
    acc = SQ == 2 ? madd_fused_if_g1(acc, q) : (SQ == 1 ? xyzz_madd_sq(acc, q) : xyzz_madd(acc, q));
}
template <class F, int VARIANT>
__global__ void __launch_bounds__(128, AccBounds<F, VARIANT>::MIN_BLOCKS) k_msm_accumulate(const uint32_t *aff, const uint32_t *offsets, const uint32_t *sorted,
                                                                           MsmGeom g, uint32_t *buckets, int first, uint32_t slot_lo,
                                                                           uint32_t slot_cnt, MsmHeavy hv, unsigned long long *err,
                                                                           const uint32_t *perm) {
    constexpr int WU = Wire<F>::WORDS_UNCOMPRESSED;
    // slots [slot_lo, slot_lo + slot_cnt) = the (window, bucket) pairs of one window group; `offsets` is that group's table;
    // perm lists the group's slots by decreasing size
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < slot_cnt; t += gridDim.x * blockDim.x) {
        const uint32_t idx = perm[t];
        const uint32_t gb = slot_lo + idx;
        const uint32_t lo = offsets[idx], hi = offsets[idx + 1];
        Xyzz<F> acc = xyzz_infinity<F>();
        if (!first) {                                          // later chunk of a streamed MSM: continue the bucket
            if (lo == hi) continue;
            acc = load_xyzz<F>(buckets, gb);
        }
        const uint32_t mine = hi - lo > hv.seg ? lo + hv.seg : hi;
        if constexpr (VARIANT == 0 || VARIANT == 5) {
#pragma unroll 1
            for (uint32_t e = lo; e < mine; e++) accumulate_entry<F>(acc, aff, __ldg(sorted + e));
        } else {
            if (lo < mine) {
                uint32_t ent = __ldg(sorted + lo), ent1 = lo + 1 < mine ? __ldg(sorted + lo + 1) : 0u;
                uint32_t w[WU];
                load_point_words<F>(w, aff, ent);
#pragma unroll 1
                for (uint32_t e = lo; e < mine; e++) {
                    uint32_t wn[WU];
                    const uint32_t ent2 = e + 2 < mine ? __ldg(sorted + e + 2) : 0u;
                    if (e + 1 < mine) load_point_words<F>(wn, aff, ent1);      // in flight during the mixed add below
                    accumulate_words<F, VARIANT == 4 ? 2 : (VARIANT == 2 ? 1 : 0)>(acc, w, ent);
#pragma unroll
                    for (int j = 0; j < WU; j++) w[j] = wn[j];
                    ent = ent1;
                    ent1 = ent2;
                }
            }
        }
        store_xyzz<F>(buckets, gb, acc);
        if (mine < hi) {                                       // overflow: hand the tail to k_msm_heavy
            const uint32_t nit = (hi - mine + hv.chunk - 1) / hv.chunk;
            const uint32_t fi = atomicAdd(&hv.counters[0], nit), hb = atomicAdd(&hv.counters[1], 1u);
            if (fi + nit <= hv.cap_items && hb < hv.cap_buckets) {
                for (uint32_t k = 0; k < nit; k++) {
                    const uint32_t a = mine + k * hv.chunk, b = hi - a > hv.chunk ? a + hv.chunk : hi;
                    hv.items[fi + k] = make_uint4(gb, a, b, 0u);
                }
                hv.hbuckets[hb] = make_uint4(gb, fi, nit, 0u);
            } else report_err(err, gb, P2B_ECUDA, 0);          // cannot happen: the capacities cover the worst case
        }
    }
}

// G2 with two lanes per bucket (msm_g2x2.cuh): lane pair (2j, 2j+1) of a block owns one bucket; each lane loads, keeps and
// stores its own component of every Fq2 coordinate.  SETS = 2 walks the list twice, once per bucket set of the pair mode.
#ifndef P2B_G2X2_MIN_BLOCKS
#define P2B_G2X2_MIN_BLOCKS 3
#endif
template <int SETS>
__global__ void __launch_bounds__(128, P2B_G2X2_MIN_BLOCKS) k_msm_accumulate_g2x2(const uint32_t *aff_a, const uint32_t *aff_b, const uint32_t *offsets,
                                                                                  const uint32_t *sorted, uint32_t *buckets_a, uint32_t *buckets_b,
                                                                                  int first, uint32_t slot_lo, uint32_t slot_cnt, MsmHeavy hv,
                                                                                  unsigned long long *err, const uint32_t *perm) {
#if defined(__CUDA_ARCH__)
    const int lane = threadIdx.x & 31, odd = lane & 1;
    const unsigned pm = 3u << (lane & ~1);
    const uint32_t npairs = gridDim.x * (blockDim.x >> 1);
    for (uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 1; t < slot_cnt; t += npairs) {
        const uint32_t idx = perm[t];
        const uint32_t gb = slot_lo + idx;
        const uint32_t lo = offsets[idx], hi = offsets[idx + 1];
        if (!first && lo == hi) continue;
        const uint32_t mine = SETS == 1 && hi - lo > hv.seg ? lo + hv.seg : hi;      // pair mode walks whole buckets (see k_msm_accumulate_pair)
#pragma unroll 1
        for (int set = 0; set < SETS; set++) {
            const uint32_t *aff = set ? aff_b : aff_a;
            uint32_t *bk = set ? buckets_b : buckets_a;
            XyzzHalf acc;
            acc.x = fp_zero<FqP>(); acc.y = fp_zero<FqP>(); acc.zz = fp_zero<FqP>(); acc.zzz = fp_zero<FqP>();
            if (!first) acc = load_xyzz_half(bk, gb, odd);
#pragma unroll 1
            for (uint32_t e = lo; e < mine; e++) {
                const uint32_t ent = __ldg(sorted + e);
                const uint4 *src = reinterpret_cast<const uint4 *>(aff + (size_t)(ent >> 1) * 32 + 8 * odd);   // x.c0 x.c1 y.c0 y.c1, 8 words each
                const uint4 x0 = __ldg(src), x1 = __ldg(src + 1), y0 = __ldg(src + 4), y1 = __ldg(src + 5);
                Fq qx, qy;
                qx.l[0] = x0.x; qx.l[1] = x0.y; qx.l[2] = x0.z; qx.l[3] = x0.w; qx.l[4] = x1.x; qx.l[5] = x1.y; qx.l[6] = x1.z; qx.l[7] = x1.w;
                qy.l[0] = y0.x; qy.l[1] = y0.y; qy.l[2] = y0.z; qy.l[3] = y0.w; qy.l[4] = y1.x; qy.l[5] = y1.y; qy.l[6] = y1.z; qy.l[7] = y1.w;
                if (!either(!(is_zero(qx) & is_zero(qy)), pm)) continue;            // all-zero = point at infinity: contributes nothing
                h2_madd(acc, qx, qy, (ent & 1u) != 0, odd != 0, pm);
            }
            store_xyzz_half(bk, gb, odd, acc);
        }
        if (SETS == 1 && mine < hi && !odd) {                 // overflow: hand the tail to k_msm_heavy (one lane of the pair)
            const uint32_t nit = (hi - mine + hv.chunk - 1) / hv.chunk;
            const uint32_t fi = atomicAdd(&hv.counters[0], nit), hb = atomicAdd(&hv.counters[1], 1u);
            if (fi + nit <= hv.cap_items && hb < hv.cap_buckets) {
                for (uint32_t k = 0; k < nit; k++) {
                    const uint32_t a = mine + k * hv.chunk, b = hi - a > hv.chunk ? a + hv.chunk : hi;
                    hv.items[fi + k] = make_uint4(gb, a, b, 0u);
                }
                hv.hbuckets[hb] = make_uint4(gb, fi, nit, 0u);
            } else report_err(err, gb, P2B_ECUDA, 0);
        }
    }
#endif
}

// Pair mode: ONE sorted list of (term, sign) entries drives TWO bucket sets -- sum k_i A_i and sum k_i B_i for two point
// arrays under the same scalars (merge_pairs, phase2/src/utils.rs:59-105; power_pairs = the same with B = A shifted by one
// point, powersoftau/src/utils.rs:133-135, where both points of an entry sit next to each other in memory).  The sort, the
// scalar traffic and the digit extraction are shared; every entry costs two mixed adds.  The whole bucket is walked by its
// thread (no overflow hand-off: the scalars of these checks are the verifier's own random numbers, never skewed; any other
// input is still summed correctly, only not load-balanced).
template <class F>
__global__ void __launch_bounds__(128, 2) k_msm_accumulate_pair(const uint32_t *aff_a, const uint32_t *aff_b, const uint32_t *offsets,
                                                                const uint32_t *sorted, uint32_t *buckets_a, uint32_t *buckets_b, int first,
                                                                uint32_t slot_lo, uint32_t slot_cnt, const uint32_t *perm) {
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < slot_cnt; t += gridDim.x * blockDim.x) {
        const uint32_t idx = perm[t];
        const uint32_t gb = slot_lo + idx;
        const uint32_t lo = offsets[idx], hi = offsets[idx + 1];
        if (!first && lo == hi) continue;
        if constexpr (FieldTraits<F>::WORDS == 8) {
            Xyzz<F> acc_a = xyzz_infinity<F>(), acc_b = xyzz_infinity<F>();
            if (!first) {
                acc_a = load_xyzz<F>(buckets_a, gb);
                acc_b = load_xyzz<F>(buckets_b, gb);
            }
#pragma unroll 1
            for (uint32_t e = lo; e < hi; e++) {
                const uint32_t ent = __ldg(sorted + e);
                accumulate_entry<F>(acc_a, aff_a, ent);
                accumulate_entry<F>(acc_b, aff_b, ent);
            }
            store_xyzz<F>(buckets_a, gb, acc_a);
            store_xyzz<F>(buckets_b, gb, acc_b);
        } else {
            // G2: one accumulator already takes the whole register file -- walk the (short, cached) list once per set
#pragma unroll 1
            for (int set = 0; set < 2; set++) {
                const uint32_t *aff = set ? aff_b : aff_a;
                uint32_t *bk = set ? buckets_b : buckets_a;
                Xyzz<F> acc = xyzz_infinity<F>();
                if (!first) acc = load_xyzz<F>(bk, gb);
#pragma unroll 1
                for (uint32_t e = lo; e < hi; e++) accumulate_entry<F>(acc, aff, __ldg(sorted + e));
                store_xyzz<F>(bk, gb, acc);
            }
        }
    }
}
// one block per item: 128 partial sums, then a tree over shared memory
template <class F> __global__ void __launch_bounds__(128) k_msm_heavy(const uint32_t *aff, const uint32_t *sorted, MsmHeavy hv) {
    constexpr int W = FieldTraits<F>::WORDS;
    __shared__ __align__(16) uint32_t sm[128 * 4 * W];
    const uint32_t count = hv.counters[0] < hv.cap_items ? hv.counters[0] : hv.cap_items;
    for (uint32_t it = blockIdx.x; it < count; it += gridDim.x) {
        const uint4 item = hv.items[it];
        Xyzz<F> acc = xyzz_infinity<F>();
#pragma unroll 1
        for (uint32_t e = item.y + threadIdx.x; e < item.z; e += 128) accumulate_entry<F>(acc, aff, __ldg(sorted + e));
        store_xyzz<F>(sm, threadIdx.x, acc);
        __syncthreads();
#pragma unroll 1
        for (uint32_t s = 64; s >= 1; s >>= 1) {
            if (threadIdx.x < s) {
                acc = xadd(acc, load_xyzz<F>(sm, threadIdx.x + s));
                store_xyzz<F>(sm, threadIdx.x, acc);
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) store_xyzz<F>(hv.partials, it, acc);
        __syncthreads();
    }
}
template <class F> __global__ void __launch_bounds__(128) k_msm_heavy_combine(uint32_t *buckets, MsmHeavy hv) {
    const uint32_t count = hv.counters[1] < hv.cap_buckets ? hv.counters[1] : hv.cap_buckets;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        const uint4 hb = hv.hbuckets[i];
        Xyzz<F> acc = load_xyzz<F>(buckets, hb.x);
#pragma unroll 1
        for (uint32_t k = 0; k < hb.z; k++) acc = xadd(acc, load_xyzz<F>(hv.partials, hb.y + k));
        store_xyzz<F>(buckets, hb.x, acc);
    }
}

// thread (w, t): buckets b = t*strip + 1 .. (t+1)*strip, walked from the top:  run = sum B_b ; acc = sum (b - t*strip) B_b
template <class F> __global__ void __launch_bounds__(128) k_msm_reduce1(const uint32_t *buckets, MsmGeom g, uint32_t *s1, uint32_t *s2) {
    const uint32_t total = g.nwin * g.tpw;
    uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= total) return;
    const uint32_t w = id / g.tpw, t = id % g.tpw;
    Xyzz<F> run = xyzz_infinity<F>(), acc = xyzz_infinity<F>();
#pragma unroll 1
    for (uint32_t j = g.strip; j >= 1; j--) {
        Xyzz<F> b = load_xyzz<F>(buckets, (size_t)w * g.nbk + t * g.strip + j);
        run = xadd(run, b);
        acc = xadd(acc, run);
    }
    store_xyzz<F>(s1, id, run);
    store_xyzz<F>(s2, id, acc);
}

// One warp-wide reduction of 32 XYZZ values, two flavours sharing the same two add sites:
//   plain    : sum_l v_l
//   weighted : sum_l l * v_l = sum_l (sum_{m > l} v_m): shift down one lane, inclusive Hillis-Steele suffix scan,
//              then the plain butterfly sum of the scan.
// Shuffles only -- no shared memory; every lane returns the total.
template <class F> __device__ __forceinline__ Xyzz<F> warp_reduce(Xyzz<F> v, bool weighted) {
    const int lane = threadIdx.x & 31;
    if (weighted) {                                            // warp-uniform
        Xyzz<F> o = shfl_xyzz(v, (lane + 1) & 31);
        v = select(lane == 31, xyzz_infinity<F>(), o);
#pragma unroll 1
        for (int d = 1; d < 32; d <<= 1) {
            o = shfl_xyzz(v, (lane + d) & 31);
            o = select(lane + d >= 32, xyzz_infinity<F>(), o);
            v = xadd(v, o);
        }
    }
#pragma unroll 1
    for (int d = 16; d >= 1; d >>= 1) {
        Xyzz<F> o = shfl_xyzz(v, lane ^ d);
        v = xadd(v, o);
    }
    return v;
}

// RED2_SPLIT blocks of 8 warps per window; the tpw strips of the window form tpw/32 rows of 32.  Pass 0: each row is reduced by
// one warp (the rows of a window are dealt over its RED2_SPLIT x 8 warps) into P = sum S1, Q = sum l*S1, A = sum S2, written
// to a per-window scratch in global memory (3 x 32 entries).  The block that finishes last (ticket counter) runs pass 1 on
// its warp 0: wp = sum row*P_row, qs = sum Q_row, as = sum A_row.  Then W = as + strip * (32 * wp + qs).
static constexpr uint32_t RED2_SPLIT = 4;
template <class F> __device__ __forceinline__ Xyzz<F> load_xyzz_cg(const uint32_t *base, size_t i) {      // L2 (written by other blocks)
    constexpr int W = FieldTraits<F>::WORDS;
    const uint4 *s = reinterpret_cast<const uint4 *>(base + i * 4 * W);
    Xyzz<F> p;
#pragma unroll
    for (int j = 0; j < W; j++) {
        const uint4 v = __ldcg(s + j);
        const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int idx = 4 * j + k;
            if (idx < W) set_word(p.x, idx, w4[k]);
            else if (idx < 2 * W) set_word(p.y, idx - W, w4[k]);
            else if (idx < 3 * W) set_word(p.zz, idx - 2 * W, w4[k]);
            else set_word(p.zzz, idx - 3 * W, w4[k]);
        }
    }
    return p;
}
template <class F> __global__ void __launch_bounds__(256) k_msm_reduce2(const uint32_t *s1, const uint32_t *s2, MsmGeom g, uint32_t *wsum,
                                                                        uint32_t *rows, uint32_t *cnt) {
    constexpr int W = FieldTraits<F>::WORDS;
    __shared__ __align__(16) uint32_t sm[4 * 4 * W];       // [wp | qs | as]
    __shared__ uint32_t is_last;
    const uint32_t w = blockIdx.x / RED2_SPLIT, part = blockIdx.x % RED2_SPLIT;
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nrows = g.tpw >> 5;
    uint32_t *wrows = rows + (size_t)w * 96 * 4 * W;
    const uint32_t gw = part * 8 + wid, stride = 8 * RED2_SPLIT;
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
        // items of this pass for this warp: pass 0 -> (row, kind) for rows gw, gw + stride, ..; pass 1 -> kind only (warp 0 of the last block)
        const uint32_t nitems = pass == 0 ? (gw < nrows ? 3 * ((nrows - gw + stride - 1) / stride) : 0u) : (wid == 0 ? 3u : 0u);
#pragma unroll 1
        for (uint32_t it = 0; it < nitems; it++) {
            const uint32_t kind = it % 3;                      // 0: P / wp (weighted in pass 1), 1: Q (weighted in pass 0) / qs, 2: A / as
            Xyzz<F> v;
            bool weighted;
            uint32_t *dst_base;
            uint32_t dst;
            if (pass == 0) {
                const uint32_t row = gw + stride * (it / 3);
                v = load_xyzz<F>(kind == 2 ? s2 : s1, (size_t)w * g.tpw + row * 32 + lane);
                weighted = kind == 1;
                dst_base = wrows;
                dst = kind * 32 + row;
            } else {
                v = xyzz_infinity<F>();
                if (lane < nrows) v = load_xyzz_cg<F>(wrows, kind * 32 + lane);     // rows beyond nrows were never written
                weighted = kind == 0;
                dst_base = sm;
                dst = kind;
            }
            v = warp_reduce(v, weighted);
            if (lane == 0) store_xyzz<F>(dst_base, dst, v);
        }
        if (pass == 0) {
            __threadfence();                                   // the row results of this block are visible before its ticket
            __syncthreads();
            if (threadIdx.x == 0) is_last = atomicAdd(&cnt[w], 1u) == RED2_SPLIT - 1 ? 1u : 0u;
            __syncthreads();
            if (!is_last) return;
            __threadfence();
        } else __syncthreads();
    }
    if (threadIdx.x == 0) {
        Xyzz<F> b = load_xyzz<F>(sm, 0);                       // wp
        uint32_t ls = 0;
        while ((1u << ls) < g.strip) ls++;
#pragma unroll 1
        for (uint32_t i = 0; i <= 5 + ls; i++) {               // b = strip * (32 * wp + qs)
            if (i == 5) b = xadd(b, load_xyzz<F>(sm, 1));
            if (i < 5 + ls) b = xdbl(b);
        }
        store_xyzz<F>(wsum, w, xadd(load_xyzz<F>(sm, 2), b));
    }
}

// ---- lane-parallel curve operations for the serial Horner chain of k_msm_final -----------------------------------------
// A chain of 255 dependent doublings on ONE thread occupies a whole warp slot of the multiplier pipe for one lane's worth
// of work (a lone warp runs a Montgomery multiplication in ~850 cycles whatever the instruction-level parallelism; measured,
// DESIGN.md section 4).  Here every lane of the warp holds the same point; at each dependency level of the formula lane l
// computes the l-th independent product and the results are exchanged with shuffles, so a doubling costs 3 multiplication
// times instead of 9 and an addition 4 instead of 14.  All 32 lanes run the same code (lanes beyond the level's width
// recompute one of the products), so there is no divergence; every lane returns the same result.
template <class F> __device__ __forceinline__ F shfl_f(const F &v, int src_lane) {
    F r;
#pragma unroll
    for (int j = 0; j < FieldTraits<F>::WORDS; j++) set_word(r, j, __shfl_sync(0xffffffffu, get_word(v, j), src_lane));
    return r;
}
template <class F> __device__ __forceinline__ Xyzz<F> xdbl_lanes(const Xyzz<F> &p) {
    const int lane = threadIdx.x & 31;
    const F u = dbl(p.y);
    F a = select(lane == 1, p.x, u);                                   // lane 0: u^2 = v, lane 1: x^2
    F t = mul(a, a);
    const F v = shfl_f(t, 0), xx = shfl_f(t, 1);
    const F m = add(dbl(xx), xx);
    a = select(lane == 1, p.x, select(lane == 2, m, select(lane == 3, p.zz, u)));      // u v = w | x v = s | m m | zz v
    F b = select(lane == 2, m, v);
    t = mul(a, b);
    const F w = shfl_f(t, 0), s = shfl_f(t, 1), mm = shfl_f(t, 2);
    Xyzz<F> r;
    r.zz = shfl_f(t, 3);
    r.x = sub(sub(mm, s), s);
    a = select(lane == 0, m, w);                                       // m (s - x3) | w y | w zzz
    b = select(lane == 0, sub(s, r.x), select(lane == 1, p.y, p.zzz));
    t = mul(a, b);
    r.y = sub(shfl_f(t, 0), shfl_f(t, 1));
    r.zzz = shfl_f(t, 2);
    return r;                                                          // infinity (zz = 0) stays infinity
}
template <class F> __device__ __forceinline__ Xyzz<F> xadd_lanes(const Xyzz<F> &p, const Xyzz<F> &q) {
    const int lane = threadIdx.x & 31;
    // level 1: x1 zz2 = u1 | x2 zz1 = u2 | y1 zzz2 = s1 | y2 zzz1 = s2 | zz1 zz2 | zzz1 zzz2
    F a = select(lane == 0, p.x, select(lane == 1, q.x, select(lane == 2, p.y, select(lane == 3, q.y, select(lane == 4, p.zz, p.zzz)))));
    F b = select(lane == 0, q.zz, select(lane == 1, p.zz, select(lane == 2, q.zzz, select(lane == 3, p.zzz, select(lane == 4, q.zz, q.zzz)))));
    F t = mul(a, b);
    const F u1 = shfl_f(t, 0), u2 = shfl_f(t, 1), s1 = shfl_f(t, 2), s2 = shfl_f(t, 3), zz12 = shfl_f(t, 4), zzz12 = shfl_f(t, 5);
    const F pp_ = sub(u2, u1), rr = sub(s2, s1);
    a = select(lane == 1, rr, pp_);                                    // P^2 = pp | R^2
    t = mul(a, a);
    const F pp = shfl_f(t, 0), rr2 = shfl_f(t, 1);
    a = select(lane == 0, pp_, select(lane == 1, u1, zz12));           // P pp = ppp | u1 pp = q | zz1 zz2 pp
    t = mul(a, pp);
    const F ppp = shfl_f(t, 0), qq = shfl_f(t, 1);
    Xyzz<F> r;
    r.zz = shfl_f(t, 2);
    r.x = sub(sub(sub(rr2, ppp), qq), qq);
    a = select(lane == 0, rr, select(lane == 1, s1, zzz12));           // R (q - x3) | s1 ppp | zzz1 zzz2 ppp
    b = select(lane == 0, sub(qq, r.x), ppp);
    t = mul(a, b);
    r.y = sub(shfl_f(t, 0), shfl_f(t, 1));
    r.zzz = shfl_f(t, 2);
    const bool p_inf = is_zero(p.zz), q_inf = is_zero(q.zz);
    if (!p_inf & !q_inf & is_zero(pp_) & is_zero(rr)) r = xdbl_lanes(p);     // warp-uniform: every lane holds the same points
    r = select(q_inf, p, r);
    return select(p_inf, q, r);
}

// Horner over the windows (most significant first), one warp (see above); result as affine wire words (out_wire)
template <class F> __global__ void __launch_bounds__(32) k_msm_final(const uint32_t *wsum, MsmGeom g, uint32_t *out_wire) {
    if (blockIdx.x) return;
    Xyzz<F> acc = load_xyzz<F>(wsum, g.nwin - 1);
#pragma unroll 1
    for (int w = (int)g.nwin - 2; w >= 0; w--) {
#pragma unroll 1
        for (uint32_t j = 0; j < win_width(g, (uint32_t)w); j++) acc = xdbl_lanes(acc);
        acc = xadd_lanes(acc, load_xyzz<F>(wsum, w));
    }
    if (threadIdx.x) return;
    bool inf = is_zero(acc.zz);
    F t = inv(mul(acc.zz, acc.zzz));
    Aff<F> a;
    a.x = mul(acc.x, mul(t, acc.zzz));        // X / ZZ
    a.y = mul(acc.y, mul(t, acc.zz));         // Y / ZZZ
    uint32_t o[Wire<F>::WORDS_UNCOMPRESSED];
    point_encode<F>(o, a, inf, ENC_UNCOMPRESSED);
    for (int j = 0; j < Wire<F>::WORDS_UNCOMPRESSED; j++) out_wire[j] = o[j];
}

// Subgroup probe: one warp per window sum W_w (XYZZ, any point of E'(Fq2)): [r] W_w by MSB-first double-and-add with the
// lane-parallel complete formulas above; *route is raised unless the result is the point at infinity, i.e. unless W_w lies in
// the order-r subgroup (#E'(Fq2) = r h with gcd(r, h) = 1, so E'[r] is exactly that subgroup).
template <class F> __global__ void __launch_bounds__(32) k_msm_probe_order(const uint32_t *wsum, uint32_t nwin, uint32_t *route) {
    if (blockIdx.x >= nwin) return;
    const Xyzz<F> w = load_xyzz<F>(wsum, blockIdx.x);
    Xyzz<F> acc = w;                                                   // bit 253 of r is its top bit
#pragma unroll 1
    for (int i = 252; i >= 0; i--) {
        acc = xdbl_lanes(acc);
        if ((FrP::p(i >> 5) >> (i & 31)) & 1u) acc = xadd_lanes(acc, w);      // warp-uniform
    }
    if (threadIdx.x == 0 && !is_zero(acc.zz)) atomicOr(route, 1u);
}

// sum of `count` affine wire points (multi-GPU combination of per-rank results)
template <class F> __global__ void k_sum_points(const uint32_t *wire, uint32_t count, uint32_t *out_wire, unsigned long long *err) {
    if (threadIdx.x || blockIdx.x) return;
    constexpr int WU = Wire<F>::WORDS_UNCOMPRESSED;
    Xyzz<F> acc = xyzz_infinity<F>();
    for (uint32_t i = 0; i < count; i++) {
        uint32_t w[WU];
        for (int j = 0; j < WU; j++) w[j] = wire[i * WU + j];
        Aff<F> a;
        bool inf;
        int rc = point_decode<F>(a, inf, w, ENC_UNCOMPRESSED, true);
        if (rc) { report_err(err, i, P2B_EDECODE, rc); continue; }
        if (!inf) acc = xyzz_madd(acc, a);
    }
    bool inf = is_zero(acc.zz);
    F t = inv(mul(acc.zz, acc.zzz));
    Aff<F> a;
    a.x = mul(acc.x, mul(t, acc.zzz));
    a.y = mul(acc.y, mul(t, acc.zz));
    uint32_t o[WU];
    point_encode<F>(o, a, inf, ENC_UNCOMPRESSED);
    for (int j = 0; j < WU; j++) out_wire[j] = o[j];
}

// ------------------------------------------------------------------------------------------------- host side
// The heavy-bucket and reduction kernels are compiled in their own translation units (msm_g{1,2}_heavy.cu,
// msm_g{1,2}_reduce.cu) so that the library builds in parallel; these are their launchers.
template <class F> void msm_launch_heavy_impl(Ctx *c, const uint32_t *aff, const uint32_t *sorted, uint32_t *buckets, const MsmHeavy &hv) {
    k_msm_heavy<F><<<c->sm_count * 4, 128, 0, c->stream>>>(aff, sorted, hv);
    k_msm_heavy_combine<F><<<8, 128, 0, c->stream>>>(buckets, hv);
}
template <class F> void msm_launch_reduce_impl(Ctx *c, const uint32_t *buckets, const MsmGeom &g, uint32_t *s1, uint32_t *s2, uint32_t *wsum,
                                               uint32_t *d_out_wire, size_t nred) {
    k_msm_reduce1<F><<<(int)((nred + 127) / 128), 128, 0, c->stream>>>(buckets, g, s1, s2);
    // scratch behind wsum (reserved by msm_typed): per-window row results and the ticket counters of k_msm_reduce2
    uint32_t *rows = wsum + (size_t)g.nwin * 4 * FieldTraits<F>::WORDS, *cnt = rows + (size_t)g.nwin * 96 * 4 * FieldTraits<F>::WORDS;
    cudaMemsetAsync(cnt, 0, (size_t)g.nwin * 4, c->stream);
    k_msm_reduce2<F><<<g.nwin * RED2_SPLIT, 256, 0, c->stream>>>(s1, s2, g, wsum, rows, cnt);
    if (d_out_wire) k_msm_final<F><<<1, 32, 0, c->stream>>>(wsum, g, d_out_wire);
}
void msm_launch_heavy_g1(Ctx *c, const uint32_t *aff, const uint32_t *sorted, uint32_t *buckets, const MsmHeavy &hv);
void msm_launch_heavy_g2(Ctx *c, const uint32_t *aff, const uint32_t *sorted, uint32_t *buckets, const MsmHeavy &hv);
void msm_launch_reduce_g1(Ctx *c, const uint32_t *buckets, const MsmGeom &g, uint32_t *s1, uint32_t *s2, uint32_t *wsum, uint32_t *d_out_wire, size_t nred);
void msm_launch_reduce_g2(Ctx *c, const uint32_t *buckets, const MsmGeom &g, uint32_t *s1, uint32_t *s2, uint32_t *wsum, uint32_t *d_out_wire, size_t nred);
void msm_launch_sum_points_g1(Ctx *c, const uint32_t *d_in, uint32_t count, uint32_t *d_out);
void msm_launch_sum_points_g2(Ctx *c, const uint32_t *d_in, uint32_t count, uint32_t *d_out);
void msm_launch_probe_order_g2(Ctx *c, const uint32_t *wsum, uint32_t nwin, uint32_t *route);

// Window geometry.  The 255 bits (r < 2^254 plus one bit of head-room for the signed-digit carry) are split into nwin
// windows whose widths differ by at most one bit -- a short top window would put n / 2^bits terms into each of a few
// buckets and serialise them on a few threads (profiles/r1a/msm_c_sweep.log: 165 s at 2^26 with a 2-bit top window).
// The narrow windows (twice the load per bucket, half the buckets) come first so that their longer threads start first.
// Width: the widest c <= 0.6 log2(n) + 3.5 that is minimal for its window count -- a fit to the sweeps on B200 with
// size-ordered buckets (2^20 -> 15, 2^22 -> 16, 2^24 -> 17, 2^26 -> 19): more windows cost n mixed adds each, wider windows
// lengthen the serial strip walk of the bucket reduction and leave the accumulation threads shorter lists.
static inline MsmGeom msm_geometry(size_t n, uint32_t scalar_bits = 0) {
    uint32_t lg = 0;
    while (((size_t)1 << (lg + 1)) <= n) lg++;
    int target = (int)(0.6 * lg + 3.5);
    if (const char *e = getenv("P2B_MSM_C")) {      // tuning override (results do not depend on the geometry)
        int v = atoi(e);
        if (v >= 2 && v <= 22) target = v;
    }
    if (target < 6) target = 6;
    if (target > 20) target = 20;
    // bits to cover: the scalar plus one bit of head-room for the signed-digit carry (255 for any canonical scalar)
    const uint32_t total = scalar_bits && scalar_bits < 254 ? scalar_bits + 1 : 255;
    if ((uint32_t)target > total) target = (int)total;
    MsmGeom g{};
    for (int c = target; c >= 2; c--) {
        const uint32_t nwin = (total + c - 1) / c;
        if ((total + nwin - 1) / nwin != (uint32_t)c) continue;       // fewer bits would do for this many windows
        g.c = (uint32_t)c;
        g.nwin = nwin;
        break;
    }
    g.nnarrow = g.nwin * g.c - total;
    g.nb = 1u << (g.c - 1);
    g.nbk = g.nb + 1;
    g.strip = g.nb > 1024 ? g.nb / 1024 : 1;
    g.tpw = g.nb / g.strip;
    return g;
}

// One MSM, or one chunk of a streamed MSM.
enum { MSM_FIRST = 1, MSM_LAST = 2 };
enum { MSM_SINGLE = 0, MSM_PAIR_SHIFTED = 1, MSM_PAIR_SEPARATE = 2 };
struct MsmJob {
    const void *d_points = nullptr;     // n points (n + 1 for MSM_PAIR_SHIFTED: B_i = A_(i+1)); encoding `in_enc`
    const void *d_points_b = nullptr;   // MSM_PAIR_SEPARATE: the second array
    const void *d_scalars = nullptr;    // n x 32 bytes big-endian, canonical
    size_t n = 0;
    uint32_t *d_out_wire = nullptr;     // result (uncompressed wire); pair modes: result B follows at + WORDS_UNCOMPRESSED
    size_t geom_n = 0;                  // >= n: fixes the scratch sizes for all chunks of a streamed MSM
    int phase = MSM_FIRST | MSM_LAST;   // MSM_FIRST starts fresh buckets, MSM_LAST reduces and writes the result
    uint64_t err_base = 0;
    size_t total_n = 0;                 // size of the WHOLE sum (window widths follow it), 0 = geom_n
    int pair = MSM_SINGLE;
    int in_enc = ENC_UNCOMPRESSED;      // ENC_UNCOMPRESSED (wire) or ENC_RAW_MONT_LE (decoded by an earlier kernel)
    int check = 0;                      // bit 0: is_on_curve on every point (CheckForCorrectness::Yes); bit 1: infinity is an error
    uint32_t scalar_bits = 0;           // the caller guarantees scalars < 2^scalar_bits (0: any canonical scalar)
    uint32_t *d_route = nullptr;        // check bit 2: device word raised when a point is off the curve (subgroup probe)
};
template <class F> int msm_typed(Ctx *c, const MsmJob &j) {
    constexpr int W = FieldTraits<F>::WORDS, WU = Wire<F>::WORDS_UNCOMPRESSED;
    const size_t n = j.n, geom_n = j.geom_n;
    const bool pair = j.pair != MSM_SINGLE;
    if (geom_n >= ((size_t)1 << 31) || n > geom_n) return ctx_fail(c, P2B_EARG, "msm: chunk must be < 2^31 terms");
    // window widths follow the size of the WHOLE sum (total_n, when this call is one chunk of a streamed MSM); the scratch
    // buffers are sized for the largest chunk (geom_n)
    MsmGeom g = msm_geometry(j.total_n ? j.total_n : (geom_n ? geom_n : 1), j.scalar_bits);
    const size_t cap_n = geom_n ? geom_n : 1;
    if ((uint64_t)g.nwin * cap_n >= (1ull << 32)) return ctx_fail(c, P2B_EARG, "msm: too many terms for one pass (use the host entry point, which streams)");
    const size_t nslots = (size_t)g.nwin * g.nbk;
    const size_t xy = (size_t)4 * W * 4;                       // bytes per XYZZ point
    const size_t nsets = pair ? 2 : 1;
    int rc;
    // msm_a: affine Montgomery points ; msm_b: hist | offsets | cursor ; msm_c: sorted entries ; msm_d: buckets | s1 | s2 | wsum
    if ((rc = dev_reserve(c, c->msm_a, (nsets * cap_n + 1) * WU * 4))) return rc;
    if ((rc = dev_reserve(c, c->msm_b, (4 * nslots + 16 + nslots / SCAN_TILE + 8) * 4))) return rc;
    if ((rc = dev_reserve(c, c->msm_c, (size_t)g.nwin * cap_n * 4))) return rc;
    const size_t nred = (size_t)g.nwin * g.tpw;
    if ((rc = dev_reserve(c, c->msm_d, (nsets * nslots + 2 * nred + g.nwin + (size_t)g.nwin * 96) * xy + (size_t)g.nwin * 4 + 64))) return rc;
    uint32_t *aff = (uint32_t *)c->msm_a.p;
    uint32_t *aff_b = j.pair == MSM_PAIR_SHIFTED ? aff + WU : aff + (cap_n + 1) * WU;
    uint32_t *hist = (uint32_t *)c->msm_b.p, *cursor = hist + nslots, *offsets = cursor + nslots;   // offsets: nslots + one per group
    uint32_t *perm = offsets + nslots + 16, *tile_sums = perm + nslots;
    uint32_t *sorted = (uint32_t *)c->msm_c.p;
    uint32_t *buckets = (uint32_t *)c->msm_d.p, *buckets_b = buckets + nslots * 4 * W;
    uint32_t *s1 = buckets + nsets * nslots * 4 * W, *s2 = s1 + nred * 4 * W, *wsum = s2 + nred * 4 * W;
    // overflow handling for skewed digit distributions (see MsmHeavy)
    MsmHeavy hv;
    uint32_t *size_hist = nullptr, *size_start = nullptr;
    {
        const size_t load = cap_n / g.nb + 1;                 // average terms per bucket of a wide window
        hv.seg = (uint32_t)(8 * load < 256 ? 256 : 8 * load);
        hv.chunk = 1u << 15;
        if (const char *e = getenv("P2B_MSM_SEG")) {           // test hook: tiny segments exercise the overflow path
            int v = atoi(e);
            if (v > 0) { hv.seg = (uint32_t)v; hv.chunk = (uint32_t)(4 * v); }
        }
        const size_t total = (size_t)g.nwin * cap_n;
        hv.cap_items = (uint32_t)(total / hv.chunk + total / hv.seg + 16);
        hv.cap_buckets = (uint32_t)(total / hv.seg + 16);
        const size_t heavy_bytes = 64 + ((size_t)hv.cap_items + hv.cap_buckets) * 16 + (size_t)hv.cap_items * xy;
        if ((rc = dev_reserve(c, c->msm_e, heavy_bytes + 2 * ((size_t)hv.seg + 2) * 4 + 64))) return rc;
        char *hp = (char *)c->msm_e.p;
        size_hist = (uint32_t *)(hp + ((heavy_bytes + 15) & ~(size_t)15));
        size_start = size_hist + hv.seg + 2;
        hv.counters = (uint32_t *)hp;
        hv.items = (uint4 *)(hp + 64);
        hv.hbuckets = hv.items + hv.cap_items;
        hv.partials = (uint32_t *)(hv.hbuckets + hv.cap_buckets);
    }
    int grid = (int)((n + 255) / 256);
    if (grid > c->sm_count * 16) grid = c->sm_count * 16;
    if (grid < 1) grid = 1;
    // The windows are processed in groups: the counting sort of group i+1 (L2 atomics and scattered 4-byte stores, on the
    // high-priority sort stream) overlaps the bucket accumulation of group i (multiplier bound, on the compute stream).
    uint32_t ngroups = n >= ((size_t)1 << 18) ? 3 : 1;
    if (ngroups > g.nwin) ngroups = g.nwin;
    cudaStream_t S = ngroups > 1 ? c->sort_stream : c->stream, C = c->stream;
    if (S != C) {
        P2B_CUDA(c, cudaEventRecord(c->msm_ev[0], C));
        P2B_CUDA(c, cudaStreamWaitEvent(S, c->msm_ev[0], 0));
    }
    prof_begin(c, P2B_PROF_MSM_SORT, S);
    P2B_CUDA(c, cudaMemsetAsync(hist, 0, nslots * 4, S));
    P2B_CUDA(c, cudaMemsetAsync(size_hist, 0, ((size_t)hv.seg + 2) * 4, S));
    // the point decode is only needed by the accumulation: it runs on the compute stream, next to the first group's sort
    const size_t np_a = n + (j.pair == MSM_PAIR_SHIFTED ? 1 : 0);
    k_msm_prepare<F><<<grid, 256, 0, C>>>((const uint32_t *)j.d_points, aff, np_a, c->d_err, j.err_base, j.in_enc, j.check, j.d_route);
    c->launches++;
    if (j.pair == MSM_PAIR_SEPARATE) {
        k_msm_prepare<F><<<grid, 256, 0, C>>>((const uint32_t *)j.d_points_b, aff_b, n, c->d_err, j.err_base, j.in_enc, j.check);
        c->launches++;
    }
    const int fst = (j.phase & MSM_FIRST) != 0;
    for (uint32_t gi = 0; gi < ngroups; gi++) {
        // the first group is small: its sort is the only one that cannot hide behind an accumulation
        static const int first_env = [] { const char *e = getenv("P2B_MSM_FIRST_GROUP"); return e ? atoi(e) : 0; }();   // tuning hook
        uint32_t first_hi = g.nwin / 7 ? g.nwin / 7 : 1;
        // below 2^23 terms two windows do not fill the GPU (one thread per bucket: 2 x 2^14 buckets at 2^20 against 56,832
        // resident threads) while the group's sort is short anyway: start with a quarter of the windows
        if (n < ((size_t)1 << 23) && g.nwin >= 8) first_hi = g.nwin / 4;
        if (first_env > 0 && (uint32_t)first_env < g.nwin) first_hi = (uint32_t)first_env;
        const uint32_t bounds[4] = {0, ngroups == 3 ? first_hi : g.nwin, ngroups == 3 ? (g.nwin + first_hi) / 2 : g.nwin, g.nwin};
        const uint32_t w_lo = bounds[gi], w_hi = ngroups == 3 ? bounds[gi + 1] : g.nwin;
        const uint32_t slot_lo = w_lo * g.nbk, slot_cnt = (w_hi - w_lo) * g.nbk;
        uint32_t *offs = offsets + slot_lo + gi;
        k_msm_hist<<<grid, 256, 0, S>>>((const uint32_t *)j.d_scalars, n, g, hist, c->d_err, j.err_base, w_lo, w_hi);
        {
            const uint32_t ntiles = (slot_cnt + SCAN_TILE - 1) / SCAN_TILE;
            k_msm_scan_tiles<<<ntiles, 256, 0, S>>>(hist + slot_lo, slot_cnt, tile_sums);
            k_msm_scan_tops<<<1, 32, 0, S>>>(tile_sums, ntiles, (uint32_t)((size_t)w_lo * cap_n));
            k_msm_scan_apply<<<ntiles, 256, 0, S>>>(hist + slot_lo, offs, cursor + slot_lo, slot_cnt, tile_sums, ntiles);
        }
        k_msm_scatter<<<grid, 256, 0, S>>>((const uint32_t *)j.d_scalars, n, g, cursor, sorted, w_lo, w_hi);
        {   // order the group's buckets by size (see k_msm_size_hist)
            const int sgrid = (int)((slot_cnt + 255) / 256) < c->sm_count * 8 ? (int)((slot_cnt + 255) / 256) : c->sm_count * 8;
            const uint32_t nbins = hv.seg + 1;
            if (nbins * 4 <= 40960) {                // block-private histograms in shared memory
                int tgrid = (int)((slot_cnt + SIZE_TILE - 1) / SIZE_TILE);
                if (tgrid > c->sm_count * 4) tgrid = c->sm_count * 4;
                k_msm_size_hist_smem<<<tgrid, 256, nbins * 4, S>>>(offs, slot_cnt, hv.seg, size_hist, nbins);
                k_msm_size_scan<<<1, 1024, 0, S>>>(size_hist, size_start, nbins);
                k_msm_size_scatter_smem<<<tgrid, 256, nbins * 4, S>>>(offs, slot_cnt, hv.seg, size_start, perm + slot_lo, nbins);
            } else {
                k_msm_size_hist<<<sgrid, 256, 0, S>>>(offs, slot_cnt, hv.seg, size_hist);
                k_msm_size_scan<<<1, 1024, 0, S>>>(size_hist, size_start, nbins);
                k_msm_size_scatter<<<sgrid, 256, 0, S>>>(offs, slot_cnt, hv.seg, size_start, perm + slot_lo);
            }
        }
        if (gi + 1 == ngroups) prof_end(c, P2B_PROF_MSM_SORT, (int)(8 * ngroups), S);
        if (S != C) {
            P2B_CUDA(c, cudaEventRecord(c->msm_ev[1 + gi], S));
            P2B_CUDA(c, cudaStreamWaitEvent(C, c->msm_ev[1 + gi], 0));
        }
        const int agrid = (int)((slot_cnt + 127) / 128);
        prof_begin(c, P2B_PROF_MSM_ACCUMULATE);
        if (pair) {
            static const int pair_g2 = [] { const char *e = getenv("P2B_ACC_VARIANT_G2"); return e ? atoi(e) : 0; }();
            if (W == 16 && pair_g2 != 9) {
                MsmHeavy none = hv;
                k_msm_accumulate_g2x2<2><<<(int)((slot_cnt + 63) / 64), 128, 0, C>>>(aff, aff_b, offs, sorted, buckets, buckets_b, fst, slot_lo, slot_cnt,
                                                                                      none, c->d_err, perm + slot_lo);
            } else
            k_msm_accumulate_pair<F><<<agrid, 128, 0, C>>>(aff, aff_b, offs, sorted, buckets, buckets_b, fst, slot_lo, slot_cnt, perm + slot_lo);
            prof_end(c, P2B_PROF_MSM_ACCUMULATE, 1);
            c->launches += 9;
            continue;
        }
        P2B_CUDA(c, cudaMemsetAsync(hv.counters, 0, 8, C));
        {
            static const int variant = [] { const char *e = getenv("P2B_ACC_VARIANT"); return e ? atoi(e) : P2B_ACC_DEFAULT_VARIANT; }();
#define P2B_ACC_LAUNCH(V) k_msm_accumulate<F, V><<<agrid, 128, 0, C>>>(aff, offs, sorted, g, buckets, fst, slot_lo, slot_cnt, hv, c->d_err, perm + slot_lo)
            static const int variant_g2 = [] { const char *e = getenv("P2B_ACC_VARIANT_G2"); return e ? atoi(e) : 0; }();
            if (W != 8) {                                       // G2: two lanes per bucket; the one-thread variants stay selectable
                if constexpr (W != 8) {
                    if (variant_g2 == 5) P2B_ACC_LAUNCH(5);
                    else if (variant_g2 == 1) P2B_ACC_LAUNCH(1);
                    else if (variant_g2 == 9) P2B_ACC_LAUNCH(0);
                    else k_msm_accumulate_g2x2<1><<<(int)((slot_cnt + 63) / 64), 128, 0, C>>>(aff, aff, offs, sorted, buckets, buckets, fst, slot_lo,
                                                                                                slot_cnt, hv, c->d_err, perm + slot_lo);
                }
            }
            else if (variant == 0) P2B_ACC_LAUNCH(0);
            else if (variant == 1) { if constexpr (W == 8) P2B_ACC_LAUNCH(1); }
            else if (variant == 2) { if constexpr (W == 8) P2B_ACC_LAUNCH(2); }
            else if (variant == 4) { if constexpr (W == 8) P2B_ACC_LAUNCH(4); }
            else { if constexpr (W == 8) P2B_ACC_LAUNCH(3); }
#undef P2B_ACC_LAUNCH
        }
        if constexpr (W == 8) msm_launch_heavy_g1(c, aff, sorted, buckets, hv);
        else msm_launch_heavy_g2(c, aff, sorted, buckets, hv);
        prof_end(c, P2B_PROF_MSM_ACCUMULATE, 3);
        c->launches += 11;
    }
    if (!(j.phase & MSM_LAST)) {
        P2B_CUDA(c, cudaGetLastError());
        return P2B_OK;
    }
    prof_begin(c, P2B_PROF_MSM_REDUCE);
    for (size_t set = 0; set < nsets; set++) {
        const uint32_t *bk = set ? buckets_b : buckets;
        uint32_t *out = j.d_out_wire ? j.d_out_wire + set * WU : nullptr;      // nullptr: window sums only (no Horner)
        if constexpr (W == 8) msm_launch_reduce_g1(c, bk, g, s1, s2, wsum, out, nred);
        else msm_launch_reduce_g2(c, bk, g, s1, s2, wsum, out, nred);
        c->launches += 3;
    }
    prof_end(c, P2B_PROF_MSM_REDUCE, (int)(3 * nsets));
    c->msm_last_wsum = wsum;            // window sums of the (last) set: W_w = sum_i digit_w(k_i) P_i
    c->msm_last_nwin = g.nwin;
    P2B_CUDA(c, cudaGetLastError());
    return P2B_OK;
}


}  // namespace p2b
