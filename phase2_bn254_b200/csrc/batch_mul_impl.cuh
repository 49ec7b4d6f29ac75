// batch_mul_impl.cuh -- K1/K2/K3/K7: batched independent scalar multiplication out[i] = [k_i] in[i] on sm_100a.
//
// GPU replacement for the reference's `batch_exp` closures
// (powersoftau/src/batched_accumulator.rs:1130-1181 with the tau-power generation of :1201-1216, and
// phase2/src/parameters.rs:424-470) fused with the chunk codec around them (read_points_chunk :889-1001,
// write_point :1052-1093).
//
// Pipeline per call (all on ctx->stream):
//   [k_decompress]   compressed wire -> raw affine Montgomery (sqrt per point; only for compressed input)
//   [k_pow_tables]   4 x 1024 table of tau^(e << 10t) (x coeff) so that tau^i costs 3 Fr mults per lane
//   k_batch_mul      one thread per point: decode (BE -> limbs -> Montgomery), scalar (array / broadcast /
//                    tau^(start+i)*coeff), [k]P by GLV + signed fixed windows (smul.cuh), Jacobian result to HBM
//   k_normalize      Montgomery-trick batched inversion (1 Fermat inversion per ~32 points), affine conversion,
//                    canonical compare for the sign bit, BE encode -- the reference's batch_normalization
//                    (ec.rs:251-299) + into_affine + into_compressed (ec.rs:920-945)
//
// Data layout in HBM: wire encodings are arrays-of-structs exactly as in the files (64/128/32/64 B per point);
// a warp reads/writes one contiguous 1-4 KB span.  Jacobian intermediates are structure-of-arrays (X[n], Y[n],
// Z[n], 32 or 64 B per element) so that consecutive lanes touch consecutive 32 B sectors.
#pragma once
#include <cstdio>
#include <cstring>
#include "codec.cuh"
#include "p2b_internal.h"
#include "smul.cuh"

namespace p2b {

// ------------------------------------------------------------------------------------------------- helpers
template <int NW> __device__ __forceinline__ void load_words(uint32_t *dst, const uint32_t *src) {
    const uint4 *s = reinterpret_cast<const uint4 *>(src);
#pragma unroll
    for (int i = 0; i < NW / 4; i++) {
        uint4 v = __ldg(s + i);
        dst[4 * i] = v.x; dst[4 * i + 1] = v.y; dst[4 * i + 2] = v.z; dst[4 * i + 3] = v.w;
    }
}
template <int NW> __device__ __forceinline__ void store_words(uint32_t *dst, const uint32_t *src) {
    uint4 *d = reinterpret_cast<uint4 *>(dst);
#pragma unroll
    for (int i = 0; i < NW / 4; i++) d[i] = make_uint4(src[4 * i], src[4 * i + 1], src[4 * i + 2], src[4 * i + 3]);
}
template <class F> __device__ __forceinline__ F load_elem(const uint32_t *base, size_t i) {
    constexpr int W = FieldTraits<F>::WORDS;
    uint32_t w[W];
    load_words<W>(w, base + i * W);
    F r;
#pragma unroll
    for (int j = 0; j < W; j++) set_word(r, j, w[j]);
    return r;
}
template <class F> __device__ __forceinline__ void store_elem(uint32_t *base, size_t i, const F &v) {
    constexpr int W = FieldTraits<F>::WORDS;
    uint32_t w[W];
#pragma unroll
    for (int j = 0; j < W; j++) w[j] = get_word(v, j);
    store_words<W>(base + i * W, w);
}
__device__ __forceinline__ void report(unsigned long long *err, uint64_t index, int kind, int sub) {
    atomicMin(err, (unsigned long long)((index << 8) | ((uint64_t)kind << 4) | (uint64_t)sub));
}

// slow generic path kept out of line so it does not weigh on the fast path's registers
template <class F> __device__ __noinline__ void mul_binary_slow(Jac<F> *out, const Aff<F> *p, const uint32_t *k) {
    uint32_t kk[8];
    for (int i = 0; i < 8; i++) kk[i] = k[i];
    *out = mul_binary<F>(*p, kk);
}

// ------------------------------------------------------------------------------------------------- tau tables
// tables[t][e] = tau^(e << (10 t)) for t = 0..3, e = 0..1023 (Montgomery Fr); table 3 is pre-multiplied by coeff.
static __global__ void k_pow_tables(Fr *tables, const uint32_t *tau_canon, const uint32_t *coeff_canon) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= 4096) return;
    int t = j >> 10;
    uint64_t e = (uint64_t)(j & 1023) << (10 * t);
    Fr tau, coeff;
    for (int i = 0; i < 8; i++) { tau.l[i] = tau_canon[i]; coeff.l[i] = coeff_canon[i]; }
    Fr v = pow_u64(to_mont(tau), e);
    if (t == 3) v = mul(v, to_mont(coeff));
    tables[j] = v;
}

// ------------------------------------------------------------------------------------------------- decompress
struct DecompressParams {
    const uint32_t *in;     // compressed wire
    uint32_t *out;          // raw affine Montgomery LE (x || y), all-zero = infinity
    size_t n;
    unsigned long long *err;
    uint64_t err_base;
};
template <class F> __global__ void __launch_bounds__(128) k_decompress(DecompressParams p) {
    constexpr int WC = Wire<F>::WORDS_COMPRESSED, WU = Wire<F>::WORDS_UNCOMPRESSED;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t w[WC];
        load_words<WC>(w, p.in + i * WC);
        Aff<F> a;
        bool inf;
        int rc = point_decode<F>(a, inf, w, ENC_COMPRESSED, false);
        if (rc) { report(p.err, p.err_base + i, P2B_EDECODE, rc); inf = true; }
        uint32_t o[WU];
        point_encode<F>(o, a, inf, ENC_RAW_MONT_LE);
        store_words<WU>(p.out + i * WU, o);
    }
}

// ------------------------------------------------------------------------------------------------- bulk codec
// out[i] = re-encoding of in[i]: decompression (sqrt), checked deserialisation (is_on_curve) and compression without any
// scalar multiplication -- BatchedAccumulator::decompress (batched_accumulator.rs:543-618) and the checked point reads of
// Parameters::read (bellman/src/groth16/mod.rs:287-383).
struct RecodeParams {
    const uint32_t *in;
    uint32_t *out;
    size_t n;
    int in_enc, out_enc, flags;
    unsigned long long *err;
    uint64_t err_base;
};
template <class F> __global__ void __launch_bounds__(128) k_recode(RecodeParams p) {
    constexpr int WU = Wire<F>::WORDS_UNCOMPRESSED, WC = Wire<F>::WORDS_COMPRESSED;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t w[WU];
        if (p.in_enc == ENC_COMPRESSED) load_words<WC>(w, p.in + i * WC);
        else load_words<WU>(w, p.in + i * WU);
        Aff<F> a;
        bool inf;
        int rc = point_decode<F>(a, inf, w, p.in_enc, (p.flags & P2B_CHECK_INPUT) != 0);
        if (rc) { report(p.err, p.err_base + i, P2B_EDECODE, rc); inf = true; }
        else if (inf && (p.flags & P2B_REJECT_INFINITY)) report(p.err, p.err_base + i, P2B_EINFINITY_IN, 0);
        uint32_t o[WU];
        point_encode<F>(o, a, inf, p.out_enc);
        if (p.out_enc == ENC_COMPRESSED) store_words<WC>(p.out + i * WC, o);
        else store_words<WU>(p.out + i * WU, o);
    }
}

// ------------------------------------------------------------------------------------------------- TMA tile staging
// The point (and scalar) tile of a block is fetched by the TMA unit with ONE bulk asynchronous copy per array
// (cp.async.bulk global -> shared, completion counted in bytes on an mbarrier) into a double buffer: while the block works
// on tile t (milliseconds of arithmetic), tile t + gridDim.x is already in flight, so no thread ever waits on HBM and the
// reads of the decompressed accumulator are full-line, perfectly coalesced bursts.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ------------------------------------------------------------------------------------------------- the hot kernel
struct BatchMulParams {
    const uint32_t *in;
    uint32_t *jx, *jy, *jz;
    size_t n;
    int in_enc;       // ENC_UNCOMPRESSED or ENC_RAW_MONT_LE
    int flags;
    int sc_mode;
    const uint32_t *scalars;   // mode 0: n x 8 words (BE bytes)
    uint32_t k[8];             // mode 1
    const Fr *tables;          // mode 2
    uint64_t start;            // mode 2
    unsigned long long *err;
    uint64_t err_base;
    uint4 *gtable;             // G2: per-thread odd-multiples tables in global memory (grid * block columns)
    UniformDigits uni;         // UNIFORM kernels (mode 1): width-5 NAF of the GLV halves of the one scalar
    const uint32_t *route;     // G2 subgroup probe (msm_g2.cu): the launch only runs when *route == route_want (nullptr: always)
    uint32_t route_want;
};

// table policy: G1 keeps its 512 B / thread table in shared memory; G2 (1 KB / thread) keeps it in L2 so that two
// 128-thread blocks fit on an SM
template <class F, int BLOCK> struct TablePolicy;
#ifndef P2B_G1_BLOCK
#define P2B_G1_BLOCK 256
#endif
#ifndef P2B_G1_MIN_BLOCKS
#define P2B_G1_MIN_BLOCKS 1
#endif
template <int BLOCK> struct TablePolicy<Fq, BLOCK> {
    static constexpr int MIN_BLOCKS = P2B_G1_MIN_BLOCKS;
    static constexpr size_t TABLE_BYTES = (size_t)BLOCK * 8 * 2 * 8 * 4;
    static __device__ __forceinline__ StridedTable<Fq> make(uint32_t *smem, const BatchMulParams &) {
        return StridedTable<Fq>{smem + threadIdx.x, BLOCK};   // private column per thread: the table needs no barriers
    }
};
template <int BLOCK> struct TablePolicy<Fq2, BLOCK> {
    static constexpr int MIN_BLOCKS = 2;
    static constexpr size_t TABLE_BYTES = 0;
    static __device__ __forceinline__ GlobalTable<Fq2> make(uint32_t *, const BatchMulParams &p) {
        return GlobalTable<Fq2>{p.gtable + (size_t)blockIdx.x * BLOCK + threadIdx.x, (size_t)gridDim.x * BLOCK};
    }
};

// GLV: use the endomorphism split (always for G1; for G2 only when the caller vouches for subgroup membership)
// UNIFORM: every point is multiplied by the same scalar (mode 1) -> warp-uniform sparse digits (mul_glv_uniform)
template <class F, int BLOCK, bool GLV, bool UNIFORM = false> __global__ void __launch_bounds__(BLOCK, TablePolicy<F, BLOCK>::MIN_BLOCKS) k_batch_mul(const __grid_constant__ BatchMulParams p) {
    extern __shared__ __align__(16) uint32_t smem[];
    constexpr int WU = Wire<F>::WORDS_UNCOMPRESSED;
    constexpr bool IS_G1 = FieldTraits<F>::WORDS == 8;
    const int tid = threadIdx.x;
    // probed G2 batches queue BOTH the split and the exact kernel; the device-side verdict picks the one that runs
    if (p.route && (__ldg(p.route) != 0u) != (p.route_want != 0u)) return;
    const auto tbl = TablePolicy<F, BLOCK>::make(smem, p);
    const size_t ntiles = (p.n + BLOCK - 1) / BLOCK;
    // staging area behind the table: 2 x point tile, 2 x scalar tile, 2 mbarriers
    constexpr size_t TILE_WORDS = (size_t)BLOCK * WU, SC_WORDS = (size_t)BLOCK * 8;
    uint32_t *stage = smem + TablePolicy<F, BLOCK>::TABLE_BYTES / 4;
    uint32_t *pt_tile[2] = {stage, stage + TILE_WORDS};
    uint32_t *sc_tile[2] = {stage + 2 * TILE_WORDS, stage + 2 * TILE_WORDS + SC_WORDS};
    uint64_t *bars = reinterpret_cast<uint64_t *>(stage + 2 * TILE_WORDS + 2 * SC_WORDS);
    auto issue = [&](size_t tile, int buf) {               // one elected thread: arm the barrier, start the bulk copies
        const size_t first = tile * BLOCK, cnt = p.n - first < (size_t)BLOCK ? p.n - first : (size_t)BLOCK;
        const uint32_t pb = (uint32_t)(cnt * WU * 4), sb = p.sc_mode == 0 ? (uint32_t)(cnt * 32) : 0u;
        mbar_expect_tx(&bars[buf], pb + sb);
        tma_load_1d(pt_tile[buf], p.in + first * WU, pb, &bars[buf]);
        if (sb) tma_load_1d(sc_tile[buf], p.scalars + first * 8, sb, &bars[buf]);
    };
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0 && (size_t)blockIdx.x < ntiles) issue(blockIdx.x, 0);
    uint32_t it = 0;
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
        const int buf = it & 1;
        const size_t i = tile * BLOCK + tid;
        // every thread of the previous iteration has copied its operands out of buffer buf^1 (barrier at the end of the
        // staging step below), so it can be refilled with the next tile while this one is being processed
        if (tid == 0 && tile + gridDim.x < ntiles) issue(tile + gridDim.x, buf ^ 1);
        mbar_wait(&bars[buf], (it >> 1) & 1);
        uint32_t w[WU], sw[8];
        if (i < p.n) {
#pragma unroll
            for (int j = 0; j < WU / 4; j++) {
                uint4 v = reinterpret_cast<const uint4 *>(pt_tile[buf] + (size_t)tid * WU)[j];
                w[4 * j] = v.x; w[4 * j + 1] = v.y; w[4 * j + 2] = v.z; w[4 * j + 3] = v.w;
            }
            if (p.sc_mode == 0) {
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    uint4 v = reinterpret_cast<const uint4 *>(sc_tile[buf] + (size_t)tid * 8)[j];
                    sw[4 * j] = v.x; sw[4 * j + 1] = v.y; sw[4 * j + 2] = v.z; sw[4 * j + 3] = v.w;
                }
            }
        }
        __syncthreads();
        if (i >= p.n) continue;
        // ---- point ----
        Aff<F> a;
        bool inf;
        int rc = point_decode<F>(a, inf, w, p.in_enc, false);
        bool oncurve = true;
        if (!rc && !inf && (GLV || (p.flags & P2B_CHECK_INPUT))) oncurve = on_curve(a);
        if (!rc && !oncurve && (p.flags & P2B_CHECK_INPUT)) rc = DEC_NOT_ON_CURVE;
        if (rc) { report(p.err, p.err_base + i, P2B_EDECODE, rc); inf = true; }
        else if (inf && (p.flags & P2B_REJECT_INFINITY)) report(p.err, p.err_base + i, P2B_EINFINITY_IN, 0);
        // ---- scalar (canonical little-endian limbs) ----
        uint32_t k[8];
        if (p.sc_mode == 0) {
            Fr kc = limbs_from_be_words<FrP>(sw);
            if (!is_canonical(kc)) report(p.err, p.err_base + i, P2B_EARG, 0);
#pragma unroll
            for (int j = 0; j < 8; j++) k[j] = kc.l[j];
        } else if (p.sc_mode == 1) {
#pragma unroll
            for (int j = 0; j < 8; j++) k[j] = p.k[j];
        } else {
            uint64_t e = p.start + i;
            Fr v = p.tables[e & 1023];
            v = mul(v, p.tables[1024 + ((e >> 10) & 1023)]);
            v = mul(v, p.tables[2048 + ((e >> 20) & 1023)]);
            v = mul(v, p.tables[3072 + ((e >> 30) & 1023)]);
            v = from_mont(v);
#pragma unroll
            for (int j = 0; j < 8; j++) k[j] = v.l[j];
        }
        // ---- [k]P ----
        Jac<F> r;
        if (inf) {
            r = jac_infinity<F>();
        } else {
            F zr[8];
            bool bad = !oncurve;      // off-curve garbage (unchecked mode): GLV does not apply
            if (!bad) {
                if constexpr (UNIFORM) r = mul_glv_uniform<F>(a, p.uni, tbl, zr, bad);
                else if constexpr (GLV) r = mul_glv<F>(a, k, tbl, zr, bad);
                else r = mul_window4<F>(a, k, tbl, zr, bad);
            }
            if (bad) mul_binary_slow<F>(&r, &a, k);
        }
        store_elem<F>(p.jx, i, r.x);
        store_elem<F>(p.jy, i, r.y);
        store_elem<F>(p.jz, i, r.z);
    }
}

// ------------------------------------------------------------------------------------------------- normalize + encode
struct NormalizeParams {
    const uint32_t *jx, *jy, *jz;
    uint32_t *prefix;     // n elements of F scratch
    uint32_t *out;        // wire
    size_t n;
    int out_enc;
    int flags;
    unsigned long long *err;
    uint64_t err_base;
};
template <class F> __global__ void __launch_bounds__(128) k_normalize(NormalizeParams p) {
    const size_t T = (size_t)gridDim.x * blockDim.x;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.n) return;
    // forward: prefix products of the finite Z's of this thread's strided chunk
    F acc = FieldTraits<F>::one();
    size_t last = t;
    for (size_t i = t; i < p.n; i += T) {
        F z = load_elem<F>(p.jz, i);
        store_elem<F>(p.prefix, i, acc);
        if (!is_zero(z)) acc = mul(acc, z);
        last = i;
    }
    F ai = inv(acc);
    constexpr int WU = Wire<F>::WORDS_UNCOMPRESSED, WC = Wire<F>::WORDS_COMPRESSED;
    for (size_t i = last;; i -= T) {
        F z = load_elem<F>(p.jz, i);
        bool inf = is_zero(z);
        F pre = load_elem<F>(p.prefix, i);
        F zinv = mul(ai, pre);
        if (!inf) ai = mul(ai, z);
        Aff<F> a;
        F zi2 = sqr(zinv);
        a.x = mul(load_elem<F>(p.jx, i), zi2);
        a.y = mul(load_elem<F>(p.jy, i), mul(zi2, zinv));
        if (inf && (p.flags & P2B_REJECT_INFINITY)) report(p.err, p.err_base + i, P2B_EINFINITY_OUT, 0);
        uint32_t o[WU];
        point_encode<F>(o, a, inf, p.out_enc);
        if (p.out_enc == ENC_COMPRESSED) store_words<WC>(p.out + i * WC, o);
        else store_words<WU>(p.out + i * WU, o);
        if (i < T) break;
    }
}
static constexpr int G1_BLOCK = P2B_G1_BLOCK;
static constexpr int G2_BLOCK = 128;

// stages: which parts of the pipeline this call queues (a probed G2 batch is PROLOGUE | KERNEL with the split kernel, then
// KERNEL | NORMALIZE with the exact one; the prologue's products -- decompressed points in c->misc, tau tables -- are reused)
enum { BM_PROLOGUE = 1, BM_KERNEL = 2, BM_NORMALIZE = 4, BM_ALL = 7 };
template <class F, int BLOCK, bool GLV, bool UNIFORM = false> int launch_typed(Ctx *c, const void *d_in, void *d_out, size_t n, const ScalarSpec &sc,
                                                       int in_enc, int out_enc, int flags, uint64_t err_base, int stages = BM_ALL,
                                                       const uint32_t *route = nullptr, uint32_t route_want = 0) {
    constexpr int W = FieldTraits<F>::WORDS;
    constexpr bool IS_G2 = W == 16;
    const size_t elem = (size_t)W * 4;
    int rc;
    if (sc.mode == 3) {                      // codec only
        RecodeParams rp{(const uint32_t *)d_in, (uint32_t *)d_out, n, in_enc, out_enc, flags, c->d_err, err_base};
        int blocks = (int)((n + 127) / 128);
        if (blocks > c->sm_count * 12) blocks = c->sm_count * 12;
        k_recode<F><<<blocks, 128, 0, c->stream>>>(rp);
        c->launches++;
        P2B_CUDA(c, cudaGetLastError());
        return P2B_OK;
    }
    if ((rc = dev_reserve(c, c->jac, 3 * n * elem))) return rc;
    if ((rc = dev_reserve(c, c->prefix, n * elem))) return rc;
    uint32_t *jx = (uint32_t *)c->jac.p, *jy = jx + n * W, *jz = jy + n * W;
    const uint32_t *in_words = (const uint32_t *)d_in;
    int kin_enc = in_enc;
    if (in_enc == P2B_ENC_COMPRESSED) {
        if ((rc = dev_reserve(c, c->misc, n * 2 * elem))) return rc;
        if (stages & BM_PROLOGUE) {
            DecompressParams dp{(const uint32_t *)d_in, (uint32_t *)c->misc.p, n, c->d_err, err_base};
            int blocks = (int)((n + 127) / 128);
            if (blocks > c->sm_count * 8) blocks = c->sm_count * 8;
            k_decompress<F><<<blocks, 128, 0, c->stream>>>(dp);
            c->launches++;
        }
        in_words = (const uint32_t *)c->misc.p;
        kin_enc = ENC_RAW_MONT_LE;
    }
    BatchMulParams bp;
    memset(&bp, 0, sizeof(bp));
    bp.in = in_words; bp.jx = jx; bp.jy = jy; bp.jz = jz; bp.n = n; bp.in_enc = kin_enc; bp.flags = flags;
    bp.sc_mode = sc.mode; bp.err = c->d_err; bp.err_base = err_base;
    bp.route = route; bp.route_want = route_want;
    if (sc.mode == 0) bp.scalars = (const uint32_t *)sc.d_scalars;
    else if (sc.mode == 1) {
        memcpy(bp.k, sc.k, 32);
        if (UNIFORM) bp.uni = uniform_digits(sc.k);
    }
    else {
        if ((rc = dev_reserve(c, c->tables, 4096 * sizeof(Fr) + 64))) return rc;
        if (stages & BM_PROLOGUE) {
            uint32_t *d_tc = (uint32_t *)((char *)c->tables.p + 4096 * sizeof(Fr));
            uint32_t tc[16];
            memcpy(tc, sc.tau, 32); memcpy(tc + 8, sc.coeff, 32);
            P2B_CUDA(c, cudaMemcpyAsync(d_tc, tc, 64, cudaMemcpyHostToDevice, c->stream));
            k_pow_tables<<<4096 / 128, 128, 0, c->stream>>>((Fr *)c->tables.p, d_tc, d_tc + 8);
            c->launches++;
        }
        bp.tables = (const Fr *)c->tables.p;
        bp.start = sc.start;
    }
    // odd-multiples table (G1) + double-buffered point / scalar tiles + 2 mbarriers
    const size_t smem = TablePolicy<F, BLOCK>::TABLE_BYTES + 2 * (size_t)BLOCK * (Wire<F>::WORDS_UNCOMPRESSED * 4 + 32) + 16;
    static bool attr_done[64] = {};          // function attributes are per device
    bool &attr_set = attr_done[c->device & 63];
    if (!attr_set && smem) {
        P2B_CUDA(c, cudaFuncSetAttribute(k_batch_mul<F, BLOCK, GLV, UNIFORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (TablePolicy<F, BLOCK>::TABLE_BYTES)
            P2B_CUDA(c, cudaFuncSetAttribute(k_batch_mul<F, BLOCK, GLV, UNIFORM>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                             cudaSharedmemCarveoutMaxShared));
        attr_set = true;
    }
    size_t ntiles = (n + BLOCK - 1) / BLOCK;
    const size_t max_grid = (size_t)c->sm_count * TablePolicy<F, BLOCK>::MIN_BLOCKS;
    int grid = (int)(ntiles < max_grid ? ntiles : max_grid);
    if (IS_G2) {
        if ((rc = dev_reserve(c, c->gtable, max_grid * BLOCK * 8 * 2 * W * 4))) return rc;
        bp.gtable = (uint4 *)c->gtable.p;
    }
    if (grid > 0 && (stages & BM_KERNEL)) {
        prof_begin(c, P2B_PROF_BATCH_MUL);
        k_batch_mul<F, BLOCK, GLV, UNIFORM><<<grid, BLOCK, smem, c->stream>>>(bp);
        prof_end(c, P2B_PROF_BATCH_MUL, 1);
        c->launches++;
    }
    if (grid > 0 && (stages & BM_NORMALIZE)) {
        // ~32 points per thread in the normalisation pass, at least one full wave of 128-thread blocks
        size_t threads = (n + 31) / 32;
        if (threads < (size_t)c->sm_count * 128) threads = n < (size_t)c->sm_count * 128 ? n : (size_t)c->sm_count * 128;
        int nblocks = (int)((threads + 127) / 128);
        NormalizeParams np{jx, jy, jz, (uint32_t *)c->prefix.p, (uint32_t *)d_out, n, out_enc, flags, c->d_err, err_base};
        prof_begin(c, P2B_PROF_NORMALIZE);
        k_normalize<F><<<nblocks, 128, 0, c->stream>>>(np);
        prof_end(c, P2B_PROF_NORMALIZE, 1);
        c->launches++;
    }
    P2B_CUDA(c, cudaGetLastError());
    return P2B_OK;
}


}  // namespace p2b
