// sparse.cu -- sparse linear maps over group elements:  out[i] = sum_{j in row i} coeff[j] * base[col[j]].
//
// This is the QAP evaluation of MPCParameters::new (phase2/src/parameters.rs:225-300): for every variable of the circuit,
// a_g1 / b_g1 / b_g2 / ext are linear combinations of the Lagrange-basis points of phase1radix2m{m} with the (coeff,
// constraint) entries of the A / B / C matrices.  The reference walks the rows on CPU threads with one wNAF scalar
// multiplication per entry; here
//   k_sparse_gather   base[col[j]] -> one contiguous array (16-byte granules, coalesced stores)
//   k_batch_mul       the library's batched scalar multiplication, one scalar per entry (batch_mul_impl.cuh)
//   k_segment_sum     one thread per segment of <= SEG entries: complete mixed additions into a Jacobian sum
//   k_normalize       Montgomery-trick inversion + encoding
// Entries whose coefficient is 1 or r - 1 (the bulk of a QAP) skip the scalar multiplication: the segment sum adds / subtracts
// the base point itself.  Rows longer than SEG are cut into segments that a second (third, ...) k_segment_sum level combines, so a variable that
// appears in every constraint (the constant ONE) does not serialise on one thread.
#include <vector>
#include "batch_mul_impl.cuh"

namespace p2b {

static constexpr uint64_t SPARSE_SEG = 256;

template <class F> __global__ void __launch_bounds__(256) k_sparse_gather(const uint4 *bases, const uint32_t *cols, size_t nnz, uint4 *out) {
    constexpr int G = Wire<F>::WORDS_UNCOMPRESSED / 4;            // 16-byte granules per point
    const size_t total = nnz * G;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const size_t j = t / G;
        const int g = (int)(t % G);
        out[t] = __ldg(bases + (size_t)cols[j] * G + g);
    }
}

// pts: affine points as Montgomery limbs (ENC_RAW_MONT_LE, all-zero = infinity); segment s = entries [off[s], off[s+1])
template <class F> __global__ void __launch_bounds__(128) k_segment_sum(const uint32_t *pts, const uint64_t *off, size_t nseg, uint32_t *jx,
                                                                        uint32_t *jy, uint32_t *jz) {
    constexpr int WU = Wire<F>::WORDS_UNCOMPRESSED, W = FieldTraits<F>::WORDS;
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    Jac<F> acc = jac_infinity<F>();
    const uint64_t lo = off[s], hi = off[s + 1];
#pragma unroll 1
    for (uint64_t j = lo; j < hi; j++) {
        uint32_t w[WU];
        load_words<WU>(w, pts + j * WU);
        uint32_t any = 0;
#pragma unroll
        for (int k = 0; k < WU; k++) any |= w[k];
        if (!any) continue;
        Aff<F> q;
#pragma unroll
        for (int k = 0; k < W; k++) { set_word(q.x, k, w[k]); set_word(q.y, k, w[W + k]); }
        acc = jac_madd(acc, q);
    }
    store_elem<F>(jx, s, acc.x);
    store_elem<F>(jy, s, acc.y);
    store_elem<F>(jz, s, acc.z);
}

// First level: entry j is either the scaled point scaled[src[j]] (kind 0) or +- the base point base_raw[src[j]] (kind 1 / 2).
template <class F> __global__ void __launch_bounds__(128) k_segment_sum_entries(const uint32_t *scaled, const uint32_t *base_raw, const uint32_t *src,
                                                                                const uint8_t *kind, const uint64_t *off, size_t nseg,
                                                                                uint32_t *jx, uint32_t *jy, uint32_t *jz) {
    constexpr int WU = Wire<F>::WORDS_UNCOMPRESSED, W = FieldTraits<F>::WORDS;
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    Jac<F> acc = jac_infinity<F>();
    const uint64_t lo = off[s], hi = off[s + 1];
#pragma unroll 1
    for (uint64_t j = lo; j < hi; j++) {
        const uint32_t kd = kind[j];
        uint32_t w[WU];
        load_words<WU>(w, (kd ? base_raw : scaled) + (size_t)src[j] * WU);
        uint32_t any = 0;
#pragma unroll
        for (int k = 0; k < WU; k++) any |= w[k];
        if (!any) continue;
        Aff<F> q;
#pragma unroll
        for (int k = 0; k < W; k++) { set_word(q.x, k, w[k]); set_word(q.y, k, w[W + k]); }
        q.y = cneg(q.y, kd == 2);
        acc = jac_madd(acc, q);
    }
    store_elem<F>(jx, s, acc.x);
    store_elem<F>(jy, s, acc.y);
    store_elem<F>(jz, s, acc.z);
}

template <class F> static int sparse_mul_host(Ctx *c, int g2, const uint8_t *bases, size_t n_bases, const uint64_t *row_offsets, const uint32_t *cols,
                                              const uint8_t *coeffs, size_t n_rows, uint8_t *out) {
    constexpr int W = FieldTraits<F>::WORDS, WU = Wire<F>::WORDS_UNCOMPRESSED;
    if (!row_offsets || (n_rows && !out)) return ctx_fail(c, P2B_EARG, "null buffer");
    if (row_offsets[0] != 0) return ctx_fail(c, P2B_EARG, "sparse: row_offsets[0] must be 0");
    for (size_t i = 0; i < n_rows; i++)
        if (row_offsets[i + 1] < row_offsets[i]) return ctx_fail(c, P2B_EARG, "sparse: row_offsets must not decrease");
    const uint64_t nnz = row_offsets[n_rows];
    if (nnz && (!bases || !cols || !coeffs)) return ctx_fail(c, P2B_EARG, "null buffer");
    if (nnz >= (1ull << 31) || n_rows >= (1ull << 31) || n_bases >= (1ull << 32)) return ctx_fail(c, P2B_EARG, "sparse: too many entries for one call");
    for (uint64_t j = 0; j < nnz; j++)
        if (cols[j] >= n_bases) {
            c->err_index = j;
            return ctx_fail(c, P2B_EARG, "sparse: column index out of range");
        }
    if (n_rows == 0) return P2B_OK;
    // classify the entries: coefficient 1 -> +P, r - 1 -> -P (no scalar multiplication), anything else -> general
    static const uint8_t ONE_BE[32] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1};
    uint8_t minus_one_be[32];
    for (int i = 0; i < 8; i++) {
        uint32_t v = FrP::p(7 - i) - (i == 7 ? 1u : 0u);         // r - 1: r is odd, no borrow
        minus_one_be[4 * i] = (uint8_t)(v >> 24); minus_one_be[4 * i + 1] = (uint8_t)(v >> 16);
        minus_one_be[4 * i + 2] = (uint8_t)(v >> 8); minus_one_be[4 * i + 3] = (uint8_t)v;
    }
    // (first a counting pass: when no entry is trivial the caller's arrays are used as they are)
    uint64_t ntriv = 0;
    for (uint64_t j = 0; j < nnz; j++) {
        const uint8_t *k = coeffs + 32 * j;
        ntriv += (!memcmp(k, ONE_BE, 32) || !memcmp(k, minus_one_be, 32)) ? 1 : 0;
    }
    const bool all_general = ntriv == 0;
    std::vector<uint8_t> kind, gen_coeffs;
    std::vector<uint32_t> src, gen_cols;
    if (!all_general) {
        kind.resize(nnz);
        src.resize(nnz);
        gen_cols.resize(nnz - ntriv);
        gen_coeffs.resize((nnz - ntriv) * 32);
        uint64_t ng = 0;
        for (uint64_t j = 0; j < nnz; j++) {
            const uint8_t *k = coeffs + 32 * j;
            if (!memcmp(k, ONE_BE, 32)) { kind[j] = 1; src[j] = cols[j]; }
            else if (!memcmp(k, minus_one_be, 32)) { kind[j] = 2; src[j] = cols[j]; }
            else {
                kind[j] = 0;
                src[j] = (uint32_t)ng;
                gen_cols[ng] = cols[j];
                memcpy(&gen_coeffs[32 * ng], k, 32);
                ng++;
            }
        }
    }
    const uint32_t *h_gen_cols = all_general ? cols : gen_cols.data();
    const uint8_t *h_gen_coeffs = all_general ? coeffs : gen_coeffs.data();
    const uint64_t ngen = nnz - ntriv;
    P2B_CUDA(c, cudaSetDevice(c->device));
    c->last_error.clear();
    P2B_CUDA(c, cudaMemsetAsync(c->d_err, 0xff, sizeof(unsigned long long), c->stream));
    const size_t psz = (size_t)WU * 4, elem = (size_t)W * 4;
    int rc;
    // stage_in[0]: bases | general cols | general coeffs | src | kind ; gfft: A = gathered / scaled general entries, B = segment
    // results, R = the bases as Montgomery affine points ; stage_in[1]: segment offsets
    const size_t cols_off = (n_bases * psz + 255) & ~(size_t)255, coef_off = (cols_off + ngen * 4 + 255) & ~(size_t)255;
    const size_t src_off = (coef_off + ngen * 32 + 255) & ~(size_t)255, kind_off = (src_off + nnz * 4 + 255) & ~(size_t)255;
    if ((rc = dev_reserve(c, c->stage_in[0], kind_off + nnz + 256))) return rc;
    const uint64_t nseg_max = nnz / SPARSE_SEG + n_rows + 1;
    const size_t a_pts = ngen > nseg_max ? ngen : nseg_max;          // A is reused for segment results from the third level on
    const size_t b_off = (a_pts * psz + 255) & ~(size_t)255, r_off = (b_off + nseg_max * psz + 255) & ~(size_t)255;
    if ((rc = dev_reserve(c, c->gfft, r_off + n_bases * psz + 256))) return rc;
    if ((rc = dev_reserve(c, c->stage_in[1], (nseg_max + 1) * 8))) return rc;
    if ((rc = dev_reserve(c, c->stage_out[0], n_rows * psz))) return rc;
    char *sin = (char *)c->stage_in[0].p;
    uint32_t *A = (uint32_t *)c->gfft.p, *B = (uint32_t *)((char *)c->gfft.p + b_off), *R = (uint32_t *)((char *)c->gfft.p + r_off);
    if (nnz) {
        if ((rc = io_h2d(c, sin, bases, n_bases * psz, c->stream))) return rc;
        if (!all_general) {
            P2B_CUDA(c, cudaMemcpyAsync(sin + src_off, src.data(), nnz * 4, cudaMemcpyHostToDevice, c->stream));
            P2B_CUDA(c, cudaMemcpyAsync(sin + kind_off, kind.data(), nnz, cudaMemcpyHostToDevice, c->stream));
        }
        ScalarSpec sc;
        memset(&sc, 0, sizeof sc);
        if (ngen < nnz) {        // +-P entries read the bases as Montgomery affine points: one checked codec pass over the bases
            sc.mode = 3;
            if ((rc = launch_batch_mul(c, g2, sin, R, n_bases, sc, P2B_ENC_UNCOMPRESSED, P2B_ENC_RAW_MONT_LE, P2B_CHECK_INPUT, 0))) return rc;
        }
        if (ngen) {
            P2B_CUDA(c, cudaMemcpyAsync(sin + cols_off, h_gen_cols, ngen * 4, cudaMemcpyHostToDevice, c->stream));
            P2B_CUDA(c, cudaMemcpyAsync(sin + coef_off, h_gen_coeffs, ngen * 32, cudaMemcpyHostToDevice, c->stream));
            size_t blocks = (ngen * (WU / 4) + 255) / 256;
            if (blocks > (size_t)c->sm_count * 16) blocks = (size_t)c->sm_count * 16;
            k_sparse_gather<F><<<(int)blocks, 256, 0, c->stream>>>((const uint4 *)sin, (const uint32_t *)(sin + cols_off), ngen, (uint4 *)A);
            c->launches++;
            sc.mode = 0;
            sc.d_scalars = sin + coef_off;
            // checked decode of the gathered points; a zero coefficient gives infinity (all-zero raw encoding), which the sum skips
            if ((rc = launch_batch_mul(c, g2, A, A, ngen, sc, P2B_ENC_UNCOMPRESSED, P2B_ENC_RAW_MONT_LE, P2B_CHECK_INPUT, 0))) return rc;
        }
    }
    // levels of segment sums
    std::vector<uint64_t> cur(row_offsets, row_offsets + n_rows + 1), seg, first;
    const uint32_t *cur_pts = A;
    uint32_t *next_pts = B;
    for (int level = 0;; level++) {
        seg.clear();
        first.resize(n_rows + 1);
        for (size_t i = 0; i < n_rows; i++) {
            first[i] = seg.size();
            uint64_t lo = cur[i];
            const uint64_t hi = cur[i + 1];
            do {
                seg.push_back(lo);
                lo = hi - lo > SPARSE_SEG ? lo + SPARSE_SEG : hi;
            } while (lo < hi);
        }
        first[n_rows] = seg.size();
        const size_t nseg = seg.size();
        seg.push_back(cur[n_rows]);
        const bool last = nseg == n_rows;
        // the previous level's kernels may still be reading stage_in[1]: the copy is stream-ordered behind them, but the
        // host vector must outlive it
        P2B_CUDA(c, cudaMemcpyAsync(c->stage_in[1].p, seg.data(), (nseg + 1) * 8, cudaMemcpyHostToDevice, c->stream));
        if ((rc = dev_reserve(c, c->jac, 3 * nseg * elem))) return rc;
        if ((rc = dev_reserve(c, c->prefix, nseg * elem))) return rc;
        uint32_t *jx = (uint32_t *)c->jac.p, *jy = jx + nseg * W, *jz = jy + nseg * W;
        if (level == 0 && !all_general)
            k_segment_sum_entries<F><<<(int)((nseg + 127) / 128), 128, 0, c->stream>>>(A, R, (const uint32_t *)(sin + src_off), (const uint8_t *)(sin + kind_off),
                                                                                    (const uint64_t *)c->stage_in[1].p, nseg, jx, jy, jz);
        else
            k_segment_sum<F><<<(int)((nseg + 127) / 128), 128, 0, c->stream>>>(cur_pts, (const uint64_t *)c->stage_in[1].p, nseg, jx, jy, jz);
        size_t threads = (nseg + 31) / 32;
        if (threads < (size_t)c->sm_count * 128) threads = nseg < (size_t)c->sm_count * 128 ? nseg : (size_t)c->sm_count * 128;
        NormalizeParams np{jx, jy, jz, (uint32_t *)c->prefix.p, last ? (uint32_t *)c->stage_out[0].p : next_pts, nseg,
                           last ? (int)ENC_UNCOMPRESSED : (int)ENC_RAW_MONT_LE, 0, c->d_err, 0};
        k_normalize<F><<<(int)((threads + 127) / 128), 128, 0, c->stream>>>(np);
        c->launches += 2;
        P2B_CUDA(c, cudaStreamSynchronize(c->stream));       // `seg` is rebuilt next round
        if (last) break;
        cur.swap(first);
        cur_pts = next_pts;
        next_pts = next_pts == B ? A : B;
    }
    if ((rc = io_d2h(c, out, c->stage_out[0].p, n_rows * psz, c->stream))) return rc;
    P2B_CUDA(c, cudaGetLastError());
    return ctx_collect_error(c);
}

}  // namespace p2b

using namespace p2b;
extern "C" {
int p2b_g1_sparse_mul(p2b_ctx *h, const uint8_t *bases, size_t n_bases, const uint64_t *row_offsets, const uint32_t *cols,
                      const uint8_t *coeffs_be32, size_t n_rows, uint8_t *out) { P2B_RANGE("p2b_g1_sparse_mul");
    return h ? sparse_mul_host<Fq>(&h->c, 0, bases, n_bases, row_offsets, cols, coeffs_be32, n_rows, out) : P2B_EARG;
}
int p2b_g2_sparse_mul(p2b_ctx *h, const uint8_t *bases, size_t n_bases, const uint64_t *row_offsets, const uint32_t *cols,
                      const uint8_t *coeffs_be32, size_t n_rows, uint8_t *out) { P2B_RANGE("p2b_g2_sparse_mul");
    return h ? sparse_mul_host<Fq2>(&h->c, 1, bases, n_bases, row_offsets, cols, coeffs_be32, n_rows, out) : P2B_EARG;
}
}
