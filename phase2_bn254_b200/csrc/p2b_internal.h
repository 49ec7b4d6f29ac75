// p2b_internal.h -- context and launcher prototypes shared by the .cu translation units of libp2b.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <utility>
#include <vector>
#include "../../include/p2b.h"

namespace p2b {

// device-side error word: min over (index << 8 | kind << 4 | sub); ~0 = no error
static constexpr unsigned long long ERR_NONE = ~0ull;

struct HostIO;
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct Ctx {
    int device = 0;
    cudaStream_t stream = nullptr;       // compute
    cudaStream_t copy_in = nullptr;      // H2D
    cudaStream_t copy_out = nullptr;     // D2H
    cudaStream_t sort_stream = nullptr;  // MSM counting sort (high priority), overlapped with the bucket accumulation
    cudaEvent_t msm_ev[4] = {};
    cudaEvent_t h2d_ev[2] = {};          // timing events around the copies of one streamed MSM chunk (link-speed probe)
    size_t h2d_probe_bytes = 0;
    double h2d_gbps = 0;                 // host -> device rate seen by the previous streamed call (0: not measured yet)
    int sm_count = 148;
    std::string last_error;
    uint64_t err_index = 0;
    int err_sub = 0;
    uint64_t launches = 0;
    unsigned long long *d_err = nullptr;   // device error word
    unsigned long long *h_err = nullptr;   // pinned mirror
    // growable device scratch
    DevBuf jac, prefix, stage_in[2], stage_out[2], scal, tables, misc, msm_a, msm_b, msm_c, msm_d, msm_e, msm_f, fft_tw_dir[2], gtable, gfft, probe;
    // per-window sums of the last MSM (device, XYZZ): read by the G2 subgroup probe (msm_g2.cu)
    uint32_t *msm_last_wsum = nullptr;
    uint32_t msm_last_nwin = 0;
    // G2 subgroup probe (g2_subgroup_probe): ChaCha20 key from the host CSPRNG (drawn at first use), running coefficient index
    uint32_t probe_key[8] = {};
    bool probe_key_set = false;
    uint64_t probe_ctr = 0;
    uint64_t probes = 0;                 // probes launched so far (p2b_g2_probe_count)
    // pinned staging rings + copy threads for pageable caller buffers (hostio.cu), created on first use
    HostIO *io = nullptr;
    cudaEvent_t ev[8] = {};
    // profiling (p2b_profile_enable): CUDA-event brackets around the dominant kernels, on `stream`
    bool prof = false;
    struct ProfSlot {
        std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
        size_t used = 0;
        uint64_t kernels = 0;
    } prof_slot[P2B_PROF_SLOTS];
    // cached FFT twiddle tables, one set per direction (fft_tw_dir[0] forward, [1] inverse)
    uint32_t fft_tw_dir_log_n[2] = {0xffffffffu, 0xffffffffu};
};

int ctx_fail(Ctx *c, int code, const std::string &msg);
int ctx_cuda(Ctx *c, cudaError_t e, const char *what);
int dev_reserve(Ctx *c, DevBuf &b, size_t bytes);
// reads the device error word back (after the stream has drained) and converts it to a P2B_* code
int ctx_collect_error(Ctx *c);

// brackets `kernels` launches queued between prof_begin / prof_end with timing events (no-ops unless profiling)
void prof_begin(Ctx *c, int slot, cudaStream_t stream = nullptr);   // nullptr: c->stream
void prof_end(Ctx *c, int slot, int kernels, cudaStream_t stream = nullptr);

// ---- hostio.cu: copies between CALLER buffers (pinned or pageable, e.g. the mmaps of the reference's binaries) and the device
int io_h2d(Ctx *c, void *d_dst, const void *h_src, size_t bytes, cudaStream_t s);
int io_d2h(Ctx *c, void *h_dst, const void *d_src, size_t bytes, cudaStream_t s);
int io_flush(Ctx *c);        // all staged D2H bytes have reached the caller's buffer (after the streams have drained)
void io_destroy(Ctx *c);
void io_stats(Ctx *c, uint64_t *staged_in, uint64_t *staged_out);

// NVTX range around every ABI call (header-only nvtx3: a no-op unless a profiler injects itself) -- the counterpart of the
// reference's only tracing, the per-step log lines of bellman/src/log.rs:56-68 and the bins' progress prints
struct NvtxRange {
    explicit NvtxRange(const char *name);
    ~NvtxRange();
};
#define P2B_RANGE(name) ::p2b::NvtxRange nvtx_range__(name)

#define P2B_CUDA(c, call)                                     \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) return ctx_cuda((c), e__, #call); \
    } while (0)

// ---- batch_mul.cu ----
struct ScalarSpec {
    int mode;                     // 0 = array (device, 32 B BE each), 1 = broadcast, 2 = powers of tau, 3 = none (codec only)
    const void *d_scalars;        // mode 0
    uint32_t k[8];                // mode 1: canonical little-endian limbs
    uint32_t tau[8], coeff[8];    // mode 2: canonical limbs (coeff = 1 when absent)
    uint64_t start;               // mode 2
};
// d_in / d_out: device buffers holding n encodings.  Work is queued on c->stream; errors land in c->d_err.
int launch_batch_mul(Ctx *c, int g2, const void *d_in, void *d_out, size_t n, const ScalarSpec &sc, int in_enc,
                     int out_enc, int flags, uint64_t err_index_base);
// Queues a proof that all n G2 points (device; uncompressed wire or RAW_MONT_LE) lie in the order-r subgroup: *d_route is a
// device word that reads 0 afterwards if they do, non-zero if the batch has to take the exact path (msm_g2.cu).
bool g2_probe_ready(Ctx *c);     // the probe's coefficient key could be drawn from the host CSPRNG
int g2_subgroup_probe(Ctx *c, const void *d_points, size_t n, int enc, uint64_t err_base, uint32_t **d_route);
size_t enc_size(int g2, int enc);
int read_scalar_be(const uint8_t *be, uint32_t k[8]);   // false if >= r

// ---- msm.cu ----
int launch_msm(Ctx *c, int g2, const void *d_points, const void *d_scalars, size_t n, void *d_result_jac);
int msm_finish_affine(Ctx *c, int g2, const void *d_jac, int count, uint8_t *out_host);

// ---- fft.cu ----
int launch_fr_fft(Ctx *c, void *d_data, uint32_t log_n, int inverse, int coset);

}  // namespace p2b

struct p2b_ctx {
    p2b::Ctx c;
};
