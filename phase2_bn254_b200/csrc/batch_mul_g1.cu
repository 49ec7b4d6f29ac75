// batch_mul_g1.cu -- G1 instantiation of the batched scalar multiplication kernels + shared host helpers.
#define P2B_BATCH_MUL_COMMON
#include <cstdlib>
#include "batch_mul_impl.cuh"

namespace p2b {
// ------------------------------------------------------------------------------------------------- host side
size_t enc_size(int g2, int enc) {
    size_t full = g2 ? 128 : 64;
    return enc == P2B_ENC_COMPRESSED ? full / 2 : full;
}
int read_scalar_be(const uint8_t *be, uint32_t k[8]) {
    static const uint32_t r[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                                  0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    for (int i = 0; i < 8; i++) {
        const uint8_t *b = be + 4 * (7 - i);
        k[i] = ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | b[3];
    }
    for (int i = 7; i >= 0; i--) {
        if (k[i] < r[i]) return 1;
        if (k[i] > r[i]) return 0;
    }
    return 0;
}


int launch_batch_mul_g2(Ctx *c, const void *d_in, void *d_out, size_t n, const ScalarSpec &sc, int in_enc, int out_enc,
                        int flags, uint64_t err_index_base, int stages, const uint32_t *route, uint32_t route_want);

int launch_batch_mul_g2_glv(Ctx *c, const void *d_in, void *d_out, size_t n, const ScalarSpec &sc, int in_enc, int out_enc,
                            int flags, uint64_t err_index_base, int stages, const uint32_t *route, uint32_t route_want);

int launch_batch_mul_g1_uniform(Ctx *c, const void *d_in, void *d_out, size_t n, const ScalarSpec &sc, int in_enc, int out_enc,
                                int flags, uint64_t err_index_base);

// smallest G2 batch that is probed (below it the probe's fixed latency -- ~40 small launches and a 253-step ladder on 8 warps --
// exceeds what the split saves); P2B_G2_PROBE_MIN overrides (test hook), P2B_G2_PROBE=0 disables the probe
static size_t g2_probe_min() {
    if (const char *e = getenv("P2B_G2_PROBE")) if (atoi(e) == 0) return ~(size_t)0;
    if (const char *e = getenv("P2B_G2_PROBE_MIN")) { long v = atol(e); if (v > 0) return (size_t)v; }
    return (size_t)1 << 17;
}

int launch_batch_mul(Ctx *c, int g2, const void *d_in, void *d_out, size_t n, const ScalarSpec &sc, int in_enc,
                     int out_enc, int flags, uint64_t err_index_base) {
    if (n == 0) return P2B_OK;
    if (in_enc < 0 || in_enc > 2 || out_enc < 0 || out_enc > 2) return ctx_fail(c, P2B_EARG, "bad encoding");
    if (g2 && (flags & P2B_G2_SUBGROUP) && !(flags & P2B_G2_EXACT))
        return launch_batch_mul_g2_glv(c, d_in, d_out, n, sc, in_enc, out_enc, flags, err_index_base, BM_ALL, nullptr, 0);
    if (g2 && sc.mode != 3 && !(flags & P2B_G2_EXACT) && n >= g2_probe_min() && g2_probe_ready(c)) {
        // Large G2 batch without the caller's promise: prove subgroup membership of the whole batch on the device (msm_g2.cu),
        // queue the split kernel AND the exact kernel, and let the verdict word pick the one that runs.
        int rc;
        uint32_t *route = nullptr;
        if ((rc = launch_batch_mul_g2_glv(c, d_in, d_out, n, sc, in_enc, out_enc, flags, err_index_base, BM_PROLOGUE, nullptr, 0))) return rc;
        const bool comp = in_enc == P2B_ENC_COMPRESSED;           // the prologue left the decompressed points in c->misc
        if ((rc = g2_subgroup_probe(c, comp ? c->misc.p : d_in, n, comp ? P2B_ENC_RAW_MONT_LE : in_enc, err_index_base, &route))) return rc;
        if ((rc = launch_batch_mul_g2_glv(c, d_in, d_out, n, sc, in_enc, out_enc, flags, err_index_base, BM_KERNEL, route, 0))) return rc;
        return launch_batch_mul_g2(c, d_in, d_out, n, sc, in_enc, out_enc, flags, err_index_base, BM_KERNEL | BM_NORMALIZE, route, 1);
    }
    if (g2) return launch_batch_mul_g2(c, d_in, d_out, n, sc, in_enc, out_enc, flags, err_index_base, BM_ALL, nullptr, 0);
    if (sc.mode == 1 && n >= 1024) return launch_batch_mul_g1_uniform(c, d_in, d_out, n, sc, in_enc, out_enc, flags, err_index_base);
    return launch_typed<Fq, G1_BLOCK, true>(c, d_in, d_out, n, sc, in_enc, out_enc, flags, err_index_base);
}

}  // namespace p2b
