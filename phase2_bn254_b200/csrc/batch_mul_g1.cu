// batch_mul_g1.cu -- G1 instantiation of the batched scalar multiplication kernels + shared host helpers.
#define P2B_BATCH_MUL_COMMON
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
                        int flags, uint64_t err_index_base);

int launch_batch_mul_g2_glv(Ctx *c, const void *d_in, void *d_out, size_t n, const ScalarSpec &sc, int in_enc, int out_enc,
                            int flags, uint64_t err_index_base);

int launch_batch_mul_g1_uniform(Ctx *c, const void *d_in, void *d_out, size_t n, const ScalarSpec &sc, int in_enc, int out_enc,
                                int flags, uint64_t err_index_base);

int launch_batch_mul(Ctx *c, int g2, const void *d_in, void *d_out, size_t n, const ScalarSpec &sc, int in_enc,
                     int out_enc, int flags, uint64_t err_index_base) {
    if (n == 0) return P2B_OK;
    if (in_enc < 0 || in_enc > 2 || out_enc < 0 || out_enc > 2) return ctx_fail(c, P2B_EARG, "bad encoding");
    if (g2 && (flags & P2B_G2_SUBGROUP)) return launch_batch_mul_g2_glv(c, d_in, d_out, n, sc, in_enc, out_enc, flags, err_index_base);
    if (g2) return launch_batch_mul_g2(c, d_in, d_out, n, sc, in_enc, out_enc, flags, err_index_base);
    if (sc.mode == 1 && n >= 1024) return launch_batch_mul_g1_uniform(c, d_in, d_out, n, sc, in_enc, out_enc, flags, err_index_base);
    return launch_typed<Fq, G1_BLOCK, true>(c, d_in, d_out, n, sc, in_enc, out_enc, flags, err_index_base);
}

}  // namespace p2b
