// batch_mul_g2_glv.cu -- G2 batched scalar multiplication with the endomorphism split (P2B_G2_SUBGROUP): only valid for
// points of the order-r subgroup, which the caller vouches for (the reference decodes G2 without a subgroup check,
// pairing/src/bn256/ec.rs:1145-1213, so this is opt-in; the default path is batch_mul_g2.cu).
#include "batch_mul_impl.cuh"

namespace p2b {

int launch_batch_mul_g2_glv(Ctx *c, const void *d_in, void *d_out, size_t n, const ScalarSpec &sc, int in_enc, int out_enc,
                            int flags, uint64_t err_index_base) {
    return launch_typed<Fq2, G2_BLOCK, true>(c, d_in, d_out, n, sc, in_enc, out_enc, flags, err_index_base);
}

}  // namespace p2b
