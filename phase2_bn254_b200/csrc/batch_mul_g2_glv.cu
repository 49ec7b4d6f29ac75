// batch_mul_g2_glv.cu -- G2 batched scalar multiplication with the endomorphism split: only valid for points of the order-r
// subgroup.  The reference decodes G2 without a subgroup check (pairing/src/bn256/ec.rs:1145-1213), so this kernel runs either
// when the caller vouches for membership (P2B_G2_SUBGROUP) or when the library has proven it for the whole batch
// (g2_subgroup_probe, msm_g2.cu: the launch is then gated by the probe's verdict word); the exact path is batch_mul_g2.cu.
#include "batch_mul_impl.cuh"

namespace p2b {

int launch_batch_mul_g2_glv(Ctx *c, const void *d_in, void *d_out, size_t n, const ScalarSpec &sc, int in_enc, int out_enc,
    int flags, uint64_t err_index_base, int stages, const uint32_t *route, uint32_t route_want) {
    return launch_typed<Fq2, G2_BLOCK, true>(c, d_in, d_out, n, sc, in_enc, out_enc, flags, err_index_base, stages, route, route_want);
}

}  // namespace p2b
