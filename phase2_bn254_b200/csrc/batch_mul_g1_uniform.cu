// batch_mul_g1_uniform.cu -- G1 batch_exp with ONE scalar for all points (MPCParameters::contribute,
// phase2/src/parameters.rs:424-470): warp-uniform width-5 NAF digits, see mul_glv_uniform in smul.cuh.
#include "batch_mul_impl.cuh"

namespace p2b {

int launch_batch_mul_g1_uniform(Ctx *c, const void *d_in, void *d_out, size_t n, const ScalarSpec &sc, int in_enc, int out_enc,
                                int flags, uint64_t err_index_base) {
    return launch_typed<Fq, G1_BLOCK, true, true>(c, d_in, d_out, n, sc, in_enc, out_enc, flags, err_index_base);
}

}  // namespace p2b
