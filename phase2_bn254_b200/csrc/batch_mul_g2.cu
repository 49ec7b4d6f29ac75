// batch_mul_g2.cu -- G2 (Fq2) instantiation of the batched scalar multiplication kernels.
#include "batch_mul_impl.cuh"

namespace p2b {

int launch_batch_mul_g2(Ctx *c, const void *d_in, void *d_out, size_t n, const ScalarSpec &sc, int in_enc, int out_enc,
    int flags, uint64_t err_index_base, int stages, const uint32_t *route, uint32_t route_want) {
    return launch_typed<Fq2, G2_BLOCK, false>(c, d_in, d_out, n, sc, in_enc, out_enc, flags, err_index_base, stages, route, route_want);
}

}  // namespace p2b
