// msm_g2.cu -- G2 (Fq2) instantiation of the Pippenger MSM kernels.
#include "msm_impl.cuh"

namespace p2b {
int msm_typed_g2(Ctx *c, const void *d_points, const void *d_scalars, size_t n, uint32_t *d_out_wire, size_t geom_n,
                 int phase, uint64_t err_base, size_t total_n) {
    return msm_typed<Fq2>(c, d_points, d_scalars, n, d_out_wire, geom_n, phase, err_base, total_n);
}
}  // namespace p2b
