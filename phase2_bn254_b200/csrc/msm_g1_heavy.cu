// msm_g1_heavy.cu -- G1 instantiation of the heavy-bucket kernels of the Pippenger MSM (k_msm_heavy, k_msm_heavy_combine).
#include "msm_impl.cuh"

namespace p2b {
void msm_launch_heavy_g1(Ctx *c, const uint32_t *aff, const uint32_t *sorted, uint32_t *buckets, const MsmHeavy &hv) {
    msm_launch_heavy_impl<Fq>(c, aff, sorted, buckets, hv);
}
}  // namespace p2b
