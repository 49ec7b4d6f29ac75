// msm_g2.cu -- G2 (Fq2) instantiation of the Pippenger MSM kernels.
#include "msm_impl.cuh"

namespace p2b {
int msm_typed_g2(Ctx *c, const MsmJob &j) { return msm_typed<Fq2>(c, j); }
}  // namespace p2b
