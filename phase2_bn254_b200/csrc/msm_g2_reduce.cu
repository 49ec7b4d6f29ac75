// msm_g2_reduce.cu -- G2 instantiation of the bucket / window reduction kernels of the Pippenger MSM (k_msm_reduce1,
// k_msm_reduce2 with its warp-shuffle scans, k_msm_final) and of k_sum_points.
#include "msm_impl.cuh"

namespace p2b {
void msm_launch_reduce_g2(Ctx *c, const uint32_t *buckets, const MsmGeom &g, uint32_t *s1, uint32_t *s2, uint32_t *wsum, uint32_t *d_out_wire,
                          size_t nred) {
    msm_launch_reduce_impl<Fq2>(c, buckets, g, s1, s2, wsum, d_out_wire, nred);
}
void msm_launch_sum_points_g2(Ctx *c, const uint32_t *d_in, uint32_t count, uint32_t *d_out) {
    k_sum_points<Fq2><<<1, 32, 0, c->stream>>>(d_in, count, d_out, c->d_err);
}
void msm_launch_probe_order_g2(Ctx *c, const uint32_t *wsum, uint32_t nwin, uint32_t *route) {
    k_msm_probe_order<Fq2><<<nwin, 32, 0, c->stream>>>(wsum, nwin, route);
}
}  // namespace p2b
