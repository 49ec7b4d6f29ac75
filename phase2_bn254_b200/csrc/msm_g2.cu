// msm_g2.cu -- G2 (Fq2) instantiation of the Pippenger MSM kernels, and the batch subgroup probe built on them.
#include <cstdio>
#include <random>
#include <sys/random.h>
#include "msm_impl.cuh"

namespace p2b {
int msm_typed_g2(Ctx *c, const MsmJob &j) { return msm_typed<Fq2>(c, j); }

// ------------------------------------------------------------------------------------------------- G2 subgroup probe
// The reference decodes G2 points without a subgroup check (pairing/src/bn256/ec.rs:1145-1213) and multiplies whatever curve
// point it is given, so the endomorphism split of `batch_exp` (smul.cuh, 1.45x fewer field multiplications) may only be used
// for points known to lie in the order-r subgroup G2 of E'(Fq2).  A per-point membership test costs a quarter of the scalar
// multiplication it would save; a test of the WHOLE BATCH costs ~3 %:
//
//   E'(Fq2) = G2 x H with #H = h = 2q - r = 10069 * 5864401 * 1875725156269 * p177 (tests/test_oracle.py), gcd(r, h) = 1.
//   For coefficients rho_i the sum W = sum rho_i P_i has H-component sum rho_i h_i (h_i = H-component of P_i), and W lies in
//   G2 iff [r] W = O.  If every P_i is in G2, so is W.  If some h_j != 0, its order is at least the smallest prime factor of
//   h, 10069, and for rho_j uniform over >= 2^13 consecutive integers -- independently of everything else in the sum --
//   rho_j h_j hits the one value that cancels the rest with probability <= 2^-13.
//
// One bucket MSM over the batch with random (8c - 1)-bit coefficients yields 8 such sums at once: its window sums
// W_w = sum_i d_w(rho_i) P_i, whose signed digits d_w(rho_j) are, given the lower windows, uniform over 2^c (top window:
// 2^(c-1)) consecutive integers, c = 14 .. 16.  A point outside G2 therefore survives all 8 tests with probability
// <= 2^-104; points that are not on the curve at all (only possible without P2B_CHECK_INPUT) are caught deterministically by
// the decode kernel, because sums of off-curve points are not group sums.  The coefficients come from ChaCha20 under a key
// drawn from the kernel's CSPRNG (getrandom) when the context first needs it; the verdict stays on the device (a word that the two
// `k_batch_mul` launches of the caller read), so the chunk pipeline never waits for the host.
// Draws the ChaCha20 key of the probe from the host's CSPRNG (once per context).  false: no entropy source -- the caller then
// takes the exact path, which needs no randomness.
bool g2_probe_ready(Ctx *c) {
    if (c->probe_key_set) return true;
    if (getrandom(c->probe_key, sizeof c->probe_key, 0) != (ssize_t)sizeof c->probe_key) {     // the kernel's CSPRNG
        try {
            std::random_device rd;                  // fallback: /dev/urandom through libstdc++
            if (rd.entropy() == 0.0) return false;  // a deterministic engine is no source of unpredictable coefficients
            for (int i = 0; i < 8; i++) c->probe_key[i] = rd();
        } catch (...) {
            return false;
        }
    }
    c->probe_key_set = true;
    return true;
}

int g2_subgroup_probe(Ctx *c, const void *d_points, size_t n, int enc, uint64_t err_base, uint32_t **d_route) {
    if (!g2_probe_ready(c)) return ctx_fail(c, P2B_EARG, "g2 subgroup probe: no entropy source");
    uint32_t lg = 0;
    while (((size_t)1 << (lg + 1)) <= n) lg++;
    int cw = (int)(0.6 * lg + 3.5);                 // the MSM's own window width for n terms (msm_geometry), kept within 14 .. 16
    if (cw < 14) cw = 14;
    if (cw > 16) cw = 16;
    const uint32_t bits = 8u * (uint32_t)cw - 1u;   // 8 windows of cw bits cover bits + 1
    int rc;
    if ((rc = dev_reserve(c, c->probe, 256 + n * 32))) return rc;
    uint32_t *route = (uint32_t *)c->probe.p, *scalars = route + 64;
    P2B_CUDA(c, cudaMemsetAsync(route, 0, 4, c->stream));
    {
        ChaKey key;
        for (int i = 0; i < 8; i++) key.k[i] = c->probe_key[i];
        int grid = (int)((n / 2 + 256) / 256);
        if (grid > c->sm_count * 16) grid = c->sm_count * 16;
        k_msm_random_scalars<<<grid, 256, 0, c->stream>>>(scalars, n, c->probe_ctr, key, bits);
        c->launches++;
        c->probe_ctr += n + (n & 1);                // whole keystream blocks: no coefficient is ever reused under this key
    }
    MsmJob j;
    j.d_points = d_points;
    j.d_scalars = scalars;
    j.n = n;
    j.geom_n = n;
    j.d_out_wire = nullptr;                         // window sums only
    j.err_base = err_base;
    j.in_enc = enc;
    j.check = 4;                                    // off-curve points raise the route word
    j.scalar_bits = bits;
    j.d_route = route;
    if ((rc = msm_typed<Fq2>(c, j))) return rc;
    msm_launch_probe_order_g2(c, c->msm_last_wsum, c->msm_last_nwin, route);
    c->launches++;
    c->probes++;
    P2B_CUDA(c, cudaGetLastError());
    *d_route = route;
    return P2B_OK;
}
}  // namespace p2b
