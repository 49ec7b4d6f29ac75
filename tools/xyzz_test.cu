// standalone device check of the warp-shuffle reductions used by k_msm_reduce2
#include "../phase2_bn254_b200/csrc/msm_impl.cuh"
using namespace p2b;
__device__ void to_wire(const Xyzz<Fq>& acc, uint32_t* out) {
    bool inf = is_zero(acc.zz);
    Fq t = inv(mul(acc.zz, acc.zzz));
    Aff<Fq> a; a.x = mul(acc.x, mul(t, acc.zzz)); a.y = mul(acc.y, mul(t, acc.zz));
    uint32_t o[16]; point_encode<Fq>(o, a, inf, ENC_UNCOMPRESSED);
    for (int j = 0; j < 16; j++) out[j] = o[j];
}
__global__ void k_test(uint32_t* out, int mode) {
    int lane = threadIdx.x & 31;
    Aff<Fq> g; g.x = fp_one<FqP>(); g.y = dbl(fp_one<FqP>());
    Xyzz<Fq> G; G.x = g.x; G.y = g.y; G.zz = fp_one<FqP>(); G.zzz = fp_one<FqP>();
    Xyzz<Fq> v = (mode == 0 ? lane == 0 : true) ? G : xyzz_infinity<Fq>();
    Xyzz<Fq> s = warp_sum(v);
    Xyzz<Fq> w = warp_weighted_sum(v);
    Xyzz<Fq> d = xdbl(G);
    Xyzz<Fq> a = xadd(G, xyzz_infinity<Fq>());
    Xyzz<Fq> b = xadd(xyzz_infinity<Fq>(), G);
    Xyzz<Fq> c = xyzz_infinity<Fq>();
    for (int i = 0; i < 5; i++) c = xdbl(c);
    Xyzz<Fq> e = xadd(c, xyzz_infinity<Fq>());
    Xyzz<Fq> f = xadd(G, e);
    if (lane == 0) { to_wire(s, out); to_wire(w, out + 16); to_wire(d, out + 32); to_wire(a, out + 48); to_wire(b, out + 64); to_wire(c, out + 80); to_wire(e, out + 96); to_wire(f, out + 112); }
}
int main() {
    uint32_t* d; cudaMalloc(&d, 4096);
    for (int mode = 0; mode < 2; mode++) {
        k_test<<<1, 32>>>(d, mode);
        uint8_t h[8 * 64]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
        printf("mode %d err=%s\n", mode, cudaGetErrorString(cudaGetLastError()));
        const char* names[8] = {"warp_sum", "warp_wsum", "dbl(G)", "G+inf", "inf+G", "dbl^5(inf)", "that+inf", "G+that"};
        for (int k = 0; k < 8; k++) { printf(" %-10s ", names[k]); for (int i = 0; i < 12; i++) printf("%02x", h[64 * k + i]); printf("\n"); }
    }
}
