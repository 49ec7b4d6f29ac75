#include "../phase2_bn254_b200/csrc/msm_impl.cuh"
using namespace p2b;
__global__ void k_fill(uint32_t* s1, uint32_t* s2, int n) {
    Xyzz<Fq> G; G.x = fp_one<FqP>(); G.y = dbl(fp_one<FqP>()); G.zz = fp_one<FqP>(); G.zzz = fp_one<FqP>();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        store_xyzz<Fq>(s1, i, i == 0 ? G : xyzz_infinity<Fq>());
        store_xyzz<Fq>(s2, i, i == 0 ? G : xyzz_infinity<Fq>());
    }
}
__global__ void k_wire(const uint32_t* ws, int n, uint32_t* out) {
    if (threadIdx.x) return;
    for (int i = 0; i < n; i++) {
        Xyzz<Fq> acc = load_xyzz<Fq>(ws, i);
        bool inf = is_zero(acc.zz);
        Fq t = inv(mul(acc.zz, acc.zzz));
        Aff<Fq> a; a.x = mul(acc.x, mul(t, acc.zzz)); a.y = mul(acc.y, mul(t, acc.zz));
        uint32_t o[16]; point_encode<Fq>(o, a, inf, ENC_UNCOMPRESSED);
        for (int j = 0; j < 16; j++) out[16 * i + j] = o[j];
    }
}
int main() {
    MsmGeom g = msm_geometry(1);
    uint32_t *s1, *s2, *ws, *out;
    size_t nred = (size_t)g.nwin * g.tpw;
    cudaMalloc(&s1, nred * 128); cudaMalloc(&s2, nred * 128); cudaMalloc(&ws, (g.nwin + 16) * 128); cudaMalloc(&out, 64 * 64);
    cudaMemset(ws, 0, (g.nwin + 16) * 128);
    k_fill<<<1, 256>>>(s1, s2, (int)nred);
    k_msm_reduce2<Fq><<<g.nwin, 256>>>(s1, s2, g, ws);
    k_wire<<<1, 32>>>(ws + (size_t)g.nwin * 32, 7, out + 16);
    k_wire<<<1, 32>>>(ws, 1, out);
    uint8_t h[8 * 64]; cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    const char* names[8] = {"W0", "wp(x32)", "qs", "as", "b", "P0", "Q0", "A0"};
    for (int k = 0; k < 8; k++) { printf(" %-8s ", names[k]); for (int i = 0; i < 12; i++) printf("%02x", h[64 * k + i]); printf("\n"); }
}
