// Raw pipe rates on sm_100a: DFMA, DADD, IMAD.WIDE, 64-bit integer adds, alone and mixed in one instruction stream.
// Purpose: is the FP64 pipe a usable second multiplier for 256-bit Montgomery arithmetic?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int NF, int NI, int NA> __global__ void k_mix(double *out, double b, double c, uint32_t ib, int iters) {
    double f[NF > 0 ? NF : 1];
    uint64_t w[NI > 0 ? NI : 1];
    uint32_t al[NA > 0 ? NA : 1], ah[NA > 0 ? NA : 1];
    for (int j = 0; j < NF; j++) f[j] = threadIdx.x + j;
    for (int j = 0; j < NI; j++) w[j] = threadIdx.x + j;
    for (int j = 0; j < NA; j++) { al[j] = threadIdx.x + j; ah[j] = j; }
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int j = 0; j < NF; j++) f[j] = __fma_rz(f[j], b, c);
#pragma unroll
            for (int j = 0; j < NI; j++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[j]) : "r"((uint32_t)w[j]), "r"(ib));
#pragma unroll
            for (int j = 0; j < NA; j++)
                asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(al[j]), "+r"(ah[j]) : "r"(ib), "r"(al[j]));
        }
    }
    double s = 0;
    for (int j = 0; j < NF; j++) s += f[j];
    for (int j = 0; j < NI; j++) s += (double)w[j];
    for (int j = 0; j < NA; j++) s += (double)al[j] + ah[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dadd(double *out, double b, int iters) {
    double f[8];
    for (int j = 0; j < 8; j++) f[j] = threadIdx.x + j;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int j = 0; j < 8; j++) f[j] = __dadd_rz(f[j], b);
    }
    double s = 0;
    for (int j = 0; j < 8; j++) s += f[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
static int sms;
static cudaEvent_t e0, e1;
template <int NF, int NI, int NA> void run(const char *name, double *out, int warps) {
    int iters = 2048, threads = warps * 32;
    k_mix<NF, NI, NA><<<sms, threads>>>(out, 1.0000001, 3.0, 5, 8);
    cudaEventRecord(e0);
    k_mix<NF, NI, NA><<<sms, threads>>>(out, 1.0000001, 3.0, 5, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double per = (double)sms * threads * iters * 4;
    double clk = 1.965e9 * ms * 1e-3 * sms;  // SM-clocks
    printf("%-28s warps/SM=%2d  %.3f ms  per clk per SM: DFMA %.1f  IMAD.WIDE %.1f  ADD64 %.1f\n", name, warps, ms, per * NF / clk,
           per * NI / clk, per * NA / clk);
}
int main() {
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *out;
    cudaMalloc(&out, 64 << 20);
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int warps : {8, 16, 32}) {
        run<8, 0, 0>("DFMA only", out, warps);
        run<0, 8, 0>("IMAD.WIDE only", out, warps);
        run<0, 0, 8>("ADD64 (2 instr) only", out, warps);
        run<4, 8, 0>("DFMA 4 : WIDE 8", out, warps);
        run<8, 8, 0>("DFMA 8 : WIDE 8", out, warps);
        run<8, 4, 0>("DFMA 8 : WIDE 4", out, warps);
        run<8, 0, 8>("DFMA 8 : ADD64 8", out, warps);
        run<0, 8, 8>("WIDE 8 : ADD64 8", out, warps);
        run<8, 8, 8>("DFMA 8 : WIDE 8 : ADD64 8", out, warps);
        run<8, 4, 8>("DFMA 8 : WIDE 4 : ADD64 8", out, warps);
        int iters = 2048, threads = warps * 32;
        cudaEventRecord(e0);
        k_dadd<<<sms, threads>>>(out, 1.5, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("%-28s warps/SM=%2d  %.3f ms  per clk per SM: DADD %.1f\n", "DADD only", warps, ms,
               (double)sms * threads * iters * 32 / (1.965e9 * ms * 1e-3 * sms));
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
}
