// IMAD / IMAD.WIDE throughput microbenchmark: the compute ceiling for 8x32-bit Montgomery arithmetic.
#include <cstdio>
#include <cuda_runtime.h>
#include "../phase2_bn254_b200/csrc/fp.cuh"
using namespace p2b;
template<int ILP> __global__ void k_wide(uint64_t* out, uint32_t a, uint32_t b, int iters){
  uint64_t acc[ILP];
  for(int j=0;j<ILP;j++) acc[j]=threadIdx.x+j;
  for(int i=0;i<iters;i++){
#pragma unroll
    for(int j=0;j<ILP;j++) acc[j] = (uint64_t)(uint32_t)(acc[j]) * b + acc[j] + a;
  }
  uint64_t s=0; for(int j=0;j<ILP;j++) s+=acc[j];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int ILP> __global__ void k_imad(uint32_t* out, uint32_t a, uint32_t b, int iters){
  uint32_t acc[ILP];
  for(int j=0;j<ILP;j++) acc[j]=threadIdx.x+j;
  for(int i=0;i<iters;i++){
#pragma unroll
    for(int j=0;j<ILP;j++) acc[j] = acc[j]*b + a;
  }
  uint32_t s=0; for(int j=0;j<ILP;j++) s+=acc[j];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void k_mul(Fq* a, const Fq* b, int iters){
  size_t t = blockIdx.x*blockDim.x+threadIdx.x;
  Fq x = a[t], y = b[t];
#pragma unroll 1
  for(int i=0;i<iters;i++){ x = mul(x,y); y = mul(y,x); }
  a[t]=x;
}
__global__ void k_mul4(Fq* a, const Fq* b, int iters){
  size_t t = blockIdx.x*blockDim.x+threadIdx.x;
  Fq x = a[t], y = b[t], z = a[t+1], w = b[t+1];
#pragma unroll 1
  for(int i=0;i<iters;i++){ x = mul(x,y); z = mul(z,w); y = mul(y,x); w = mul(w,z); }
  a[t]=add(x,z);
}
int main(){
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  void* out; cudaMalloc(&out, 64<<20); cudaMemset(out, 1, 64<<20);
  void* b; cudaMalloc(&b, 64<<20); cudaMemset(b, 3, 64<<20);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  for (int warps : {4, 8, 16, 32}) {
    int threads = warps*32; int iters = 4096;
    k_wide<8><<<sms, threads>>>((uint64_t*)out, 3, 5, 16);
    cudaEventRecord(e0); k_wide<8><<<sms, threads>>>((uint64_t*)out, 3, 5, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms,e0,e1);
    double ops = (double)sms*threads*iters*8;
    printf("IMAD.WIDE warps/SM=%2d: %.2f Tops/s  (%.1f per clk per SM @1.9GHz)\n", warps, ops/ms/1e9, ops/ms/1e6/sms/1.9e3);
    cudaEventRecord(e0); k_imad<8><<<sms, threads>>>((uint32_t*)out, 3, 5, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms,e0,e1);
    printf("IMAD      warps/SM=%2d: %.2f Tops/s  (%.1f per clk per SM @1.9GHz)\n", warps, ops/ms/1e9, ops/ms/1e6/sms/1.9e3);
    int mi = 2000;
    cudaEventRecord(e0); k_mul<<<sms, threads>>>((Fq*)out, (Fq*)b, mi); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms,e0,e1);
    double muls = (double)sms*threads*mi*2;
    printf("mont_mul dep-chain warps/SM=%2d: %.2f Gmul/s\n", warps, muls/ms/1e6);
    cudaEventRecord(e0); k_mul4<<<sms, threads>>>((Fq*)out, (Fq*)b, mi); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms,e0,e1);
    printf("mont_mul 2-chains  warps/SM=%2d: %.2f Gmul/s\n", warps, muls*2/ms/1e6);
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
}
