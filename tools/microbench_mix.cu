// Are DMMA.8x8x4 (fp64 "tensor") and DFMA (fp64 CUDA-core) separate pipes on B200? Run them in the same kernel
// (even warps DMMA, odd warps DFMA) and compare with each alone.
#include <cuda_runtime.h>
#include <cstdio>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1;} }while(0)

__global__ void mix(double* out, int iters, int mode) {  // mode 0: all DMMA, 1: all DFMA, 2: even warps DMMA / odd DFMA
  const int warp = threadIdx.x >> 5;
  const bool do_mma = mode == 0 || (mode == 2 && (warp & 1) == 0);
  double c[8][2];
  for (int i = 0; i < 8; i++) { c[i][0] = i; c[i][1] = -i; }
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  if (do_mma) {
    for (int it = 0; it < iters; it++)
#pragma unroll
      for (int i = 0; i < 8; i++)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  } else {
    for (int it = 0; it < 4 * iters; it++)   // 4x the iterations: same flops per warp as the DMMA warps
#pragma unroll
      for (int i = 0; i < 8; i++) { c[i][0] = fma(c[i][0], a, b); c[i][1] = fma(c[i][1], b, a); }
  }
  double s = 0;
  for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  double* out; CK(cudaMalloc(&out, (size_t)p.multiProcessorCount * 2 * 1024 * 8));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int iters = 20000, grid = p.multiProcessorCount * 2, threads = 512;
  for (int mode = 0; mode < 3; mode++) {
    mix<<<grid, threads>>>(out, iters, mode); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); mix<<<grid, threads>>>(out, iters, mode); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double warps = (double)grid * threads / 32;
    double fl_mma = 0, fl_fma = 0;
    if (mode == 0) fl_mma = warps * iters * 8 * 512.0;
    if (mode == 1) fl_fma = warps * 4.0 * iters * 16 * 64.0;
    if (mode == 2) { fl_mma = warps / 2 * iters * 8 * 512.0; fl_fma = warps / 2 * 4.0 * iters * 16 * 64.0; }
    printf("mode %d: %.3f ms  DMMA %.2f TF/s + DFMA %.2f TF/s = %.2f TF/s\n", mode, ms, fl_mma / ms * 1e-9, fl_fma / ms * 1e-9,
           (fl_mma + fl_fma) / ms * 1e-9);
  }
  return 0;
}
