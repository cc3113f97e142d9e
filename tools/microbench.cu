// Microbenchmarks that size the fp64 roofline on B200 (sm_100a): DMMA.8x8x4 issue rate, DFMA rate.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)

template<int NACC>
__global__ void dmma_rate(double* out, int iters) {
  double c[NACC][2];
  #pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = 0.0; c[i][1] = 0.0; }
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; it++) {
    #pragma unroll
    for (int i = 0; i < NACC; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  #pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template<int NACC>
__global__ void dfma_rate(double* out, int iters) {
  double c[NACC];
  #pragma unroll
  for (int i = 0; i < NACC; i++) c[i] = i;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; it++) {
    #pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0;
  #pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template<typename F> float time_ms(F f) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); return ms;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s SMs %d clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  double* out; CK(cudaMalloc(&out, 148 * 8 * 1024 * sizeof(double)));
  int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    for (int bps : {1, 2}) {
      int grid = p.multiProcessorCount * bps;
      float ms = time_ms([&]{ dmma_rate<8><<<grid, warps * 32>>>(out, iters); });
      double flops = (double)grid * warps * iters * 8 * 512.0;
      printf("DMMA.8x8x4 nacc=8 warps/cta=%2d cta/sm=%d : %.2f TFLOP/s  (%.1f flop/clk/SM @1.965GHz)\n", warps, bps,
             flops / ms * 1e-9, flops / ms * 1e-9 * 1e12 / p.multiProcessorCount / 1.965e9);
    }
  }
  {
    int grid = p.multiProcessorCount; int warps = 8;
    float ms = time_ms([&]{ dmma_rate<2><<<grid, warps * 32>>>(out, iters); });
    double flops = (double)grid * warps * iters * 2 * 512.0;
    printf("DMMA.8x8x4 nacc=2 warps/cta=8 : %.2f TFLOP/s\n", flops / ms * 1e-9);
    ms = time_ms([&]{ dmma_rate<16><<<grid, warps * 32>>>(out, iters); });
    flops = (double)grid * warps * iters * 16 * 512.0;
    printf("DMMA.8x8x4 nacc=16 warps/cta=8 : %.2f TFLOP/s\n", flops / ms * 1e-9);
    // latency: 1 warp, 1 accumulator chain
    ms = time_ms([&]{ dmma_rate<1><<<1, 32>>>(out, iters); });
    printf("DMMA.8x8x4 dependent-chain latency: %.1f ns/op (%.1f clk @1.965GHz)\n", ms * 1e6 / iters, ms * 1e6 / iters * 1.965);
  }
  for (int warps : {8, 16, 32}) {
    int grid = p.multiProcessorCount * 2;
    float ms = time_ms([&]{ dfma_rate<8><<<grid, warps * 32>>>(out, iters); });
    double flops = (double)grid * warps * 32 * iters * 8 * 2.0;
    printf("DFMA nacc=8 warps/cta=%2d cta/sm=2 : %.2f TFLOP/s\n", warps, flops / ms * 1e-9);
  }
  return 0;
}
