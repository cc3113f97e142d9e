// Issue-pattern microbenchmark for the digit-split GEMM (not part of the library): one CTA per SM, 8 A planes (128 x 32 B)
// and 8 B planes (N x 32 B) resident in shared memory, NO TMA, no epilogue; one thread issues the digit-pair MMAs of a
// k-chunk over and over in different orders. Tells apart "issue loop" / "accumulator switching" / "shared-memory read"
// limits of tcgen05.mma kind::i8 at N = 64 and N = 128.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/microbench_ozaki_pattern.cu -o tools/_build/mb_ozaki_pattern
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (uint32_t)(2 << 4) | (uint32_t)(1 << 7) | (uint32_t)(1 << 10) | (uint32_t)((N >> 3) << 17) | (uint32_t)((M >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da),
               "l"(db), "r"(idesc), "r"(acc) : "memory");
}

// ORDER 0: by exponent group (g outer, s inner; consecutive MMAs share the accumulator)      [what the library does]
// ORDER 1: by A plane (s outer, t inner; consecutive MMAs share the A operand, accumulator changes every time)
// ORDER 2: all pairs into ONE accumulator (reference: pure shared-memory / issue rate)
// G0..G1: range of groups issued (N = 128 can only hold 4 groups in TMEM: 4..7 = 26 pairs, 0..3 = 10 pairs)
template <int N, int ORDER, int G0, int G1, int SYNC = 0>
__global__ void __launch_bounds__(128, 1) pattern_kernel(long long* cycles, int nchunks, int* npairs) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(8) uint64_t bar2[4], bar3;
  constexpr int A_BYTES = 128 * 32, B_BYTES = N * 32;
  for (int i = threadIdx.x; i < 8 * (A_BYTES + B_BYTES) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    for (int i = 0; i < 4; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2[i])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar3)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t taddr = tmem_base;
  constexpr uint32_t idesc = make_idesc(128, N);
  constexpr uint64_t hi = ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46);
  if (threadIdx.x == 0) {
    const uint32_t a_lo = (smem_u32(smem) & 0x3FFFF) >> 4, b_lo = a_lo + ((8 * A_BYTES) >> 4);
    int pairs = 0;
    const long long t0 = clock64();
    for (int ch = 0; ch < nchunks; ch++) {
      const uint32_t acc0 = ch > 0;
      if (ORDER == 1) {
#pragma unroll
        for (int s = 0; s < 8; s++)
#pragma unroll
          for (int t = 0; t < 8; t++) {
            const int g = s + t;
            if (g >= G0 && g <= G1) {
              umma(taddr + (uint32_t)((g - G0) * N), hi | (uint64_t)(a_lo + s * (A_BYTES >> 4)), hi | (uint64_t)(b_lo + t * (B_BYTES >> 4)), idesc,
                   (s > 0 && t < 7 && g - 1 >= 0 && s - 1 + t + 1 == g) ? 1u : acc0 | (s > 0 ? 1u : 0u));
              if (ch == 0) pairs++;
            }
          }
      } else {
#pragma unroll
        for (int g = G0; g <= G1; g++)
#pragma unroll
          for (int s = 0; s <= g; s++) {
            const uint32_t d = ORDER == 2 ? taddr : taddr + (uint32_t)((g - G0) * N);
            umma(d, hi | (uint64_t)(a_lo + s * (A_BYTES >> 4)), hi | (uint64_t)(b_lo + (g - s) * (B_BYTES >> 4)), idesc, s > 0 ? 1u : acc0);
            if (ch == 0) pairs++;
          }
      }
      if (SYNC >= 1)   // what the library does after every k-chunk: release the stage
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2[ch & 3])) : "memory");
      if (SYNC >= 3) {  // ... and wait for the next stage (here: a barrier whose phase -1 is complete: returns at once)
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 1;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar3)) : "memory");
      }
      if (SYNC >= 2) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    bool done = false;
    for (long spin = 0; spin < (1L << 26) && !done; spin++) {
      uint32_t ok;
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
      done = ok != 0;
    }
    cycles[blockIdx.x] = done ? clock64() - t0 : -1;
    if (blockIdx.x == 0) *npairs = pairs;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(taddr));
}

template <int N, int ORDER, int G0, int G1, int SYNC = 0>
static void run(const char* name, int nsm) {
  const int smem = 8 * (128 * 32 + N * 32);
  cudaFuncSetAttribute(pattern_kernel<N, ORDER, G0, G1, SYNC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* d; int* np;
  cudaMalloc(&d, nsm * sizeof(long long)); cudaMalloc(&np, 4);
  const int nchunks = 256;
  for (int rep = 0; rep < 2; rep++) {
    pattern_kernel<N, ORDER, G0, G1, SYNC><<<nsm, 128, smem>>>(d, nchunks, np);
    cudaError_t err = cudaDeviceSynchronize();
    long long h[1024]; int pairs = 0;
    cudaMemcpy(h, d, nsm * sizeof(long long), cudaMemcpyDeviceToHost); cudaMemcpy(&pairs, np, 4, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < nsm; i++) mx = h[i] > mx ? h[i] : mx;
    if (rep == 1)
      printf("%-64s %s  %2d pairs/chunk: %7.1f clk per chunk = %5.1f clk per 128x%dx32 MMA (%.0f%% of the 8192 MAC/clk rate)\n", name,
             cudaGetErrorString(err), pairs, (double)mx / nchunks, (double)mx / nchunks / pairs, N,
             100.0 * 128.0 * N * 32 / 8192.0 / ((double)mx / nchunks / pairs));
  }
  cudaFree(d); cudaFree(np);
}

// TMEM drain: 8 warps (lane quarter = warp % 4, column half = warp / 4) read 8 groups x 64 columns of s32 and fold them into
// 32 fp64 accumulators per thread, the way the epilogue of the library does. MODE 0: ld.x32 -> wait -> math per group;
// MODE 1: the next group's ld is issued before the math of the current one; MODE 2: loads only (no math).
template <int MODE>
__global__ void __launch_bounds__(256, 1) drain_kernel(long long* cycles, double* sink, int reps) {
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 32);
  double acc[32];
  for (int j = 0; j < 32; j++) acc[j] = 0.0;
  __syncthreads();
  const long long t0 = clock64();
#define LD32(V, ADDR) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" \
      : "=r"(V[0]), "=r"(V[1]), "=r"(V[2]), "=r"(V[3]), "=r"(V[4]), "=r"(V[5]), "=r"(V[6]), "=r"(V[7]), "=r"(V[8]), "=r"(V[9]), "=r"(V[10]), "=r"(V[11]), "=r"(V[12]), "=r"(V[13]), "=r"(V[14]), "=r"(V[15]), "=r"(V[16]), "=r"(V[17]), "=r"(V[18]), "=r"(V[19]), "=r"(V[20]), "=r"(V[21]), "=r"(V[22]), "=r"(V[23]), "=r"(V[24]), "=r"(V[25]), "=r"(V[26]), "=r"(V[27]), "=r"(V[28]), "=r"(V[29]), "=r"(V[30]), "=r"(V[31]) : "r"(ADDR))
#define WAITLD() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")
#define MATH(V, G) _Pragma("unroll") for (int j = 0; j < 32; j++) acc[j] = fma(__hiloint2double(0x43300000, (int)(V[j] ^ 0x80000000u)) - 4503601774854144.0, __longlong_as_double((long long)(1023 - 7 * (G)) << 52), acc[j])
  for (int rep = 0; rep < reps; rep++) {
    if (MODE == 1) {
      uint32_t va[32], vb[32];
      LD32(va, taddr + 7 * 64); WAITLD();
#pragma unroll
      for (int g = 7; g >= 0; g -= 2) {
        LD32(vb, taddr + (g - 1) * 64);
        MATH(va, g);
        WAITLD();
        if (g >= 2) LD32(va, taddr + (g - 2) * 64);
        MATH(vb, g - 1);
        WAITLD();
      }
    } else {
#pragma unroll
      for (int g = 7; g >= 0; g--) {
        uint32_t v[32];
        LD32(v, taddr + g * 64); WAITLD();
        if (MODE == 0) MATH(v, g);
        else acc[0] += (double)v[lane & 31];
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
  double s = 0; for (int j = 0; j < 32; j++) s += acc[j];
  if (s == 12345.678) sink[0] = s;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
}

template <int MODE>
static void run_drain(const char* name, int nsm) {
  long long* d; double* sink; cudaMalloc(&d, nsm * 8); cudaMalloc(&sink, 8);
  const int reps = 64;
  for (int r = 0; r < 2; r++) drain_kernel<MODE><<<nsm, 256>>>(d, sink, reps);
  cudaError_t err = cudaDeviceSynchronize();
  long long h[1024]; cudaMemcpy(h, d, nsm * 8, cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < nsm; i++) mx = h[i] > mx ? h[i] : mx;
  printf("%-64s %s  %7.1f clk per drain of 8 groups x 128 lanes x 64 columns (256 KB): %.1f B/clk\n", name, cudaGetErrorString(err),
         (double)mx / reps, 262144.0 / ((double)mx / reps));
  cudaFree(d); cudaFree(sink);
}
static void drain_bench(int nsm) {
  run_drain<0>("TMEM drain, 8 warps, ld.x32 -> wait -> math per group", nsm);
  run_drain<1>("TMEM drain, 8 warps, next ld issued before the math", nsm);
  run_drain<2>("TMEM drain, 8 warps, loads only", nsm);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
  run<64, 0, 0, 7>("N=64, 36 pairs, by group (library order)", p.multiProcessorCount);
  run<64, 1, 0, 7>("N=64, 36 pairs, by A plane (accumulator changes every MMA)", p.multiProcessorCount);
  run<64, 2, 0, 7>("N=64, 36 pairs, ONE accumulator", p.multiProcessorCount);
  run<128, 0, 4, 7>("N=128, groups 4..7 (26 pairs), by group", p.multiProcessorCount);
  run<128, 0, 0, 3>("N=128, groups 0..3 (10 pairs), by group", p.multiProcessorCount);
  run<128, 2, 4, 7>("N=128, 26 pairs, ONE accumulator", p.multiProcessorCount);
  run<256, 0, 6, 7>("N=256, groups 6..7 (15 pairs), by group", p.multiProcessorCount);
  run<64, 0, 0, 7, 1>("N=64, 36 pairs, tcgen05.commit after every chunk", p.multiProcessorCount);
  run<64, 0, 0, 7, 2>("N=64, 36 pairs, commit + fence::after_thread_sync", p.multiProcessorCount);
  run<64, 0, 0, 7, 3>("N=64, 36 pairs, commit + mbarrier wait + fence", p.multiProcessorCount);
  run<128, 0, 4, 7, 3>("N=128, 26 pairs, commit + mbarrier wait + fence", p.multiProcessorCount);
  drain_bench(p.multiProcessorCount);
  return 0;
}
