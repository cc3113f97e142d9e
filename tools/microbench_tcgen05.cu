// Issue-rate microbenchmark of tcgen05.mma on B200 for the operand kinds an Ozaki-split fp64 GEMM could use
// (kind::i8 with exact s32 accumulation, kind::f8f6f4 and kind::f16 for comparison). Not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/microbench_tcgen05.cu -o tools/_build/mb_tcgen05
// One CTA per SM; one thread issues NMMA back-to-back 128 x 256 x (32 bytes of K) MMAs from shared-memory descriptors
// (operand contents are irrelevant for the rate; they rotate over four tiles so that the reads are real) into one TMEM
// accumulator, commits them to an mbarrier and waits. Reports ops per clock per SM and the chip-wide rate.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// UMMA shared-memory descriptor (cute/arch/mma_sm100_desc.hpp:98-113): no swizzle, K-major, 8-row x 16-byte core matrices
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);             // start address [0,14)
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;    // leading-dimension byte offset [16,30)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;    // stride-dimension byte offset [32,46)
  d |= (uint64_t)1 << 46;                               // descriptor version 1 (sm_100)
  return d;                                             // base offset 0, layout type 0 = SWIZZLE_NONE
}

// instruction descriptor (mma_sm100_desc.hpp:412-434)
__host__ __device__ constexpr uint32_t make_idesc(int c_fmt, int a_fmt, int b_fmt, int M, int N) {
  return (uint32_t)(c_fmt << 4) | (uint32_t)(a_fmt << 7) | (uint32_t)(b_fmt << 10) | (uint32_t)((N >> 3) << 17) |
         (uint32_t)((M >> 4) << 24);
}

template <int KIND, int N>   // 0 = i8 (s8 x s8 -> s32), 1 = f8f6f4 (e4m3 -> f32), 2 = f16 (bf16 -> f32); N = MMA width
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(long long* cycles, int nmma) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) uint64_t bar;
  constexpr int M = 128;
  constexpr int A_BYTES = M * 32, B_BYTES = N * 32;     // one MMA consumes 32 bytes of K per row
  for (int i = threadIdx.x; i < 4 * (A_BYTES + B_BYTES) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");      // generic-proxy fill of the operand tiles -> async proxy (MMA reads)
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t taddr = tmem_base;
  constexpr uint32_t idesc = KIND == 0 ? make_idesc(2, 1, 1, M, N) : KIND == 1 ? make_idesc(1, 0, 0, M, N) : make_idesc(1, 1, 1, M, N);
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int i = 0; i < nmma; i++) {
      const uint32_t a = smem_u32(smem) + (i & 3) * (A_BYTES + B_BYTES);
      const uint64_t da = make_desc(a, 128, 256), db = make_desc(a + A_BYTES, 128, 256);
      const uint32_t acc = i > 0;
      if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(taddr), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
      else if (KIND == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(taddr), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
      else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(taddr), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    bool done = false;
    for (long spin = 0; spin < (1L << 26) && !done; spin++) {   // bounded: a bad descriptor must not hang the GPU
      uint32_t ok;
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
      done = ok != 0;
    }
    if (!done) { cycles[blockIdx.x] = -1; }
    else
    cycles[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(taddr));
}

template <int KIND, int N = 256>
static void run(const char* name, int k_elems, int nsm, double ghz) {
  const int smem = 4 * (128 * 32 + N * 32);
  cudaFuncSetAttribute(mma_rate_kernel<KIND, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* d; cudaMalloc(&d, nsm * sizeof(long long));
  const int nmma = 4096;
  for (int rep = 0; rep < 3; rep++) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    mma_rate_kernel<KIND, N><<<nsm, 128, smem>>>(d, nmma);
    cudaEventRecord(e1);
    cudaError_t err = cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    long long h[1024]; cudaMemcpy(h, d, nsm * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < nsm; i++) mx = h[i] > mx ? h[i] : mx;
    const double ops = 2.0 * 128 * N * k_elems * (double)nmma;
    printf("%-18s rep %d: %s  %.1f clk per 128x%dx%d MMA, %.0f ops/clk/SM, chip %.1f Tops/s (kernel %.3f ms -> %.1f Tops/s)\n", name, rep,
           cudaGetErrorString(err), (double)mx / nmma, N, k_elems, ops / mx, ops / mx * nsm * ghz * 1e-3, ms, ops * nsm / (ms * 1e-3) * 1e-12);
  }
  cudaFree(d);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz * 1e-6;
  printf("%s, %d SMs, %.3f GHz\n", p.name, p.multiProcessorCount, ghz);
  run<0>("kind::i8 s8xs8->s32", 32, p.multiProcessorCount, ghz);
  run<1>("kind::f8f6f4 e4m3", 32, p.multiProcessorCount, ghz);
  run<2>("kind::f16 bf16", 16, p.multiProcessorCount, ghz);
  // narrower MMAs: is the rate still 8192 MAC/clk/SM when the A tile (4 KB) is re-read for every N = 128 / 64 / 32 columns?
  run<0, 128>("kind::i8 N=128", 32, p.multiProcessorCount, ghz);
  run<0, 64>("kind::i8 N=64", 32, p.multiProcessorCount, ghz);
  run<0, 32>("kind::i8 N=32", 32, p.multiProcessorCount, ghz);
  return 0;
}
