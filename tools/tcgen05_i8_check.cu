// Correctness probe for the next-round kernel: ONE 128 x 256 x 64 int8 product on tcgen05 (kind::i8, s32 accumulation in
// TMEM) from no-swizzle K-major shared-memory descriptors, read back with tcgen05.ld and compared with the CPU. Pins the
// operand layout (core matrices of 8 rows x 16 bytes; which descriptor field strides K and which strides M/N) and the
// TMEM addressing (lane = row, column = n) before any real kernel is built on it. Not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/tcgen05_i8_check.cu -o tools/_build/tcgen05_i8_check
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int c_fmt, int a_fmt, int b_fmt, int M, int N) {
  return (uint32_t)(c_fmt << 4) | (uint32_t)(a_fmt << 7) | (uint32_t)(b_fmt << 10) | (uint32_t)((N >> 3) << 17) |
         (uint32_t)((M >> 4) << 24);
}
constexpr int M = 128, N = 256, KSTEPS = 2, KB = 32;   // K = 64
constexpr int A_BYTES = M * KB, B_BYTES = N * KB;

// A: [M][K] row-major int8 in global, B: [N][K]; D: [M][N] int32.  fill_k_stride / fill_mn_stride: where the kernel
// PLACES core matrices; desc_lbo / desc_sbo: what it TELLS the hardware.
__global__ void __launch_bounds__(128, 1) check_kernel(const int8_t* A, const int8_t* B, int32_t* D, int fill_k_stride,
                                                       int fill_mn_stride, int desc_lbo, int desc_sbo) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) uint64_t bar;
  const int K = KSTEPS * KB;
  for (int idx = threadIdx.x; idx < M * K; idx += blockDim.x) {
    const int r = idx / K, k = idx % K, step = k / KB, kk = k % KB;
    smem[step * (A_BYTES + B_BYTES) + (r / 8) * fill_mn_stride + (kk / 16) * fill_k_stride + (r % 8) * 16 + kk % 16] = (unsigned char)A[idx];
  }
  for (int idx = threadIdx.x; idx < N * K; idx += blockDim.x) {
    const int r = idx / K, k = idx % K, step = k / KB, kk = k % KB;
    smem[step * (A_BYTES + B_BYTES) + A_BYTES + (r / 8) * fill_mn_stride + (kk / 16) * fill_k_stride + (r % 8) * 16 + kk % 16] = (unsigned char)B[idx];
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t taddr = tmem_base;
  constexpr uint32_t idesc = make_idesc(2, 1, 1, M, N);   // s32 accumulate, signed 8-bit A and B, K-major both
  if (threadIdx.x == 0) {
    for (int step = 0; step < KSTEPS; step++) {
      const uint32_t a = smem_u32(smem) + step * (A_BYTES + B_BYTES);
      const uint64_t da = make_desc(a, desc_lbo, desc_sbo), db = make_desc(a + A_BYTES, desc_lbo, desc_sbo);
      const uint32_t acc = step > 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(taddr), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // everybody waits for the MMAs (bounded spin)
  bool done = false;
  for (long spin = 0; spin < (1L << 24) && !done; spin++) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    done = ok != 0;
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = warp * 32 + lane;                       // TMEM lane = accumulator row; a warp owns its 32-lane quarter
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    const uint32_t addr = taddr + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; j++) D[(long)row * N + c0 + j] = done ? (int32_t)v[j] : -123456789;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(taddr));
}

int main() {
  const int K = KSTEPS * KB;
  std::vector<int8_t> A((size_t)M * K), B((size_t)N * K);
  srand(1);
  for (auto& x : A) x = (int8_t)(rand() % 255 - 127);
  for (auto& x : B) x = (int8_t)(rand() % 255 - 127);
  std::vector<int32_t> ref((size_t)M * N), out((size_t)M * N);
  for (int i = 0; i < M; i++)
    for (int j = 0; j < N; j++) {
      int32_t s = 0;
      for (int k = 0; k < K; k++) s += (int32_t)A[(size_t)i * K + k] * (int32_t)B[(size_t)j * K + k];
      ref[(size_t)i * N + j] = s;
    }
  int8_t *dA, *dB; int32_t* dD;
  cudaMalloc(&dA, A.size()); cudaMalloc(&dB, B.size()); cudaMalloc(&dD, out.size() * 4);
  cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice);
  const int smem = KSTEPS * (A_BYTES + B_BYTES);
  cudaFuncSetAttribute(check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  // the data is always PLACED with K-chunks 128 B apart and 8-row groups 256 B apart; the two runs tell the hardware
  // (LBO, SBO) = (128, 256) and (256, 128)
  const int cfg[2][2] = {{128, 256}, {256, 128}};
  for (int v = 0; v < 2; v++) {
    cudaMemset(dD, 0xff, out.size() * 4);
    check_kernel<<<1, 128, smem>>>(dA, dB, dD, 128, 256, cfg[v][0], cfg[v][1]);
    cudaError_t err = cudaDeviceSynchronize();
    cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
    long bad = 0;
    for (size_t i = 0; i < out.size(); i++) bad += out[i] != ref[i];
    printf("descriptor LBO=%d SBO=%d (data placed: K-chunk stride 128 B, 8-row-group stride 256 B): %s, mismatches %ld / %zu, D[0][0..3] = %d %d %d %d (ref %d %d %d %d)\n",
           cfg[v][0], cfg[v][1], cudaGetErrorString(err), bad, out.size(), out[0], out[1], out[2], out[3], ref[0], ref[1], ref[2], ref[3]);
  }
  return 0;
}
