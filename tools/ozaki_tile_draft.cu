// Next-round building block, correctness first (NOT part of the library, NOT tuned): one 128 x 64 tile of C = A B^T in fp64
// accuracy computed on the INT8 tensor path (tcgen05.mma kind::i8, exact s32 accumulation in TMEM) by the digit split that
// tools/proto_ozaki.py validates numerically:
//   rows of A and B scaled to (-1, 1) by a power of two, S signed 7-bit digits per element (int8 planes in shared memory in
//   the no-swizzle K-major core-matrix layout pinned by tools/tcgen05_i8_check.cu), the slice pairs with s + t <= S + 1
//   accumulated per group g = s + t in its own 64 TMEM columns (8 groups x 64 = 512 columns), groups converted and summed in
//   fp64 smallest first, rows/columns rescaled.
// No pipelining, no TMA: one CTA, 128 threads, K processed in chunks of 32. Compared with a long-double reference on the host.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/ozaki_tile_draft.cu -o tools/_build/ozaki_tile_draft
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int TM = 128, TN = 64, KC = 32, S = 8;
constexpr int A_PLANE = TM * KC, B_PLANE = TN * KC;     // bytes per digit plane and chunk

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {   // LBO = 128 B (16-byte K chunks), SBO = 256 B (8-row groups)
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {   // s32 accumulate, signed 8-bit A and B, both K-major
  return (uint32_t)(2 << 4) | (uint32_t)(1 << 7) | (uint32_t)(1 << 10) | (uint32_t)((N >> 3) << 17) | (uint32_t)((M >> 4) << 24);
}
__device__ __forceinline__ int core_off(int row, int kk) { return (row / 8) * 256 + (kk / 16) * 128 + (row % 8) * 16 + kk % 16; }

// A: [TM][K] row-major fp64, B: [TN][K] row-major fp64, C: [TM][TN] row-major fp64
__global__ void __launch_bounds__(128, 1) ozaki_tile_kernel(const double* A, const double* B, double* C, int K) {
  extern __shared__ __align__(1024) unsigned char smem[];   // [S][A_PLANE] then [S][B_PLANE]
  unsigned char* sA = smem;
  unsigned char* sB = smem + S * A_PLANE;
  __shared__ int eA[TM], eB[TN];
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // ---- row exponents: |x| < 2^e for every element of the row -------------------------------------------------------
  {
    double amax = 0.0;
    for (int k = 0; k < K; k++) amax = fmax(amax, fabs(A[(long)tid * K + k]));
    int e = 0;
    if (amax > 0.0) frexp(amax, &e);
    eA[tid] = e;
    if (tid < TN) {
      double bmax = 0.0;
      for (int k = 0; k < K; k++) bmax = fmax(bmax, fabs(B[(long)tid * K + k]));
      int f = 0;
      if (bmax > 0.0) frexp(bmax, &f);
      eB[tid] = f;
    }
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t taddr = tmem_base;
  constexpr uint32_t idesc = make_idesc(TM, TN);
  const int nchunk = K / KC;
  bool ok_all = true;
  for (int ch = 0; ch < nchunk; ch++) {
    // ---- digit planes of this K chunk ----------------------------------------------------------------------------
    for (int idx = tid; idx < TM * KC; idx += 128) {
      const int row = idx / KC, kk = idx % KC;
      double r = ldexp(A[(long)row * K + ch * KC + kk], -eA[row]);
      const int off = core_off(row, kk);
#pragma unroll
      for (int s = 0; s < S; s++) {
        r *= 128.0;
        const double d = trunc(r);
        r -= d;
        sA[s * A_PLANE + off] = (unsigned char)(signed char)(int)d;
      }
    }
    for (int idx = tid; idx < TN * KC; idx += 128) {
      const int row = idx / KC, kk = idx % KC;
      double r = ldexp(B[(long)row * K + ch * KC + kk], -eB[row]);
      const int off = core_off(row, kk);
#pragma unroll
      for (int s = 0; s < S; s++) {
        r *= 128.0;
        const double d = trunc(r);
        r -= d;
        sB[s * B_PLANE + off] = (unsigned char)(signed char)(int)d;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (tid == 0) {
      // group g = s + t (0-based digits: g = 0 .. S-1) accumulates in TMEM columns [g*TN, (g+1)*TN)
      for (int g = 0; g < S; g++)
        for (int s = 0; s <= g; s++) {
          const int t = g - s;
          const uint64_t da = make_desc(smem_u32(sA + s * A_PLANE)), db = make_desc(smem_u32(sB + t * B_PLANE));
          const uint32_t acc = (ch > 0 || s > 0) ? 1u : 0u;
          const uint32_t dcol = taddr + (uint32_t)(g * TN);
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(dcol), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // all threads: wait until this chunk's MMAs have consumed the planes (bounded spin)
    bool done = false;
    for (long spin = 0; spin < (1L << 24) && !done; spin++) {
      uint32_t ok;
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"((uint32_t)(ch & 1)) : "memory");
      done = ok != 0;
    }
    ok_all = ok_all && done;
    asm volatile("tcgen05.fence::after_thread_sync;");
  }
  // ---- epilogue: groups from TMEM, smallest magnitude first, in fp64 ---------------------------------------------------
  const int row = warp * 32 + lane;
  double acc[TN];
#pragma unroll
  for (int j = 0; j < TN; j++) acc[j] = 0.0;
  for (int g = S - 1; g >= 0; g--) {
    const double scale = ldexp(1.0, -7 * (g + 2));       // digits are 1-based in the expansion: 2^(-7 (s+1)) 2^(-7 (t+1))
#pragma unroll
    for (int c0 = 0; c0 < TN; c0 += 32) {
      uint32_t v[32];
      const uint32_t addr = taddr + ((uint32_t)(warp * 32) << 16) + (uint32_t)(g * TN + c0);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
          "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
            "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
            "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
            "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(addr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; j++) acc[c0 + j] = fma((double)(int32_t)v[j], scale, acc[c0 + j]);
    }
  }
  for (int j = 0; j < TN; j++) C[(long)row * TN + j] = ok_all ? ldexp(acc[j], eA[row] + eB[j]) : nan("");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(taddr));
}

int main(int argc, char** argv) {
  const int K = argc > 1 ? atoi(argv[1]) : 1024;
  std::vector<double> A((size_t)TM * K), B((size_t)TN * K), C((size_t)TM * TN);
  srand(3);
  auto rnd = []() { return (rand() / (double)RAND_MAX) * 2.0 - 1.0; };
  for (int i = 0; i < TM; i++) { const double sc = exp(8.0 * rnd()); for (int k = 0; k < K; k++) A[(size_t)i * K + k] = sc * rnd() * exp(2.0 * rnd()); }
  for (int j = 0; j < TN; j++) { const double sc = exp(8.0 * rnd()); for (int k = 0; k < K; k++) B[(size_t)j * K + k] = sc * rnd() * exp(2.0 * rnd()); }
  double *dA, *dB, *dC;
  cudaMalloc(&dA, A.size() * 8); cudaMalloc(&dB, B.size() * 8); cudaMalloc(&dC, C.size() * 8);
  cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice);
  const int smem = S * (A_PLANE + B_PLANE);
  cudaFuncSetAttribute(ozaki_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  ozaki_tile_kernel<<<1, 128, smem>>>(dA, dB, dC, K);
  cudaError_t err = cudaDeviceSynchronize();
  cudaMemcpy(C.data(), dC, C.size() * 8, cudaMemcpyDeviceToHost);
  double worst_oz = 0, worst_f64 = 0; long nans = 0;
  for (int i = 0; i < TM; i++)
    for (int j = 0; j < TN; j++) {
      long double ref = 0, den = 0; double f64 = 0;
      for (int k = 0; k < K; k++) {
        ref += (long double)A[(size_t)i * K + k] * (long double)B[(size_t)j * K + k];
        den += fabsl((long double)A[(size_t)i * K + k] * (long double)B[(size_t)j * K + k]);
        f64 = fma(A[(size_t)i * K + k], B[(size_t)j * K + k], f64);
      }
      const double c = C[(size_t)i * TN + j];
      if (c != c) { nans++; continue; }
      worst_oz = fmax(worst_oz, (double)(fabsl((long double)c - ref) / den));
      worst_f64 = fmax(worst_f64, (double)(fabsl((long double)f64 - ref) / den));
    }
  printf("K=%d S=%d digits (%d slice-pair MMAs per 32-deep chunk): %s, NaN entries %ld, max |err| / (|A||B|^T): int8 split %.3e, plain fp64 fma loop %.3e\n",
         K, S, S * (S + 1) / 2, cudaGetErrorString(err), nans, worst_oz, worst_f64);
  return 0;
}
