// UNTESTED DRAFT for the next round (written after round 1's GPU budget was spent; it compiles for sm_100a, it has never
// run). Pipelined version of tools/ozaki_tile_draft.cu, which IS validated on the hardware:
//
//   C (M x N, fp64, column-major)  =  A (M x K) * B (N x K)^T      operands column-major, m-contiguous like the engine's
//
//   1. split_planes_kernel  : per operand, row exponents and S = 8 signed 7-bit digit planes (int8), stored tile by tile in
//                             exactly the shared-memory image tcgen05 wants (no-swizzle K-major core matrices), so that one
//                             plane of one (row tile, 32-deep K chunk) is ONE contiguous block = one cp.async.bulk.
//   2. ozaki_gemm_kernel    : one CTA per 128 x 64 output tile. warp 0 = producer (16 bulk copies per K chunk into a 3-stage
//                             ring, mbarrier expect_tx), warp 1 = MMA issuer (36 tcgen05.mma kind::i8 per chunk into 8 exponent
//                             groups x 64 TMEM columns, tcgen05.commit frees the stage), warps 2..5 = epilogue (tcgen05.ld,
//                             s32 -> fp64, groups summed smallest first, 2^(e_m + f_n), store).
//   Sizing: with N = 64 per MMA the A plane (4 KB) is re-read per instruction and shared-memory bandwidth,
//   not the MMA rate, is the expected limit (~2.4x DMMA); the 2-CTA / 128-wide variants come after this one is correct.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/ozaki_gemm_draft.cu -o tools/_build/ozaki_gemm_draft
//   tools/_build/ozaki_gemm_draft [M N K]      (checks against a long-double reference for small sizes, times the kernel)
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int TM = 128, TN = 64, KC = 32, S = 8, STAGES = 3;
constexpr int A_PLANE = TM * KC, B_PLANE = TN * KC;                 // 4096, 2048 bytes
constexpr int STAGE_BYTES = S * (A_PLANE + B_PLANE);                // 49152
constexpr int NTHREADS = 192;                                       // 6 warps

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}"
               ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {   // LBO = 128 B (16-byte K chunks), SBO = 256 B (8-row groups)
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {   // s32 accumulate, signed 8-bit A and B, both K-major
  return (uint32_t)(2 << 4) | (uint32_t)(1 << 7) | (uint32_t)(1 << 10) | (uint32_t)((N >> 3) << 17) | (uint32_t)((M >> 4) << 24);
}
__host__ __device__ __forceinline__ int core_off(int row, int kk) { return (row / 8) * 256 + (kk / 16) * 128 + (row % 8) * 16 + kk % 16; }

// ---- 1. exponents + digit planes ------------------------------------------------------------------------------------
// X: rows x K column-major (element (r, k) at r + k*ld). expo[r]: |x(r, :)| < 2^expo. planes: for row tile rt (TR rows), chunk
// kc, digit s: TR*KC bytes at (((rt * nkc) + kc) * S + s) * TR*KC, inside it the core-matrix image.
template <int TR>
__global__ void row_exponent_kernel(const double* __restrict__ X, long ld, long rows, long K, int* __restrict__ expo) {
  const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  double amax = 0.0;
  for (long k = 0; k < K; k++) amax = fmax(amax, fabs(X[r + k * ld]));
  int e = 0;
  if (amax > 0.0) frexp(amax, &e);
  expo[r] = e;
}
template <int TR>
__global__ void split_planes_kernel(const double* __restrict__ X, long ld, long rows, long K, const int* __restrict__ expo,
                                    int8_t* __restrict__ planes) {
  const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;   // one thread = one row x one K chunk
  const long kc = blockIdx.y;
  if (r >= rows) return;
  const long nkc = K / KC, rt = r / TR;
  const int rl = (int)(r % TR), e = expo[r];
  int8_t* base = planes + ((rt * nkc + kc) * S) * (long)(TR * KC);
  for (int kk = 0; kk < KC; kk++) {
    double x = ldexp(X[r + (kc * KC + kk) * ld], -e);
    const int off = core_off(rl, kk);
#pragma unroll
    for (int s = 0; s < S; s++) {
      x *= 128.0;
      const double d = trunc(x);
      x -= d;
      base[(long)s * (TR * KC) + off] = (int8_t)(int)d;
    }
  }
}

// ---- 2. the GEMM ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 1)
ozaki_gemm_kernel(const int8_t* __restrict__ Ap, const int* __restrict__ eA, const int8_t* __restrict__ Bp,
                  const int* __restrict__ eB, double* __restrict__ C, long ldc, int nkc) {
  extern __shared__ __align__(1024) unsigned char smem[];          // STAGES x [S x A_PLANE | S x B_PLANE]
  __shared__ __align__(8) uint64_t full[STAGES], empty[STAGES], accum_bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rt = blockIdx.x, ct = blockIdx.y;                       // 128-row tile of A, 64-row tile of B
  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(&accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // the MMA warp owns the TMEM allocation (all 512 columns: 8 groups x 64)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t taddr = tmem_base;

  if (warp == 0) {
    // ================= producer: 16 contiguous blocks per chunk =================================================
    const int8_t* a_src = Ap + ((long)rt * nkc) * S * A_PLANE;
    const int8_t* b_src = Bp + ((long)ct * nkc) * S * B_PLANE;
    for (int kc = 0; kc < nkc; kc++) {
      const int st = kc % STAGES;
      if (kc >= STAGES) mbar_wait(&empty[st], ((kc / STAGES) & 1) ^ 1);
      if (lane == 0) mbar_arrive_expect_tx(&full[st], STAGE_BYTES);
      __syncwarp();
      unsigned char* dst = smem + st * STAGE_BYTES;
      if (lane < S) bulk_g2s(dst + lane * A_PLANE, a_src + ((long)kc * S + lane) * A_PLANE, A_PLANE, &full[st]);
      else if (lane < 2 * S)
        bulk_g2s(dst + S * A_PLANE + (lane - S) * B_PLANE, b_src + ((long)kc * S + (lane - S)) * B_PLANE, B_PLANE, &full[st]);
    }
  } else if (warp == 1) {
    // ================= MMA issuer ==================================================================================
    constexpr uint32_t idesc = make_idesc(TM, TN);
    for (int kc = 0; kc < nkc; kc++) {
      const int st = kc % STAGES;
      mbar_wait(&full[st], (kc / STAGES) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;");
      if (lane == 0) {
        const uint32_t a0 = smem_u32(smem + st * STAGE_BYTES), b0 = a0 + S * A_PLANE;
        for (int g = 0; g < S; g++)
          for (int s = 0; s <= g; s++) {
            const int t = g - s;
            const uint64_t da = make_desc(a0 + s * A_PLANE), db = make_desc(b0 + t * B_PLANE);
            const uint32_t acc = (kc > 0 || s > 0) ? 1u : 0u;
            const uint32_t dcol = taddr + (uint32_t)(g * TN);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(dcol), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
          }
        umma_commit(&empty[st]);                 // the stage may be refilled once these MMAs have read it
        if (kc == nkc - 1) umma_commit(&accum_bar);
      }
      __syncwarp();
    }
  } else {
    // ================= epilogue warps 2..5: TMEM lane quarter = warp % 4 =============================================
    mbar_wait(&accum_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;");
    const int q = warp & 3;
    const int row = q * 32 + lane;
    double acc[TN];
#pragma unroll
    for (int j = 0; j < TN; j++) acc[j] = 0.0;
    for (int g = S - 1; g >= 0; g--) {
      const double scale = ldexp(1.0, -7 * (g + 2));
#pragma unroll
      for (int c0 = 0; c0 < TN; c0 += 32) {
        uint32_t v[32];
        const uint32_t addr = taddr + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * TN + c0);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
            "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
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
    const long gm = (long)rt * TM + row;
    const int em = eA[gm];
    for (int j = 0; j < TN; j++) {
      const long gn = (long)ct * TN + j;
      C[gm + gn * ldc] = ldexp(acc[j], em + eB[gn]);      // coalesced along m across the lanes of a warp
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(taddr));
}

int main(int argc, char** argv) {
  const long M = argc > 1 ? atol(argv[1]) : 256, N = argc > 2 ? atol(argv[2]) : 128, K = argc > 3 ? atol(argv[3]) : 1024;
  if (M % TM || N % TN || K % KC) { printf("M %% 128, N %% 64, K %% 32 must be 0\n"); return 1; }
  std::vector<double> A((size_t)M * K), B((size_t)N * K), C((size_t)M * N);
  srand(5);
  auto rnd = []() { return (rand() / (double)RAND_MAX) * 2.0 - 1.0; };
  std::vector<double> ra(M), rb(N);
  for (auto& x : ra) x = exp(8.0 * rnd());
  for (auto& x : rb) x = exp(8.0 * rnd());
  for (long k = 0; k < K; k++) for (long i = 0; i < M; i++) A[i + k * M] = ra[i] * rnd() * exp(2.0 * rnd());
  for (long k = 0; k < K; k++) for (long j = 0; j < N; j++) B[j + k * N] = rb[j] * rnd() * exp(2.0 * rnd());
  double *dA, *dB, *dC; int *eA, *eB; int8_t *pA, *pB;
  cudaMalloc(&dA, A.size() * 8); cudaMalloc(&dB, B.size() * 8); cudaMalloc(&dC, C.size() * 8);
  cudaMalloc(&eA, M * 4); cudaMalloc(&eB, N * 4);
  cudaMalloc(&pA, (size_t)M * K * S); cudaMalloc(&pB, (size_t)N * K * S);
  cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice);
  const int nkc = (int)(K / KC);
  cudaFuncSetAttribute(ozaki_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * STAGE_BYTES);
  cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
  float ms_split = 0, ms_gemm = 0;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    row_exponent_kernel<TM><<<(unsigned)((M + 127) / 128), 128>>>(dA, M, M, K, eA);
    row_exponent_kernel<TN><<<(unsigned)((N + 127) / 128), 128>>>(dB, N, N, K, eB);
    split_planes_kernel<TM><<<dim3((unsigned)((M + 127) / 128), nkc), 128>>>(dA, M, M, K, eA, pA);
    split_planes_kernel<TN><<<dim3((unsigned)((N + 127) / 128), nkc), 128>>>(dB, N, N, K, eB, pB);
    cudaEventRecord(e1);
    ozaki_gemm_kernel<<<dim3((unsigned)(M / TM), (unsigned)(N / TN)), NTHREADS, STAGES * STAGE_BYTES>>>(pA, eA, pB, eB, dC, M, nkc);
    cudaEventRecord(e2);
    cudaError_t err = cudaEventSynchronize(e2);
    cudaEventElapsedTime(&ms_split, e0, e1); cudaEventElapsedTime(&ms_gemm, e1, e2);
    printf("rep %d: %s  split %.3f ms, gemm %.3f ms = %.2f TF/s fp64-equivalent\n", rep, cudaGetErrorString(err), ms_split, ms_gemm,
           2.0 * M * N * K / (ms_gemm * 1e-3) * 1e-12);
  }
  cudaMemcpy(C.data(), dC, C.size() * 8, cudaMemcpyDeviceToHost);
  if ((double)M * N * K <= 3e8) {
    double worst = 0;
    for (long i = 0; i < M; i++)
      for (long j = 0; j < N; j++) {
        long double ref = 0, den = 0;
        for (long k = 0; k < K; k++) {
          const long double p = (long double)A[i + k * M] * (long double)B[j + k * N];
          ref += p; den += fabsl(p);
        }
        worst = fmax(worst, (double)(fabsl((long double)C[i + j * M] - ref) / den));
      }
    printf("max |err| / (|A||B|^T) = %.3e  (fp64 grade: ~1e-15)\n", worst);
  }
  return 0;
}
