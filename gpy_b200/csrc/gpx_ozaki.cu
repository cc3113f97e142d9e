// gpx_ozaki.cu — the trailing update and K^-1 = U U^T of the factor-and-invert sweep on the 5th-generation tensor cores.
//
// tcgen05.mma has no f64 kind, so fp64-grade products are formed by an Ozaki split on kind::i8 (exact s32 accumulation):
//   every row of a panel P (rows x K) is scaled by a power of two to (-1/2, 1/2) and cut into 8 signed digits of 7 bits, rounded
//   to nearest (|digit| <= 64; int8 planes),
//   P_r P_c^T = sum_{s+t <= 7} 2^(e_r + e_c - 7 (s+t+2)) D_s(r) D_t(c)^T; the 36 digit-pair products of a 32-deep k-chunk are
//   36 tcgen05.mma (128 x 64 x 32) accumulated per exponent group g = s + t in its own 64 TMEM columns (8 groups = all 512
//   columns of the SM), |digit product sum| <= 64^2 * K * 8 < 2^31 for K <= 65536; the epilogue converts the groups to
//   fp64, sums them smallest first, rescales by the row/column exponents and applies the result to the fp64 target tile.
// Replaces the same reference work as the DMMA GEMM of gpx_gemm.cu: LAPACK dpotrf / dtrtri / dpotri behind
// GPy/util/linalg.py:58,142,209-212 (see DESIGN.md §5 for the digit budget against the 1e-8 / 1e-6 tolerances).
//
// Kernel anatomy (one CTA per SM, persistent over a tile list): warp 0 = TMA producer (cp.async.bulk.tensor, 4-D tensor map
// over the pre-tiled digit planes, 4-stage mbarrier ring of 32-deep k-chunks), warp 1 = MMA issuer (one thread, tcgen05.mma
// from shared-memory descriptors, tcgen05.commit frees the stage / publishes the accumulators), warps 2..9 = epilogue
// (tcgen05.ld, one TMEM lane quarter x 32 columns each).
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <type_traits>
#include <cstdlib>
#include <cstring>

#include "gpx_common.cuh"
#include "gpx_ozaki.cuh"

#include <algorithm>
#include <vector>

namespace gpx {

constexpr int OZ_STAGES = 4;
constexpr int OZ_A_BYTES = OZ_TM * OZ_KC;                         // 4096: one digit plane of the A tile, one k-chunk
constexpr int OZ_B_BYTES = OZ_TN * OZ_KC;                         // 2048
constexpr int OZ_STAGE_BYTES = OZ_S * (OZ_A_BYTES + OZ_B_BYTES);  // 49152
constexpr int OZ_EPI_WARPS = 8;
constexpr int OZ_THREADS = (2 + OZ_EPI_WARPS) * 32;               // 320
constexpr int OZ_SMEM = OZ_STAGES * OZ_STAGE_BYTES + 1024 + 256;  // ring + alignment slack + barriers

// ---------------------------------------------------------------------------------------------------------------
// PTX helpers (tcgen05 / TMA tensor copies); mbarrier helpers come from gpx_common.cuh
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}
// mbarrier wait with back-off: the polling loop of a waiting warp takes issue slots from the one thread that feeds the tensor
// core, so the roles that wait for long (epilogue: a whole tile; producer: a free stage) sleep between polls
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity, unsigned ns) {
  for (;;) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return;
    __nanosleep(ns);
  }
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool elect_one() {   // one lane of the (converged) warp
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// shared-memory matrix descriptor: no swizzle, K-major, core matrices 8 rows x 16 B; LBO = 128 B between the two core matrices
// along K, SBO = 256 B between 8-row groups (layout pinned by tools/tcgen05_i8_check.cu on the hardware)
__device__ __forceinline__ uint64_t oz_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46);
}
// instruction descriptor: D = s32, A = B = signed 8-bit, both K-major, N at bit 17 (units of 8), M at bit 24 (units of 16)
__host__ __device__ constexpr uint32_t oz_idesc(int M, int N) {
  return (uint32_t)(2 << 4) | (uint32_t)(1 << 7) | (uint32_t)(1 << 10) | (uint32_t)((N >> 3) << 17) | (uint32_t)((M >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
               "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t addr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
      "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(addr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// 1. digit split: row exponents + 8 signed 7-bit digit planes of a panel, written in the tiled image of gpx_ozaki.cuh
// ---------------------------------------------------------------------------------------------------------------
constexpr int SPLIT_ROWS = 64;
constexpr int SPLIT_KS = 8;      // row maxima: k-slices per row (partial maxima, reduced by the digit kernel)
constexpr int SPLIT_KCB = 4;     // digit kernel: k-chunks (of 32) per CTA
// The split is two kernels so that both have thousands of CTAs: (1) partial row maxima over k-slices, (2) digits of a
// 64-row x 128-column piece per CTA. As ONE kernel (a CTA walked the whole K range of its 64 rows: 256 CTAs at N = 16384, each
// thread 2 x 256 dependent strided loads) it ran at 1.3 TB/s of HBM traffic, 211 us per 16384 x 1024 panel.
__global__ void __launch_bounds__(256) oz_rowmax_kernel(const double* __restrict__ P, long ld, long rows, int nkc_used,
                                                       double* __restrict__ amax_part) {
  __shared__ double smax[2][128];
  const int tid = threadIdx.x, rl = tid & 127, half = tid >> 7;
  const long row = (long)blockIdx.x * 128 + rl;
  const long K = (long)nkc_used * OZ_KC;
  const long kper = (K + SPLIT_KS - 1) / SPLIT_KS;
  const long k0 = (long)blockIdx.y * kper, k1 = min(K, k0 + kper);
  const double* prow = P + row;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  long k = k0 + half;
  for (; k + 6 < k1; k += 8) {
    a0 = fmax(a0, fabs(prow[k * ld]));
    a1 = fmax(a1, fabs(prow[(k + 2) * ld]));
    a2 = fmax(a2, fabs(prow[(k + 4) * ld]));
    a3 = fmax(a3, fabs(prow[(k + 6) * ld]));
  }
  for (; k < k1; k += 2) a0 = fmax(a0, fabs(prow[k * ld]));
  smax[half][rl] = fmax(fmax(a0, a1), fmax(a2, a3));
  __syncthreads();
  if (half == 0) amax_part[(long)blockIdx.y * rows + row] = fmax(smax[0][rl], smax[1][rl]);
}

// nkc = k-chunks of the plane LAYOUT (strides), nkc_used <= nkc = k-chunks actually present in this panel (a short last block)
__global__ void __launch_bounds__(256) oz_split_kernel(const double* __restrict__ P, long ld, long rows, int nkc, int nkc_used,
                                                      const double* __restrict__ amax_part, int8_t* __restrict__ planes,
                                                      double* __restrict__ scale) {
  __shared__ double sinv[SPLIT_ROWS];
  const int tid = threadIdx.x, rl = tid & (SPLIT_ROWS - 1), part = tid >> 6;
  const long row = (long)blockIdx.x * SPLIT_ROWS + rl;
  const double* prow = P + row;
  if (part == 0) {
    double amax = 0.0;
#pragma unroll
    for (int q = 0; q < SPLIT_KS; q++) amax = fmax(amax, amax_part[(long)q * rows + row]);
    double inv = 0.0, sc = 0.0;
    if (amax >= 1e-290 && amax <= 1e290) {   // rows of zeros (padding) and rows with infinities get all-zero digits
      int e = 0;
      frexp(amax, &e);                        // amax = m 2^e, m in [0.5, 1): |x| 2^-(e+1) < 1/2 for the whole row
      inv = ldexp(1.0, -(e + 1));
      sc = ldexp(1.0, e + 1 - 7);             // value = 2^(e+1) sum_s d_s 2^(-7 (s+1)); the 2^-7 of both operands folded in here
    }
    sinv[rl] = inv;
    if (blockIdx.y == 0) scale[row] = sc;
  }
  __syncthreads();
  const double inv = sinv[rl];
  const long ngrp = rows / 8;
  const int u_beg = blockIdx.y * SPLIT_KCB * 2, u_end = min(nkc_used * 2, u_beg + SPLIT_KCB * 2);
  for (int u = u_beg + part; u < u_end; u += 4) {   // unit = 16 consecutive k of one row = one 16-byte line of a core matrix
    const int kc = u >> 1, half = u & 1;
    double xs[16];
#pragma unroll
    for (int kk = 0; kk < 16; kk++) xs[kk] = prow[((long)kc * OZ_KC + half * 16 + kk) * ld];   // 16 loads in flight
    uint32_t w[OZ_S][4];
#pragma unroll
    for (int s = 0; s < OZ_S; s++) { w[s][0] = 0; w[s][1] = 0; w[s][2] = 0; w[s][3] = 0; }
#pragma unroll
    for (int kk = 0; kk < 16; kk++) {
      double x = xs[kk] * inv;                                 // exact (power of two), |x| < 1/2
#pragma unroll
      for (int s = 0; s < OZ_S; s++) {
        // round-to-nearest digits WITHOUT the conversion unit (F2I / I2F on fp64 run at a few lanes per SM):
        // x + 1.5 * 2^52 rounds x to an integer whose two's complement sits in the low word
        x *= 128.0;                                            // exact; |x| < 64 (first digit), <= 64 afterwards
        const double t = x + 6755399441055744.0;               // 1.5 * 2^52
        const int d = __double2loint(t);                       // rint(x), |d| <= 64
        x -= (t - 6755399441055744.0);                         // exact remainder, |x| <= 1/2
        w[s][kk >> 2] |= ((uint32_t)d & 0xffu) << (8 * (kk & 3));
      }
    }
    int8_t* base = planes + ((long)kc * ngrp + row / 8) * 256 + half * 128 + (row % 8) * 16;
#pragma unroll
    for (int s = 0; s < OZ_S; s++)
      *reinterpret_cast<uint4*>(base + (long)s * nkc * ngrp * 256) = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
  }
}

int launch_oz_split(const double* P, long ld, long K, OzPlanes& pl, cudaStream_t st) {
  if (K % OZ_KC || K / OZ_KC > pl.nkc) { set_error("launch_oz_split: bad panel width"); return -2; }
  const int nkc_used = (int)(K / OZ_KC);
  oz_rowmax_kernel<<<dim3((unsigned)(pl.rows / 128), SPLIT_KS), 256, 0, st>>>(P, ld, pl.rows, nkc_used, pl.amax_part);
  oz_split_kernel<<<dim3((unsigned)(pl.rows / SPLIT_ROWS), (unsigned)((nkc_used + SPLIT_KCB - 1) / SPLIT_KCB)), 256, 0, st>>>(
      P, ld, pl.rows, pl.nkc, nkc_used, pl.amax_part, pl.planes, pl.scale);
  GPX_CUDA(cudaGetLastError());
  return 0;
}

// apply one thread's row of an output tile to the fp64 target: NCOL columns, leading dimension ld. The loads of a group of 8
// columns are all issued before the first store: written as `C[j * ld] -= ...` the compiler must assume that a store may
// alias the next load (ld is a run-time value) and serialises 64 global round trips per thread (~20 us per tile, measured).
template <int NCOL>
__device__ __forceinline__ void oz_apply_row(double* __restrict__ C, long ld, const double (&acc)[NCOL], double si,
                                             const double* __restrict__ scol, int kind) {
#pragma unroll
  for (int j0 = 0; j0 < NCOL; j0 += 8) {
    double old[8];
    if (kind != OZ_LAUUM_SET) {
#pragma unroll
      for (int j = 0; j < 8; j++) old[j] = C[(long)(j0 + j) * ld];
    }
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const double v = acc[j0 + j] * (si * __ldg(scol + j0 + j));
      C[(long)(j0 + j) * ld] = kind == OZ_UPDATE ? old[j] - v : (kind == OZ_LAUUM_ACC ? old[j] + v : v);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 2. the GEMM
// ---------------------------------------------------------------------------------------------------------------
// one 32-deep k-chunk: the ND (ND + 1) / 2 digit-pair products, exponent group g = s + t in TMEM columns [64 g, 64 g + 64).
// a_lo = low descriptor word (address field) of digit plane 0 of the A tile in this stage; the B planes follow the 8 A planes.
template <int ND, int G_BEG, int G_END>
__device__ __forceinline__ void oz_issue_groups(uint32_t taddr, uint32_t a_lo, uint32_t acc0) {
  constexpr uint32_t idesc = oz_idesc(OZ_TM, OZ_TN);
  constexpr uint64_t hi = ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46);   // LBO, SBO, version
  const uint32_t b_lo = a_lo + (uint32_t)((OZ_S * OZ_A_BYTES) >> 4);
#pragma unroll
  for (int g = G_BEG; g < G_END; g++) {
#pragma unroll
    for (int s = 0; s <= g; s++) {
      const uint64_t da = hi | (uint64_t)(a_lo + (uint32_t)(s * (OZ_A_BYTES >> 4)));
      const uint64_t db = hi | (uint64_t)(b_lo + (uint32_t)((g - s) * (OZ_B_BYTES >> 4)));
      umma_i8(taddr + (uint32_t)(g * OZ_TN), da, db, idesc, s > 0 ? 1u : acc0);
    }
  }
}

__global__ void __launch_bounds__(OZ_THREADS, 1)
oz_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const OzParams p) {
  extern __shared__ unsigned char oz_smem_raw[];
  unsigned char* ring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(oz_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + OZ_STAGES * OZ_STAGE_BYTES);
  uint64_t* empty = full + OZ_STAGES;
  uint64_t* tmem_full = empty + OZ_STAGES;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < OZ_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, OZ_EPI_WARPS);
    fence_mbar_init();
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
  }
  if (warp == 1) {   // the MMA warp owns the TMEM allocation: all 512 columns (8 exponent groups x 64)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t taddr = *tmem_slot;
  const int nkc = p.nkc;
  // a CTA owns p.tpc CONSECUTIVE tiles of the list (neighbours in the banded order share operand panels in L2) and then
  // retires: SM slots come free every few tiles, so the high-priority side stream of the sweep (diagonal block, panel of the
  // next step) gets onto the machine while this launch is still running — a fully persistent grid would shut it out
  const int ti_beg = blockIdx.x * p.tpc, ti_end = min(p.ntiles, ti_beg + p.tpc);

  if (warp == 0) {
    // ================= TMA producer: lane s < nd loads digit plane s of both operands ================================
    uint32_t it = 0;
    for (int ti = ti_beg; ti < ti_end; ti++) {
      const uint32_t t = p.tiles[ti];
      const int r = t & 0xfff, c64 = (t >> 12) & 0x1fff;
      const int nd = ((t >> 27) & 1) ? p.dig_up : p.dig_lo;
      for (int kc = 0; kc < nkc; kc++, it++) {
        const int st = it % OZ_STAGES;
        if (p.dbg & 16) mbar_wait(&empty[st], ((it / OZ_STAGES) & 1) ^ 1);
        else mbar_wait_backoff(&empty[st], ((it / OZ_STAGES) & 1) ^ 1, 64);
        if (p.dbg & 2) { if (lane == 0) mbar_arrive(&full[st]); continue; }
        // complete_tx of a copy may precede the expect_tx below: the phase cannot complete before lane 0's arrival
        unsigned char* dst = ring + st * OZ_STAGE_BYTES;
        if (lane < nd) {
          tma_load_4d(dst + lane * OZ_A_BYTES, &mapA, 0, r * (OZ_TM / 8), kc, lane, &full[st]);
          tma_load_4d(dst + OZ_S * OZ_A_BYTES + lane * OZ_B_BYTES, &mapB, 0, c64 * (OZ_TN / 8), kc, lane, &full[st]);
        }
        if (lane == 0) mbar_arrive_expect_tx(&full[st], (uint32_t)nd * (OZ_A_BYTES + OZ_B_BYTES));
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer ====================================================================================
    // The issue loop must sustain one tcgen05.mma per ~48 clk from ONE thread: descriptors are not rebuilt per instruction
    // (the first version spent ~100 clk of uniform-datapath arithmetic per MMA and ran at half the shared-memory bound);
    // the low descriptor word of plane s is the stage's base word + s * (plane bytes >> 4), the 36 instructions of a
    // k-chunk are straight-line code (template on the digit count). The whole warp executes the loop convergently and an
    // elected lane issues: inside an `if (lane == 0)` region the compiler can lose track of warp-uniformity and then
    // feeds every MMA operand through R2UR from vector registers (measured: 1.3x slower issue).
    {   // the WHOLE warp runs the (warp-uniform) control flow; one elected lane issues MMAs and commits
      uint32_t it = 0, tl = 0;
      const uint32_t ring_lo = (smem_u32(ring) & 0x3FFFF) >> 4;
      for (int ti = ti_beg; ti < ti_end; ti++, tl++) {
        const uint32_t t = p.tiles[ti];
        const int nd = ((t >> 27) & 1) ? p.dig_up : p.dig_lo;
        mbar_wait(tmem_empty, (tl & 1) ^ 1);     // the epilogue has read the previous tile out of TMEM
        tc_fence_after();
        for (int kc = 0; kc < nkc; kc++, it++) {
          const int st = it % OZ_STAGES;
          mbar_wait(&full[st], (it / OZ_STAGES) & 1);
          tc_fence_after();
          const uint32_t a_lo = ring_lo + (uint32_t)st * (OZ_STAGE_BYTES >> 4);
          const uint32_t acc0 = kc > 0 ? 1u : 0u;
          if (elect_one()) {
            if (!(p.dbg & 1)) {
              if (nd == 8) oz_issue_groups<8, 0, 8>(taddr, a_lo, acc0);
              else if (nd == 7) oz_issue_groups<7, 0, 7>(taddr, a_lo, acc0);
              else if (nd == 6) oz_issue_groups<6, 0, 6>(taddr, a_lo, acc0);
              else if (nd == 5) oz_issue_groups<5, 0, 5>(taddr, a_lo, acc0);
              else oz_issue_groups<4, 0, 4>(taddr, a_lo, acc0);
            }
            umma_commit(&empty[st]);              // the stage may be refilled once these MMAs have read it
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(tmem_full);  // accumulators of this tile complete
        __syncwarp();
      }
    }
  } else {
    // ================= epilogue warps: TMEM lane quarter = warp % 4, column half = (warp - 2) / 4 ====================
    const int q = warp & 3, h = (warp - 2) >> 2;
    uint32_t tl = 0;
    for (int ti = ti_beg; ti < ti_end; ti++, tl++) {
      const uint32_t t = p.tiles[ti];
      const int r = t & 0xfff, c64 = (t >> 12) & 0x1fff, kind = (t >> 25) & 3;
      const int nd = ((t >> 27) & 1) ? p.dig_up : p.dig_lo;
      if (p.dbg & 8) mbar_wait(tmem_full, tl & 1);
      else mbar_wait_backoff(tmem_full, tl & 1, 128);
      tc_fence_after();
      double acc[32];
#pragma unroll
      for (int j = 0; j < 32; j++) acc[j] = 0.0;
      for (int g = ((p.dbg & 4) ? 0 : nd) - 1; g >= 0; g--) {         // smallest magnitude first
        uint32_t v[32];
        tmem_ld32(taddr + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * OZ_TN + h * 32), v);
        const double sc = __longlong_as_double((long long)(1023 - 7 * g) << 52);   // 2^(-7 g)
        // s32 -> fp64 without the conversion unit (I2F.F64 is a few lanes per SM: it made the TMEM drain ~15k clk per tile):
        // the bit pattern {0x43300000, v ^ 0x80000000} is the double 2^52 + 2^31 + v, one exact DADD takes the offset off
#pragma unroll
        for (int j = 0; j < 32; j++)
          acc[j] = fma(__hiloint2double(0x43300000, (int)(v[j] ^ 0x80000000u)) - 4503601774854144.0, sc, acc[j]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty);     // TMEM may be overwritten by the next tile's MMAs
      if (p.dbg & 4) continue;
      const long gi = (long)r * OZ_TM + q * 32 + lane;
      const long gj0 = (long)c64 * OZ_TN + h * 32;
      const double si = p.scale[gi];
      if (kind == OZ_UPDATE) oz_apply_row<32>(p.S + gi + gj0 * p.lds, p.lds, acc, si, p.scale + gj0, kind);
      else oz_apply_row<32>(p.Kinv + gi + gj0 * p.ldk, p.ldk, acc, si, p.scale + gj0, kind);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(taddr));
}

// ---------------------------------------------------------------------------------------------------------------
// 2b. the two-pass variant: 128 x 128 output tiles, tcgen05.mma 128 x 128 x 32.
// At N = 64 the tensor core re-reads the 4 KB A plane for every 2 KB B plane and the 128 B/clk shared-memory read port
// caps it at 2/3 of its rate (48 clk instead of 32 per MMA, tools/microbench_ozaki_pattern.cu); at N = 128 it runs at the
// full 8192 MAC/clk. Eight exponent groups x 128 columns do not fit the 512 TMEM columns, so a tile is computed in two
// passes over K: first the low-order groups 4..7 (26 digit pairs, all planes), drained into fp64 registers, then the
// high-order groups 0..3 (10 pairs, planes 0..3 only). Operand traffic per output element is 1.5x that of the one-pass
// kernel (24 instead of 16 plane-tile loads per 128 x 128 x 32), the MMA time 2/3.
// ---------------------------------------------------------------------------------------------------------------
constexpr int OZ2_TN = 128;
constexpr int OZ2_STAGES = 3;
constexpr int OZ2_PLANE = OZ_TM * OZ_KC;                          // 4096 bytes: one digit plane of a 128-row tile, one k-chunk
constexpr int OZ2_STAGE_BYTES = 2 * OZ_S * OZ2_PLANE;             // 65536: up to 8 A planes + 8 B planes
constexpr int OZ2_SMEM = OZ2_STAGES * OZ2_STAGE_BYTES + 1024 + 256;
constexpr int OZ2_THREADS = 384;                                   // three warpgroups (setmaxnreg works per warpgroup)

// groups [G_BEG, G_END) of one k-chunk, highest group first; group g accumulates in TMEM columns [(g - G_BASE) * 128, +128)
template <int G, int G_BASE>
__device__ __forceinline__ void oz2_issue_group(uint32_t taddr, uint32_t a_lo, uint32_t acc0) {
  constexpr uint32_t idesc = oz_idesc(OZ_TM, OZ2_TN);
  constexpr uint64_t hi = ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46);
  const uint32_t b_lo = a_lo + (uint32_t)((OZ_S * OZ2_PLANE) >> 4);
#pragma unroll
  for (int s = 0; s <= G; s++) {
    const uint64_t da = hi | (uint64_t)(a_lo + (uint32_t)(s * (OZ2_PLANE >> 4)));
    const uint64_t db = hi | (uint64_t)(b_lo + (uint32_t)((G - s) * (OZ2_PLANE >> 4)));
    umma_i8(taddr + (uint32_t)((G - G_BASE) * OZ2_TN), da, db, idesc, s > 0 ? 1u : acc0);
  }
}
template <int G_BEG, int G_END, int G_BASE>
__device__ __forceinline__ void oz2_issue_groups(uint32_t taddr, uint32_t a_lo, uint32_t acc0) {
  if constexpr (G_END > G_BEG) {
    oz2_issue_group<G_END - 1, G_BASE>(taddr, a_lo, acc0);
    oz2_issue_groups<G_BEG, G_END - 1, G_BASE>(taddr, a_lo, acc0);
  }
}
// First k-chunk of a pass: the accumulator of group g (TMEM slot g - G_BASE) is overwritten as soon as the epilogue has drained
// THAT slot of the previous pass (it drains the slots in the same descending order), not the whole TMEM: the tensor core waits
// for one slot's drain (~670 clk) instead of four. Bit s of usebits is the parity of the number of passes that used slot s (the mbarrier phase to wait for).
template <int G_BEG, int G_END, int G_BASE>
__device__ __forceinline__ void oz2_first_chunk(uint32_t taddr, uint32_t a_lo, uint64_t* tmem_empty, uint32_t& usebits, bool issue) {
  if constexpr (G_END > G_BEG) {
    constexpr int slot = G_END - 1 - G_BASE;
    mbar_wait(&tmem_empty[slot], ((usebits >> slot) & 1u) ^ 1u);
    usebits ^= 1u << slot;
    tc_fence_after();
    if (elect_one()) {
      if (issue) oz2_issue_group<G_END - 1, G_BASE>(taddr, a_lo, 0u);
    }
    __syncwarp();
    oz2_first_chunk<G_BEG, G_END - 1, G_BASE>(taddr, a_lo, tmem_empty, usebits, issue);
  }
}
// k-chunks of a tile: the whole panel width, except OZ_PANEL tiles (triangular B: columns of tile c' see k < (c' + 1) * 128)
__device__ __forceinline__ int oz2_tile_nkc(uint32_t t, int nkc) {
  return ((t >> 25) & 3) == OZ_PANEL ? min(nkc, (int)(((t >> 12) & 0x1fff) + 1) * (OZ2_TN / OZ_KC)) : nkc;
}
// Warp roles: warps 0..7 = epilogue (TMEM lane quarter = warp % 4, column half = warp / 4; two warpgroups that raise their
// register budget to 224 with setmaxnreg: 64 fp64 accumulators per thread live across the two passes), warp 8 = TMA
// producer, warp 9 = MMA issuer, warps 10-11 idle (the third warpgroup hands its registers over: 56 each).
__global__ void __launch_bounds__(OZ2_THREADS, 1)
oz_gemm2_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const OzParams p) {
  extern __shared__ unsigned char oz_smem_raw[];
  unsigned char* ring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(oz_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + OZ2_STAGES * OZ2_STAGE_BYTES);
  uint64_t* empty = full + OZ2_STAGES;
  uint64_t* tmem_full = empty + OZ2_STAGES;
  uint64_t* tmem_empty = tmem_full + 1;                    // one per TMEM slot (128 accumulator columns)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < OZ2_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
#pragma unroll
    for (int s = 0; s < 4; s++) mbar_init(&tmem_empty[s], OZ_EPI_WARPS);
    fence_mbar_init();
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
  }
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t taddr = *tmem_slot;
  const int nkc = p.nkc;
  const int ti_beg = blockIdx.x * p.tpc, ti_end = min(p.ntiles, ti_beg + p.tpc);

  if (warp >= 8) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 8) {
    // ================= TMA producer =================================================================================
    uint32_t it = 0;
    for (int ti = ti_beg; ti < ti_end; ti++) {
      const uint32_t t = p.tiles[ti];
      const int r = t & 0xfff, c = (t >> 12) & 0x1fff;
      const int nd = ((t >> 27) & 1) ? p.dig_up : p.dig_lo;   // >= 5 here: both passes always run, so that the stage
      const int nkc_t = oz2_tile_nkc(t, nkc);
      for (int pass = 0; pass < 2; pass++) {                    // counter stays warp-uniform (descriptors in uniform registers)
        const int npl = pass == 0 ? nd : 4;        // low-order groups need every plane, groups 0..3 only planes 0..3
        for (int kc = 0; kc < nkc_t; kc++, it++) {
          const int st = it % OZ2_STAGES;
          if (p.dbg & 16) mbar_wait(&empty[st], ((it / OZ2_STAGES) & 1) ^ 1);
          else mbar_wait_backoff(&empty[st], ((it / OZ2_STAGES) & 1) ^ 1, 64);
          if (p.dbg & 2) { if (lane == 0) mbar_arrive(&full[st]); continue; }
          unsigned char* dst = ring + st * OZ2_STAGE_BYTES;
          if (lane < npl) {
            tma_load_4d(dst + lane * OZ2_PLANE, &mapA, 0, r * (OZ_TM / 8), kc, lane, &full[st]);
            tma_load_4d(dst + (OZ_S + lane) * OZ2_PLANE, &mapB, 0, c * (OZ2_TN / 8), kc, lane, &full[st]);
          }
          if (lane == 0) mbar_arrive_expect_tx(&full[st], (uint32_t)npl * 2 * OZ2_PLANE);
        }
      }
    }
  } else if (warp == 9) {
    // ================= MMA issuer (one thread) ======================================================================
    {   // the WHOLE warp runs the (warp-uniform) control flow; one elected lane issues MMAs and commits
      uint32_t ph = 0;
      uint32_t usebits = 0u;
      int st = 0;
      const uint32_t ring_lo = (smem_u32(ring) & 0x3FFFF) >> 4;
      const bool issue = !(p.dbg & 1);
      for (int ti = ti_beg; ti < ti_end; ti++) {
        const uint32_t t = p.tiles[ti];
        const int nd = ((t >> 27) & 1) ? p.dig_up : p.dig_lo;
        const int nkc_t = oz2_tile_nkc(t, nkc);
        // the two passes are two copies of the loop (compile-time pass): with a run-time pass variable selecting the MMA
        // block the compiler keeps the descriptors in vector registers (R2UR per operand, ~1.3x slower issue)
        auto run_pass = [&](auto pass_tag) {
          constexpr int pass = decltype(pass_tag)::value;
          for (int kc = 0; kc < nkc_t; kc++) {
            mbar_wait(&full[st], ph);
            tc_fence_after();
            const uint32_t a_lo = ring_lo + (uint32_t)st * (OZ2_STAGE_BYTES >> 4);
            if (kc == 0) {   // slot by slot behind the epilogue's drain of the previous pass
              if (pass == 1) oz2_first_chunk<0, 4, 0>(taddr, a_lo, tmem_empty, usebits, issue);
              else if (nd == 8) oz2_first_chunk<4, 8, 4>(taddr, a_lo, tmem_empty, usebits, issue);
              else if (nd == 7) oz2_first_chunk<4, 7, 4>(taddr, a_lo, tmem_empty, usebits, issue);
              else if (nd == 6) oz2_first_chunk<4, 6, 4>(taddr, a_lo, tmem_empty, usebits, issue);
              else oz2_first_chunk<4, 5, 4>(taddr, a_lo, tmem_empty, usebits, issue);
              if (elect_one()) umma_commit(&empty[st]);
            } else if (elect_one()) {
              if (issue) {
                if (pass == 1) oz2_issue_groups<0, 4, 0>(taddr, a_lo, 1u);          // groups 0 .. 3
                else if (nd == 8) oz2_issue_groups<4, 8, 4>(taddr, a_lo, 1u);       // low-order groups 4 .. nd-1
                else if (nd == 7) oz2_issue_groups<4, 7, 4>(taddr, a_lo, 1u);
                else if (nd == 6) oz2_issue_groups<4, 6, 4>(taddr, a_lo, 1u);
                else oz2_issue_groups<4, 5, 4>(taddr, a_lo, 1u);
              }
              umma_commit(&empty[st]);
            }
            __syncwarp();
            if (++st == OZ2_STAGES) { st = 0; ph ^= 1; }
          }
          if (elect_one()) umma_commit(tmem_full);
          __syncwarp();
        };
        run_pass(std::integral_constant<int, 0>{});
        run_pass(std::integral_constant<int, 1>{});
      }
    }
  }
  } else {
    // ================= epilogue warps: TMEM lane quarter = warp % 4, 64 of the 128 columns each =======================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int q = warp & 3, h = warp >> 2;
    uint32_t hs = 0;
    for (int ti = ti_beg; ti < ti_end; ti++) {
      const uint32_t t = p.tiles[ti];
      const int r = t & 0xfff, c = (t >> 12) & 0x1fff, kind = (t >> 25) & 3;
      const int nd = ((t >> 27) & 1) ? p.dig_up : p.dig_lo;
      double acc[64];
#pragma unroll
      for (int j = 0; j < 64; j++) acc[j] = 0.0;
      for (int pass = 0; pass < 2; pass++, hs++) {
        if (p.dbg & 8) mbar_wait(tmem_full, hs & 1);
        else mbar_wait_backoff(tmem_full, hs & 1, 128);
        tc_fence_after();
        const int gbase = pass == 0 ? 4 : 0, gtop = pass == 0 ? nd : 4;
        for (int g = gtop - 1; g >= gbase; g--) {   // smallest magnitude first (pass 0 before pass 1)
          const double sc = __longlong_as_double((long long)(1023 - 7 * g) << 52);   // 2^(-7 g)
          uint32_t v0[32], v1[32];
          if (!(p.dbg & 4)) {
            const uint32_t ta = taddr + ((uint32_t)(q * 32) << 16) + (uint32_t)((g - gbase) * OZ2_TN + h * 64);
            tmem_ld32(ta, v0);
            tmem_ld32(ta + 32, v1);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0 && !(p.dbg & 32)) mbar_arrive(&tmem_empty[g - gbase]);   // this slot may take the next pass's MMAs: the fp64 sum follows
          if (!(p.dbg & 4)) {
#pragma unroll
            for (int j = 0; j < 32; j++)
              acc[j] = fma(__hiloint2double(0x43300000, (int)(v0[j] ^ 0x80000000u)) - 4503601774854144.0, sc, acc[j]);
#pragma unroll
            for (int j = 0; j < 32; j++)
              acc[32 + j] = fma(__hiloint2double(0x43300000, (int)(v1[j] ^ 0x80000000u)) - 4503601774854144.0, sc, acc[32 + j]);
          }
        }
        if (p.dbg & 32) {   // measurement: release the whole TMEM only after the full drain (the behaviour before the per-slot barriers)
          __syncwarp();
          if (lane == 0)
            for (int g = gtop - 1; g >= gbase; g--) mbar_arrive(&tmem_empty[g - gbase]);
        }
      }
      if (p.dbg & 4) continue;
      const long gi = (long)r * OZ_TM + q * 32 + lane;
      const long gj0 = (long)c * OZ2_TN + h * 64;
      const double si = p.scale[gi];
      if (kind == OZ_UPDATE) oz_apply_row<64>(p.S + gi + gj0 * p.lds, p.lds, acc, si, p.scale + gj0, kind);
      else if (kind == OZ_PANEL) {
        oz_apply_row<64>(p.P + gi + gj0 * p.ldp, p.ldp, acc, si, p.scaleB + gj0, OZ_LAUUM_SET);
        // the panel rows also take their final place in the workspace (U block column above, L panel below the diagonal
        // block): its digit planes were taken before this launch, nothing reads the fp64 block column any more
        if (p.Pfinal) oz_apply_row<64>(p.Pfinal + gi + gj0 * p.lds, p.lds, acc, si, p.scaleB + gj0, OZ_LAUUM_SET);
      }
      else oz_apply_row<64>(p.Kinv + gi + gj0 * p.ldk, p.ldk, acc, si, p.scale + gj0, kind);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(taddr));
}

// ---------------------------------------------------------------------------------------------------------------
// tile lists of the sweep (host): per panel step the launches U0 | U1 | U2 (+ the K^-1 tiles of the step)
// ---------------------------------------------------------------------------------------------------------------
// tiles in bands of 8 row tiles x 16 column tiles (64 wide): the ~148 tiles in flight share 8 A panels and 16 B panels in L2
// (column tiles: cw per 128 columns - two 64-wide tiles for the one-pass kernel, one 128-wide tile for the two-pass kernel)
template <class Valid, class Emit>
static void oz_banded(const std::vector<int>& rows, int ct_beg, int ct_end, int cw, Valid valid, Emit emit) {
  for (size_t b = 0; b < rows.size(); b += 8)
    for (int cc = ct_beg; cc < ct_end; cc += 8 * cw)
      for (size_t i = b; i < std::min(rows.size(), b + 8); i++)
        for (int ct = cc; ct < std::min(ct_end, cc + 8 * cw); ct++)
          if (valid(rows[i], ct / cw)) emit(rows[i], ct);
}

// own_G > 1 (memory-distributed multi-GPU sweep): only the row tiles of the block rows this rank owns (block-cyclic over the
// panels of NB rows, like the DMMA update of gpx_dist.cu) are listed; every tile of the matrix is in exactly one rank's list.
void oz_build_lists(long Npad, long NB, int cw, int own_G, int own_g, std::vector<uint32_t>& tiles, std::vector<OzStep>& steps) {
  const int nt = (int)(Npad / TILE), nbt_full = (int)(NB / TILE);
  auto mine = [&](int r) { return own_G <= 1 || ((r / nbt_full) % own_G) == own_g; };
  tiles.clear();
  steps.clear();
  for (long o = 0; o < Npad; o += NB) {
    const long nb = std::min(NB, Npad - o);
    const int kt0 = (int)(o / TILE), kt1 = kt0 + (int)(nb / TILE);
    const int next_nbt = kt1 < nt ? (int)(std::min(NB, Npad - (o + nb)) / TILE) : 0;
    OzStep st;
    auto emit_update = [&](int cbeg, int cend) {   // S(r, c) -= P_r P_c^T, c in [cbeg, cend), r in [0, kt1) U [c, nt)
      std::vector<int> rows;
      for (int r = 0; r < kt1; r++) if (mine(r)) rows.push_back(r);
      for (int r = cbeg; r < nt; r++) if (mine(r)) rows.push_back(r);
      oz_banded(rows, cw * cbeg, cw * cend, cw, [&](int r, int cc) { return r < kt1 || cc <= r; },
                [&](int r, int ct) { tiles.push_back(oz_tile(r, ct, OZ_UPDATE, r < kt1 ? 1 : 0)); });
    };
    auto count_up = [&](int off, int n) { int u = 0; for (int i = off; i < off + n; i++) u += (tiles[i] >> 27) & 1; return u; };
    // U0: the tiles of the NEXT diagonal block (rows and columns of block k+1): all that D(k+1) waits for
    st.u0_off = (int)tiles.size();
    if (kt1 < nt) {
      std::vector<int> rows;
      for (int r = kt1; r < kt1 + next_nbt; r++) if (mine(r)) rows.push_back(r);
      oz_banded(rows, cw * kt1, cw * (kt1 + next_nbt), cw, [&](int r, int cc) { return cc <= r; },
                [&](int r, int ct) { tiles.push_back(oz_tile(r, ct, OZ_UPDATE, 0)); });
    }
    st.u0_n = (int)tiles.size() - st.u0_off;
    st.u0_up = 0;
    // U1: the rest of block column k+1 (rows above the block and below it)
    st.u1_off = (int)tiles.size();
    if (kt1 < nt) {
      std::vector<int> rows;
      for (int r = 0; r < kt1; r++) if (mine(r)) rows.push_back(r);
      for (int r = kt1 + next_nbt; r < nt; r++) if (mine(r)) rows.push_back(r);
      oz_banded(rows, cw * kt1, cw * (kt1 + next_nbt), cw, [&](int r, int cc) { return r < kt1 || cc <= r; },
                [&](int r, int ct) { tiles.push_back(oz_tile(r, ct, OZ_UPDATE, r < kt1 ? 1 : 0)); });
    }
    st.u1_n = (int)tiles.size() - st.u1_off;
    st.u1_up = count_up(st.u1_off, st.u1_n);
    st.u2_off = (int)tiles.size();
    if (kt1 + next_nbt < nt) emit_update(kt1 + next_nbt, nt);
    st.u2_upd = (int)tiles.size() - st.u2_off;
    st.u2_upd_up = count_up(st.u2_off, st.u2_upd);
    {   // K^-1(r, c) (+)= P_r P_c^T for c <= r < kt1: rows of block k see their first contribution at this step
      std::vector<int> rows;
      for (int r = 0; r < kt1; r++) if (mine(r)) rows.push_back(r);
      oz_banded(rows, 0, cw * kt1, cw, [&](int r, int cc) { return cc <= r; },
                [&](int r, int ct) { tiles.push_back(oz_tile(r, ct, r >= kt0 ? OZ_LAUUM_SET : OZ_LAUUM_ACC, 1)); });
    }
    st.u2_n = (int)tiles.size() - st.u2_off;
    st.u2_up = count_up(st.u2_off, st.u2_n);
    {   // panel GEMM on the tensor cores: P(r, c') for the rows outside the diagonal block and block k+1, c' < nb / 128
      st.pan_off = (int)tiles.size();
      const int nbt = kt1 - kt0;
      std::vector<int> rows;
      for (int r = 0; r < kt0; r++) if (mine(r)) rows.push_back(r);
      for (int r = kt1 + next_nbt; r < nt; r++) if (mine(r)) rows.push_back(r);
      for (size_t b = 0; b < rows.size(); b += 8)              // bands of 8 row tiles; longest k-range (last column tile) first
        for (int cc = nbt - 1; cc >= 0; cc--)
          for (size_t i = b; i < std::min(rows.size(), b + 8); i++) tiles.push_back(oz_tile(rows[i], cc, OZ_PANEL, rows[i] < kt0 ? 1 : 0));
      st.pan_n = (int)tiles.size() - st.pan_off;
      st.pan_up = count_up(st.pan_off, st.pan_n);
    }
    steps.push_back(st);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

int oz_init() {
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    GPX_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || !fn) { set_error("cuTensorMapEncodeTiled is not available in this driver"); return -1; }
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  GPX_CUDA(cudaFuncSetAttribute(oz_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM));
  GPX_CUDA(cudaFuncSetAttribute(oz_gemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ2_SMEM));
  return 0;
}

static int oz_make_map(CUtensorMap* map, const OzPlanes& pl, int box_rows) {
  const cuuint64_t ngrp = (cuuint64_t)(pl.rows / 8);
  cuuint64_t dims[4] = {256, ngrp, (cuuint64_t)pl.nkc, (cuuint64_t)OZ_S};
  cuuint64_t strides[3] = {256, ngrp * 256, (cuuint64_t)pl.nkc * ngrp * 256};
  cuuint32_t box[4] = {256, (cuuint32_t)(box_rows / 8), 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, pl.planes, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")"); return -1; }
  return 0;
}

int oz_planes_alloc(OzPlanes& pl, long rows, long K) {
  if (rows % OZ_TM || K % OZ_KC) { set_error("oz_planes_alloc: rows % 128 or K % 32"); return -2; }
  pl.rows = rows;
  pl.nkc = (int)(K / OZ_KC);
  GPX_CUDA(cudaMalloc(&pl.planes, (size_t)OZ_S * rows * K));
  GPX_CUDA(cudaMalloc(&pl.scale, (size_t)rows * 8));
  GPX_CUDA(cudaMalloc(&pl.amax_part, (size_t)rows * 8 * 8));   // SPLIT_KS partial row maxima
  if (oz_make_map(&pl.mapA, pl, OZ_TM) || oz_make_map(&pl.mapB, pl, OZ_TN)) return -1;
  return 0;
}

void oz_planes_free(OzPlanes& pl) {
  if (pl.planes) cudaFree(pl.planes);
  if (pl.scale) cudaFree(pl.scale);
  if (pl.amax_part) cudaFree(pl.amax_part);
  pl.planes = nullptr; pl.scale = nullptr; pl.amax_part = nullptr; pl.rows = 0; pl.nkc = 0;
}

int launch_oz_gemm(const OzPlanes& pl, const OzParams& p_in, int num_sms, cudaStream_t st, const OzPlanes* plB) {
  if (p_in.ntiles <= 0) return 0;
  OzParams p = p_in;
  // default: 8 narrow / 4 wide tiles per CTA when there is enough work (measured: 1 -> 2 -> 4 -> 8 tiles per CTA = 118 / 104 / 98 / 91 ms)
  if (p.tpc <= 0) p.tpc = std::max(1, std::min(p.wide ? 4 : 8, p.ntiles / std::max(1, num_sms)));
  const int grid = (p.ntiles + p.tpc - 1) / p.tpc;
  if (p.wide && (p.dig_lo < 5 || p.dig_up < 5)) { set_error("the two-pass kernel needs at least 5 digits"); return -2; }
  if (plB && !p.wide) { set_error("OZ_PANEL tiles need the two-pass kernel"); return -2; }
  if (p.wide) oz_gemm2_kernel<<<grid, OZ2_THREADS, OZ2_SMEM, st>>>(pl.mapA, plB ? plB->mapA : pl.mapA, p);
  else oz_gemm_kernel<<<grid, OZ_THREADS, OZ_SMEM, st>>>(pl.mapA, pl.mapB, p);
  GPX_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// 3. gradient reductions from the stored K^-1 (same sums as the fused LAUUM epilogue of gpx_gemm.cu):
//    dL_dK = 1/2 (alpha alpha^T - P K^-1)  (exact_gaussian_inference.py:70), tr -> noise (:72, gaussian.py:78-79),
//    sum K o dL_dK / variance (stationary.py:199), ARD / iso lengthscale sums (:202-213, 225-243). One CTA per lower
//    128 x 128 tile; the thread-mapped index is the ROW (contiguous in the column-major K^-1), 64 columns per thread.
// ---------------------------------------------------------------------------------------------------------------
template <int DREG>
__global__ void __launch_bounds__(256) grad_kinv_kernel(GradKinvParams p) {
  extern __shared__ __align__(128) unsigned char gk_smem[];
  const int D = p.kp.D, P = p.P;
  double* sXc = reinterpret_cast<double*>(gk_smem);   // [D][128] column points
  double* sSc = sXc + (size_t)D * TILE;               // [128]
  double* sAc = sSc + TILE;                           // [P][128]
  double* sRed = sAc + (size_t)P * TILE;              // [8 warps][nred]
  // lower tiles only: blockIdx.x enumerates (r, c), c <= r
  const int csplit = p.csplit, tileidx = (int)blockIdx.x / csplit, sidx = (int)blockIdx.x % csplit;
  int r = (int)((sqrtf(8.f * (float)tileidx + 1.f) - 1.f) * 0.5f);
  while (r * (r + 1) / 2 > tileidx) --r;
  while ((r + 1) * (r + 2) / 2 <= tileidx) ++r;
  const int c = tileidx - r * (r + 1) / 2;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int idx = tid; idx < D * TILE; idx += 256) sXc[idx] = p.XsT[(long)(idx / TILE) * p.ldx + (long)c * TILE + idx % TILE];
  for (int idx = tid; idx < P * TILE; idx += 256) sAc[idx] = p.alpha[(long)(idx / TILE) * p.ldx + (long)c * TILE + idx % TILE];
  if (tid < TILE) sSc[tid] = p.sq[(long)c * TILE + tid];
  __syncthreads();
  const int il = tid & (TILE - 1), half = tid >> 7;
  const long gi = (long)r * TILE + il;
  const bool ard = p.kp.ard != 0;
  const int nl = ard ? D : 1, nred = nl + 2;
  double xi[DREG], gq[DREG], ai[MAX_P];
#pragma unroll
  for (int q = 0; q < DREG; q++) { gq[q] = 0.0; xi[q] = q < D ? p.XsT[(long)q * p.ldx + gi] : 0.0; }
#pragma unroll
  for (int q = 0; q < MAX_P; q++) ai[q] = q < P ? p.alpha[(long)q * p.ldx + gi] : 0.0;
  const double si = p.sq[gi];
  const double w = (r > c) ? 2.0 : 1.0;               // strictly-lower tiles stand for their mirror image as well
  const double variance = p.kp.variance, inv_ls = p.kp.inv_ls_iso;
  const int kind = p.kp.kind;
  double gvar = 0.0, giso = 0.0, gnoise = 0.0;
  const double* kcol = p.Kinv + gi + ((long)c * TILE + half * 64) * p.ld;
  if (gi < p.N) {
    const int jj_beg = sidx * (64 / csplit);
    const int jj_end = (int)min((long)(jj_beg + 64 / csplit), p.N - (long)c * TILE - half * 64);   // columns beyond N are padding
    // two columns in flight per thread: one column is a single dependent chain (dot product -> exp -> reductions) and the
    // 16 warps an SM holds at 92 registers leave the fp64 pipe mostly idle (1.8 ms for the 134 M elements of N = 16384)
#pragma unroll 2
    for (int jj = jj_beg; jj < jj_end; jj++) {
      const int jl = half * 64 + jj;
      const long gj = (long)c * TILE + jl;
      const double kinv = kcol[(long)jj * p.ld];
      double dot = 0.0;
#pragma unroll
      for (int q = 0; q < DREG; q++)
        if (q < D) dot = fma(xi[q], sXc[q * TILE + jl], dot);
      double r2 = si + sSc[jl] - 2.0 * dot;
      if (gi == gj) r2 = 0.0;
      r2 = fmax(r2, 0.0);
      double aa = 0.0;
#pragma unroll
      for (int q = 0; q < MAX_P; q++)
        if (q < P) aa = fma(ai[q], sAc[q * TILE + jl], aa);
      const double dl = 0.5 * (aa - (double)P * kinv);
      if (gi == gj) {
        gnoise += dl;
        if (p.dnoise_out) p.dnoise_out[gi] = dl;
      }
      if (kind == GPX_RBF) {
        // dK/dr / r = -K for the RBF kernel (rbf.py:177-178 over stationary.py:205): neither the square root nor the division
        // of the general form is needed; at r = 0 the reference's 1/0 := 0 multiplies (x_i - x_j)^2 = 0 either way
        const double r2s = r2 * (inv_ls * inv_ls);
        const double k = exp(-0.5 * r2s);
        const double kd = w * k * dl;
        gvar += kd;
        const double tmpv = -variance * kd;
        if (ard) {
#pragma unroll
          for (int q = 0; q < DREG; q++)
            if (q < D) {
              const double df = xi[q] - sXc[q * TILE + jl];
              gq[q] = fma(tmpv, df * df, gq[q]);
            }
        } else {
          giso = fma(tmpv, r2s, giso);
        }
        continue;
      }
      const double rr = sqrt(r2) * inv_ls;
      double k, dk;
      k_dk_of_r_unit(kind, rr, k, dk);
      gvar = fma(w * k, dl, gvar);
      const double G = variance * dk * dl;
      if (ard) {
        const double tmpv = (rr != 0.0) ? w * G / rr : 0.0;     // stationary.py:205,225-232: 1/r with 1/0 := 0
#pragma unroll
        for (int q = 0; q < DREG; q++)
          if (q < D) {
            const double df = xi[q] - sXc[q * TILE + jl];
            gq[q] = fma(tmpv, df * df, gq[q]);
          }
      } else {
        giso = fma(w * G, rr, giso);
      }
    }
  }
  gvar = warp_sum(gvar);
  gnoise = warp_sum(gnoise);
  if (lane == 0) { sRed[warp * nred] = gvar; sRed[warp * nred + nred - 1] = gnoise; }
  if (ard) {
#pragma unroll
    for (int q = 0; q < DREG; q++)
      if (q < D) {
        const double s = warp_sum(gq[q]);
        if (lane == 0) sRed[warp * nred + 1 + q] = s;
      }
  } else {
    giso = warp_sum(giso);
    if (lane == 0) sRed[warp * nred + 1] = giso;
  }
  __syncthreads();
  if (tid < nred) {
    double s = 0.0;
#pragma unroll
    for (int wdx = 0; wdx < 8; wdx++) s += sRed[wdx * nred + tid];
    p.partials[(((long)r * p.nt + c) * csplit + sidx) * nred + tid] = s;
  }
}

int grad_kinv_csplit(int nt, int nred) {
  const int tiles = nt * (nt + 1) / 2;
  int cs = 1;
  while (cs < 8 && tiles * cs < 592 && (cs * 2) * nred <= MAX_D + 2) cs *= 2;   // ~4 waves of CTAs; partials sized for MAX_D + 2 per tile
  return cs;
}

template <int DREG>
static int launch_grad_kinv_t(const GradKinvParams& p, unsigned grid, size_t smem, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    GPX_CUDA(cudaFuncSetAttribute(grad_kinv_kernel<DREG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)((MAX_D + 1 + MAX_P) * TILE * 8 + 8 * (MAX_D + 2) * 8)));
    attr_set = true;
  }
  grad_kinv_kernel<DREG><<<grid, 256, smem, st>>>(p);
  GPX_CUDA(cudaGetLastError());
  return 0;
}

int launch_grad_kinv(const GradKinvParams& p, cudaStream_t st) {
  const int D = p.kp.D;
  const size_t smem = (size_t)(D + 1 + p.P) * TILE * 8 + 8 * (MAX_D + 2) * 8;
  if (p.csplit != 1 && p.csplit != 2 && p.csplit != 4 && p.csplit != 8) { set_error("grad_kinv: csplit must be 1, 2, 4 or 8"); return -2; }
  const unsigned grid = (unsigned)(p.nt * (p.nt + 1) / 2 * p.csplit);
  if (D <= 8) return launch_grad_kinv_t<8>(p, grid, smem, st);
  if (D <= 16) return launch_grad_kinv_t<16>(p, grid, smem, st);
  if (D <= 32) return launch_grad_kinv_t<32>(p, grid, smem, st);
  return launch_grad_kinv_t<64>(p, grid, smem, st);
}

}  // namespace gpx
