// gpx_sparse.cu — the N-dependent work of sparse GP regression (VarDTC) on the device.
//
// Reference: GPy/inference/latent_function_inference/var_dtc.py:66-215 with the gradient wiring of
// GPy/core/sparse_gp.py:108-119 (Gaussian likelihood, homoscedastic noise, certain inputs). Everything that scales with
// the number of data points N stays in HBM and never crosses PCIe:
//   gpx_sparse_stats : psi1 = K(X, Z) in both layouts (8 N M bytes each), G = psi1^T psi1 (M x M, DMMA GEMM with
//                      k-depth N) and psi1^T Y (M x P)            -> var_dtc.py:126-132,139-141 (A, psi1Vf need only these)
//   gpx_sparse_grads : dL_dKnm^T = W2 psi1^T + C (beta Y)^T (M x N, never materialised on the host) reduced straight to
//                      d/d(variance, lengthscale) and dL/dZ       -> var_dtc.py:219-234, sparse_gp.py:112,118
// The M x M algebra in between (two Choleskys, back-substitutions; var_dtc.py:135-156) is driven from the host mirror
// (gpy_b200/sparse.py) through gpx_pdinv and small host products: it does not depend on N.
#include <algorithm>
#include <cstring>
#include <vector>

#include "gpx_common.cuh"
#include "gpx_ctx.cuh"
#include "gpx_kernels.cuh"

using namespace gpx;

#define GPX_CHECK(x)            \
  do {                          \
    int rc__ = (x);             \
    if (rc__ != 0) return rc__; \
  } while (0)
#define GPX_FAIL(msg)       \
  do {                      \
    gpx::set_error(msg);    \
    return -2;              \
  } while (0)

struct SparseState {
  long N = 0, Npad = 0, M = 0, Mpad = 0;
  int D = 0, P = 0;
  double *X = nullptr, *XsT = nullptr, *sqX = nullptr, *Y = nullptr, *Yb = nullptr;   // Y: [P][Npad]
  double *Z = nullptr, *ZsT = nullptr, *sqZ = nullptr;
  double *Kuf = nullptr;   // Mpad x Npad column-major: (a, n) at a + n*Mpad
  double *Kfu = nullptr;   // Npad x Mpad column-major: (n, a) at n + a*Npad
  double *dLt = nullptr;   // Mpad x Npad: dL_dKnm^T
  double *Gm = nullptr, *W2 = nullptr, *Cm = nullptr;   // Mpad x Mpad, Mpad x Mpad, [P][Mpad]
  double *part = nullptr; size_t part_cap = 0;
  KernParams kp{};
  bool have_stats = false;
};

namespace {
void free_m(SparseState* s) {
  double** ptrs[] = {&s->Z, &s->ZsT, &s->sqZ, &s->Kuf, &s->Kfu, &s->dLt, &s->Gm, &s->W2, &s->Cm};
  for (auto p : ptrs) { if (*p) cudaFree(*p); *p = nullptr; }
  s->M = s->Mpad = 0;
  s->have_stats = false;
}
void free_all(SparseState* s) {
  free_m(s);
  double** ptrs[] = {&s->X, &s->XsT, &s->sqX, &s->Y, &s->Yb, &s->part};
  for (auto p : ptrs) { if (*p) cudaFree(*p); *p = nullptr; }
  s->part_cap = 0;
  s->N = s->Npad = 0;
}
int ensure_part(SparseState* s, size_t bytes) {
  if (s->part_cap >= bytes) return 0;
  if (s->part) cudaFree(s->part);
  s->part = nullptr;
  GPX_CUDA(cudaMalloc(&s->part, bytes));
  s->part_cap = bytes;
  return 0;
}
__global__ void scale_kernel(const double* __restrict__ in, double a, long n, double* __restrict__ out) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a * in[i];
}
}  // namespace

namespace gpx {
void sparse_free(gpx_ctx* c) {
  if (!c->sparse) return;
  free_all(c->sparse);
  delete c->sparse;
  c->sparse = nullptr;
}
}  // namespace gpx

extern "C" {

int gpx_sparse_set_data(gpx_ctx* c, const double* X, int64_t N, int D, const double* Y, int P) {
  if (!c || !X || !Y) GPX_FAIL("null argument");
  if (N < 1 || D < 1 || D > MAX_D || P < 1 || P > MAX_P) GPX_FAIL("bad shape (D <= 64, P <= 8)");
  GPX_CUDA(cudaSetDevice(c->device));
  if (!c->sparse) c->sparse = new SparseState();
  SparseState* s = c->sparse;
  const long Npad = (N + TILE - 1) / TILE * TILE;
  if (Npad != s->Npad || D != s->D || P != s->P) {
    GPX_CUDA(cudaStreamSynchronize(c->st));
    free_all(s);
    s->Npad = Npad; s->D = D; s->P = P;
    GPX_CUDA(cudaMalloc(&s->X, (size_t)Npad * D * 8));
    GPX_CUDA(cudaMalloc(&s->XsT, (size_t)Npad * D * 8));
    GPX_CUDA(cudaMalloc(&s->sqX, (size_t)Npad * 8));
    GPX_CUDA(cudaMalloc(&s->Y, (size_t)Npad * P * 8));
    GPX_CUDA(cudaMalloc(&s->Yb, (size_t)Npad * P * 8));
  }
  s->N = N;
  s->have_stats = false;
  GPX_CUDA(cudaMemcpyAsync(s->X, X, (size_t)N * D * 8, cudaMemcpyHostToDevice, c->st));
  GPX_CUDA(cudaMemcpyAsync(s->Yb, Y, (size_t)N * P * 8, cudaMemcpyHostToDevice, c->st));
  GPX_CHECK(launch_transpose_pad(s->Yb, N, P, Npad, s->Y, c->st));
  c->total_launches++;
  GPX_CUDA(cudaStreamSynchronize(c->st));
  return 0;
}

int gpx_sparse_stats(gpx_ctx* c, int kind, int ard, double variance, const double* lengthscale, const double* Z,
                     int64_t M, double* G, double* psi1tY) {
  if (!c || !c->sparse || !c->sparse->X) GPX_FAIL("gpx_sparse_set_data has not been called");
  if (!Z || !G || !psi1tY || !lengthscale) GPX_FAIL("null argument");
  if (M < 1) GPX_FAIL("M must be positive");
  GPX_CUDA(cudaSetDevice(c->device));
  SparseState* s = c->sparse;
  cudaStream_t st = c->st;
  GPX_CHECK(fill_kp(s->kp, kind, ard, s->D, variance, lengthscale));
  const long Mpad = (M + TILE - 1) / TILE * TILE, Npad = s->Npad, N = s->N;
  if (Mpad != s->Mpad) {
    GPX_CUDA(cudaStreamSynchronize(st));
    free_m(s);
    s->Mpad = Mpad;
    GPX_CUDA(cudaMalloc(&s->Z, (size_t)Mpad * s->D * 8));
    GPX_CUDA(cudaMalloc(&s->ZsT, (size_t)Mpad * s->D * 8));
    GPX_CUDA(cudaMalloc(&s->sqZ, (size_t)Mpad * 8));
    GPX_CUDA(cudaMalloc(&s->Kuf, (size_t)Mpad * Npad * 8));
    GPX_CUDA(cudaMalloc(&s->Kfu, (size_t)Mpad * Npad * 8));
    GPX_CUDA(cudaMalloc(&s->dLt, (size_t)Mpad * Npad * 8));
    GPX_CUDA(cudaMalloc(&s->Gm, (size_t)Mpad * Mpad * 8));
    GPX_CUDA(cudaMalloc(&s->W2, (size_t)Mpad * Mpad * 8));
    GPX_CUDA(cudaMalloc(&s->Cm, (size_t)Mpad * s->P * 8));
    GPX_CUDA(cudaMemsetAsync(s->Kuf, 0, (size_t)Mpad * Npad * 8, st));   // padding stays zero: only valid entries
    GPX_CUDA(cudaMemsetAsync(s->Kfu, 0, (size_t)Mpad * Npad * 8, st));   // are ever written
  }
  s->M = M;
  GPX_CUDA(cudaMemcpyAsync(s->Z, Z, (size_t)M * s->D * 8, cudaMemcpyHostToDevice, st));
  GPX_CHECK(launch_prep_x(s->X, N, Npad, s->kp, s->XsT, s->sqX, st));
  GPX_CHECK(launch_prep_x(s->Z, M, Mpad, s->kp, s->ZsT, s->sqZ, st));
  const int mt = (int)(Mpad / TILE), ntl = (int)(Npad / TILE);
  KBuildParams kb;
  memset(&kb, 0, sizeof(kb));
  kb.kp = s->kp; kb.sym = 0; kb.same = 0;
  // psi1^T: thread-mapped operand = inducing points
  kb.rowsT = s->ZsT; kb.ld_rows = Mpad; kb.sq_rows = s->sqZ; kb.colsT = s->XsT; kb.ld_cols = Npad; kb.sq_cols = s->sqX;
  kb.out = s->Kuf; kb.ld = Mpad; kb.nrows = M; kb.ncols = N;
  GPX_CHECK(launch_kbuild(kb, mt, ntl, st));
  // psi1: thread-mapped operand = data points
  kb.rowsT = s->XsT; kb.ld_rows = Npad; kb.sq_rows = s->sqX; kb.colsT = s->ZsT; kb.ld_cols = Mpad; kb.sq_cols = s->sqZ;
  kb.out = s->Kfu; kb.ld = Npad; kb.nrows = N; kb.ncols = M;
  GPX_CHECK(launch_kbuild(kb, ntl, mt, st));
  // G = psi1^T psi1 (lower tiles), k-depth = Npad
  {
    GemmParams pg = gemm_defaults();
    pg.mode = GEMM_PANEL; pg.plain = 2;
    pg.A = s->Kuf; pg.lda = Mpad; pg.B = s->Kuf; pg.ldb = Mpad; pg.C = s->Gm; pg.ldc = Mpad;
    pg.K = (int)Npad; pg.nt = mt; pg.ncols = mt;
    GPX_CHECK(launch_gemm(pg, dim3(1, 1), st));
  }
  GPX_CHECK(launch_col_dot(s->Kfu, Npad, N, M, s->P, s->Y, Npad, s->Cm, Mpad, st));
  c->total_launches += 6;
  std::vector<double> hG((size_t)Mpad * Mpad), hC((size_t)Mpad * s->P);
  GPX_CUDA(cudaMemcpyAsync(hG.data(), s->Gm, hG.size() * 8, cudaMemcpyDeviceToHost, st));
  GPX_CUDA(cudaMemcpyAsync(hC.data(), s->Cm, hC.size() * 8, cudaMemcpyDeviceToHost, st));
  GPX_CUDA(cudaStreamSynchronize(st));
  for (long j = 0; j < M; j++)
    for (long i = 0; i < M; i++) {
      const long ti = i / TILE, tj = j / TILE;
      G[i + j * M] = (ti >= tj) ? hG[i + j * Mpad] : hG[j + i * Mpad];   // lower tiles were computed: mirror
    }
  for (long i = 0; i < M; i++)
    for (int q = 0; q < s->P; q++) psi1tY[i * s->P + q] = hC[(size_t)q * Mpad + i];
  s->have_stats = true;
  return 0;
}

int gpx_sparse_grads(gpx_ctx* c, const double* W2, const double* Cmat, double beta, double* dvariance,
                     double* dlengthscale, double* dZ) {
  if (!c || !c->sparse || !c->sparse->have_stats) GPX_FAIL("gpx_sparse_stats has not been called");
  if (!W2 || !Cmat || !dvariance || !dlengthscale || !dZ) GPX_FAIL("null argument");
  GPX_CUDA(cudaSetDevice(c->device));
  SparseState* s = c->sparse;
  cudaStream_t st = c->st;
  const long M = s->M, Mpad = s->Mpad, N = s->N, Npad = s->Npad;
  const int D = s->D, P = s->P;
  const int mt = (int)(Mpad / TILE), ntl = (int)(Npad / TILE);
  GPX_CUDA(cudaMemsetAsync(s->W2, 0, (size_t)Mpad * Mpad * 8, st));
  GPX_CUDA(cudaMemcpy2DAsync(s->W2, Mpad * 8, W2, M * 8, (size_t)M * 8, M, cudaMemcpyHostToDevice, st));
  std::vector<double> hC((size_t)Mpad * P, 0.0);
  for (long i = 0; i < M; i++)
    for (int q = 0; q < P; q++) hC[(size_t)q * Mpad + i] = Cmat[i * P + q];
  GPX_CUDA(cudaMemcpyAsync(s->Cm, hC.data(), hC.size() * 8, cudaMemcpyHostToDevice, st));
  scale_kernel<<<(unsigned)((Npad * P + 255) / 256), 256, 0, st>>>(s->Y, beta, Npad * P, s->Yb);
  GPX_CUDA(cudaGetLastError());
  // dLt = W2 * psi1^T  (M x N), W2 symmetric
  {
    GemmParams pg = gemm_defaults();
    pg.mode = GEMM_PANEL; pg.plain = 1;
    pg.A = s->W2; pg.lda = Mpad; pg.B = s->Kfu; pg.ldb = Npad; pg.C = s->dLt; pg.ldc = Mpad;
    pg.K = (int)Mpad; pg.nt = mt; pg.ncols = ntl;
    GPX_CHECK(launch_gemm(pg, dim3(1, 1), st));
  }
  // kernel-parameter gradients of sum(dL_dKnm * K(X, Z)): rows i = data points, columns j = inducing points
  const int nl = s->kp.ard ? D : 1, nred = nl + 1;
  const int nchunk = (int)std::max<long>(1, std::min<long>((N + 31) / 32, (4 * 148 + mt - 1) / mt));
  const long mchunk = ((N + nchunk - 1) / nchunk + 31) / 32 * 32;
  const int nch = (int)((N + mchunk - 1) / mchunk);
  const size_t part_full = (size_t)mt * ntl * nred * 8, part_x = (size_t)nch * M * D * 8 + (size_t)M * D * 8;
  GPX_CHECK(ensure_part(s, std::max(part_full, part_x)));
  GradFullParams gp;
  memset(&gp, 0, sizeof(gp));
  gp.x1T = s->XsT; gp.ld1 = Npad; gp.sq1 = s->sqX; gp.N = N;
  gp.x2T = s->ZsT; gp.ld2 = Mpad; gp.sq2 = s->sqZ; gp.M = M;
  gp.dL_dK = s->dLt; gp.ldd = Mpad; gp.same = 0; gp.partials = s->part; gp.kp = s->kp;
  gp.ci = s->Yb; gp.ldci = Npad; gp.cj = s->Cm; gp.ldcj = Mpad; gp.cP = P;
  GPX_CHECK(launch_grad_full(gp, mt, ntl, st));
  std::vector<double> hp((size_t)mt * ntl * nred);
  GPX_CUDA(cudaMemcpyAsync(hp.data(), s->part, hp.size() * 8, cudaMemcpyDeviceToHost, st));
  GPX_CUDA(cudaStreamSynchronize(st));
  {
    std::vector<double> tot(nred, 0.0);
    for (size_t t = 0; t < (size_t)mt * ntl; t++)
      for (int q = 0; q < nred; q++) tot[q] += hp[t * nred + q];
    *dvariance = tot[0];
    for (int q = 0; q < nl; q++) dlengthscale[q] = -tot[1 + q] / s->kp.ls[q];
  }
  // dL/dZ: gradients_X(dL_dKnm^T, Z, X): rows = inducing points, columns = data points, dL read transposed
  GradFullParams gx;
  memset(&gx, 0, sizeof(gx));
  gx.x1T = s->ZsT; gx.ld1 = Mpad; gx.sq1 = s->sqZ; gx.N = M;
  gx.x2T = s->XsT; gx.ld2 = Npad; gx.sq2 = s->sqX; gx.M = N;
  gx.dL_dK = s->dLt; gx.ldd = Mpad; gx.transposed = 1; gx.same = 0; gx.kp = s->kp;
  gx.ci = s->Cm; gx.ldci = Mpad; gx.cj = s->Yb; gx.ldcj = Npad; gx.cP = P;
  double* dout = s->part + (size_t)nch * M * D;
  GPX_CHECK(launch_gradx(gx, nch, mchunk, s->part, dout, st));
  GPX_CUDA(cudaMemcpyAsync(dZ, dout, (size_t)M * D * 8, cudaMemcpyDeviceToHost, st));
  GPX_CUDA(cudaStreamSynchronize(st));
  c->total_launches += 5;
  return 0;
}

}  // extern "C"
