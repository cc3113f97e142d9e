/* oracle_c.c — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * Plain-C restatement of the one native loop on the reference's hot path:
 *   GPy/kern/src/stationary_cython.pyx:53-62  lengthscale_grads(N, M, Q, tmp, X, X2, grad)
 * which is what Stationary._lengthscale_grads_cython (GPy/kern/src/stationary.py:237-243) calls. The
 * reference loop is SERIAL (q outermost, then n, then m, `with nogil`, no prange) and accumulates into grad[q];
 * this restatement keeps that order so that timing it reproduces the reference's single-thread ARD pass.
 */
void oracle_lengthscale_grads(int N, int M, int Q, const double* tmp, const double* X, const double* X2, double* grad) {
  for (int q = 0; q < Q; q++) {
    double g = 0.0;
    for (int n = 0; n < N; n++) {
      const double xn = X[(long)n * Q + q];
      const double* t = tmp + (long)n * M;
      for (int m = 0; m < M; m++) {
        const double dist = xn - X2[(long)m * Q + q];
        g += t[m] * dist * dist;
      }
    }
    grad[q] = g;
  }
}

/* GPy/util/linalg_cython.pyx:9-18 symmetrify: copy lower -> upper (or upper -> lower), in place, N x N C-order. */
void oracle_symmetrify(int N, double* A, int upper) {
  if (upper) {
    for (int i = 0; i < N; i++)
      for (int j = 0; j < i; j++) A[(long)i * N + j] = A[(long)j * N + i];
  } else {
    for (int i = 0; i < N; i++)
      for (int j = 0; j < i; j++) A[(long)j * N + i] = A[(long)i * N + j];
  }
}
