/* cg_port.c -- CPU port (C + OpenMP) of the restated oracle's two scalable loops:
 * CSR SpMV and (Jacobi-)preconditioned conjugate gradients on the periodic-merged
 * pressure operator (oracle/restated.py: PressureSystem.solve_cg, itself the
 * restatement of np.linalg.solve(A_pressure, b_p), code/StokesColor.py:554-555).
 *
 * TEST INFRASTRUCTURE ONLY: built into oracle/_ref/libcgport.so by oracle/Makefile
 * and used (a) by tests/ to cross-check the port against scipy and (b) by bench.py
 * as the multi-threaded CPU baseline ("kind": "port").  Never loaded by the product.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int cgport_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void cgport_spmv(int64_t n, const int32_t* rowptr, const int32_t* colidx, const double* vals,
                 const double* x, double* y) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    double s = 0.0;
    for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) s += vals[k] * x[colidx[k]];
    y[i] = s;
  }
}

static double dot(int64_t n, const double* a, const double* b) {
  double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

/* Same recurrence as libfluidsim's CG and scipy.sparse.linalg.cg.  Stops when
 * ||r|| <= rtol*||b|| or after maxit iterations; returns the iterations done.
 * project_mean: b is made mean-free first, x's mean removed at the end. */
int cgport_cg(int64_t n, const int32_t* rowptr, const int32_t* colidx, const double* vals,
              const double* b_in, double* x, double rtol, int maxit, int jacobi, int project_mean,
              double* relres) {
  double* b = (double*)malloc(sizeof(double) * n);
  double* r = (double*)malloc(sizeof(double) * n);
  double* p = (double*)malloc(sizeof(double) * n);
  double* Ap = (double*)malloc(sizeof(double) * n);
  double* dinv = (double*)malloc(sizeof(double) * n);
  double mean = 0.0;
  if (project_mean) {
    for (int64_t i = 0; i < n; ++i) mean += b_in[i];
    mean /= (double)n;
  }
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    b[i] = b_in[i] - mean;
    double d = 1.0;
    if (jacobi)
      for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k)
        if (colidx[k] == i && vals[k] != 0.0) d = 1.0 / vals[k];
    dinv[i] = d;
  }
  cgport_spmv(n, rowptr, colidx, vals, x, Ap);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    r[i] = b[i] - Ap[i];
    p[i] = dinv[i] * r[i];
  }
  double bb = dot(n, b, b), rr = dot(n, r, r), rz = dot(n, r, p);
  int it = 0;
  const double tol2 = rtol * rtol;
  while (bb > 0.0 && rr > tol2 * bb && it < maxit) {
    cgport_spmv(n, rowptr, colidx, vals, p, Ap);
    double pAp = dot(n, p, Ap);
    double alpha = pAp != 0.0 ? rz / pAp : 0.0;
    double rr_new = 0.0, rz_new = 0.0;
#pragma omp parallel for reduction(+ : rr_new, rz_new) schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      x[i] += alpha * p[i];
      r[i] -= alpha * Ap[i];
      rr_new += r[i] * r[i];
      rz_new += r[i] * (dinv[i] * r[i]);
    }
    double beta = rz != 0.0 ? rz_new / rz : 0.0;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) p[i] = dinv[i] * r[i] + beta * p[i];
    rr = rr_new;
    rz = rz_new;
    ++it;
  }
  if (project_mean) {
    double m = 0.0;
    for (int64_t i = 0; i < n; ++i) m += x[i];
    m /= (double)n;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) x[i] -= m;
  }
  if (relres) *relres = bb > 0.0 ? sqrt(rr / bb) : 0.0;
  free(b); free(r); free(p); free(Ap); free(dinv);
  return it;
}
