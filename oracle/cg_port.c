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

/* ---------------------------------------------------------------------------------------------
 * Multigrid-preconditioned CG: the CPU statement of the algorithm libfluidsim runs by default on
 * large pressure systems (CG + one V(1,1) damped-Jacobi cycle of a smoothed-aggregation hierarchy
 * per iteration).  The hierarchy (A_l, P_l, R_l = P_l^T as CSR, 1/diag(A_l), dense (pseudo-)inverse
 * of the coarsest operator) is built by oracle/amg_cpu.py with scipy; this file only applies it.
 * Used by bench.py --impl reference / cpu_baseline as the same-algorithm CPU arm.
 */
typedef struct cgport_level {
  int64_t n;
  const int32_t *a_rp, *a_ci; const double* a_v;     /* A_l            (n x n)        */
  const int32_t *p_rp, *p_ci; const double* p_v;     /* P_l            (n x n_next)   */
  const int32_t *r_rp, *r_ci; const double* r_v;     /* R_l = P_l^T    (n_next x n)   */
  const double* dinv;
  double *x, *b, *t;                                 /* work vectors of this level (n) */
} cgport_level;

static void vcycle(const cgport_level* L, int nlev, int l, const double* cinv, double omega, const double* b, double* x) {
  const int64_t n = L[l].n;
  if (l == nlev - 1) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      const double* row = cinv + i * n;
      double s = 0.0;
      for (int64_t j = 0; j < n; ++j) s += row[j] * b[j];
      x[i] = s;
    }
    return;
  }
  const cgport_level* lv = &L[l];
  const cgport_level* nx = &L[l + 1];
  double* t = lv->t;
  /* pre-smooth from zero, residual */
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) x[i] = omega * lv->dinv[i] * b[i];
  cgport_spmv(n, lv->a_rp, lv->a_ci, lv->a_v, x, t);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) t[i] = b[i] - t[i];
  cgport_spmv(nx->n, lv->r_rp, lv->r_ci, lv->r_v, t, nx->b);
  vcycle(L, nlev, l + 1, cinv, omega, nx->b, nx->x);
  cgport_spmv(n, lv->p_rp, lv->p_ci, lv->p_v, nx->x, t);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) x[i] += t[i];
  /* post-smooth */
  cgport_spmv(n, lv->a_rp, lv->a_ci, lv->a_v, x, t);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) x[i] += omega * lv->dinv[i] * (b[i] - t[i]);
}

/* one application of the preconditioner z = M^-1 r (exposed for the tests) */
void cgport_vcycle(const cgport_level* L, int nlev, const double* cinv, double omega, const double* r, double* z) {
  vcycle(L, nlev, 0, cinv, omega, r, z);
}

int cgport_pcg_amg(const cgport_level* L, int nlev, const double* cinv, double omega, const double* b_in, double* x,
                   double rtol, int maxit, int project_mean, double* relres) {
  const int64_t n = L[0].n;
  double* b = (double*)malloc(sizeof(double) * n);
  double* r = (double*)malloc(sizeof(double) * n);
  double* p = (double*)malloc(sizeof(double) * n);
  double* Ap = (double*)malloc(sizeof(double) * n);
  double* z = (double*)malloc(sizeof(double) * n);
  double mean = 0.0;
  if (project_mean) {
#pragma omp parallel for reduction(+ : mean) schedule(static)
    for (int64_t i = 0; i < n; ++i) mean += b_in[i];
    mean /= (double)n;
  }
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) b[i] = b_in[i] - mean;
  cgport_spmv(n, L[0].a_rp, L[0].a_ci, L[0].a_v, x, Ap);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) r[i] = b[i] - Ap[i];
  double bb = dot(n, b, b), rr = dot(n, r, r);
  const double tol2 = rtol * rtol;
  int it = 0;
  if (bb > 0.0 && rr > tol2 * bb) {
    vcycle(L, nlev, 0, cinv, omega, r, z);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) p[i] = z[i];
    double rz = dot(n, r, z);
    while (it < maxit) {
      cgport_spmv(n, L[0].a_rp, L[0].a_ci, L[0].a_v, p, Ap);
      const double pAp = dot(n, p, Ap);
      const double alpha = pAp != 0.0 ? rz / pAp : 0.0;
      double rr_new = 0.0;
#pragma omp parallel for reduction(+ : rr_new) schedule(static)
      for (int64_t i = 0; i < n; ++i) {
        x[i] += alpha * p[i];
        r[i] -= alpha * Ap[i];
        rr_new += r[i] * r[i];
      }
      rr = rr_new;
      ++it;
      if (rr <= tol2 * bb) break;
      vcycle(L, nlev, 0, cinv, omega, r, z);
      const double rz_new = dot(n, r, z);
      const double beta = rz != 0.0 ? rz_new / rz : 0.0;
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
      rz = rz_new;
    }
  }
  if (project_mean) {
    double m = 0.0;
#pragma omp parallel for reduction(+ : m) schedule(static)
    for (int64_t i = 0; i < n; ++i) m += x[i];
    m /= (double)n;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) x[i] -= m;
  }
  if (relres) *relres = bb > 0.0 ? sqrt(rr / bb) : 0.0;
  free(b); free(r); free(p); free(Ap); free(z);
  return it;
}

void cgport_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ---- vector kernels of the solution-subspace projection (oracle/cpu_step.py: Recycler; GPU: csrc/recycle.cu).
 * X: k vectors of length n, row-major with leading dimension ld. */
void cgport_multidot(int64_t n, int k, const double* X, int64_t ld, const double* a, double* out) {
  for (int j = 0; j < k; ++j) {
    const double* x = X + (int64_t)j * ld;
    double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (int64_t i = 0; i < n; ++i) s += x[i] * a[i];
    out[j] = s;
  }
}

/* out = beta * in + sum_j coef[j] X_j  (in may be NULL; out may alias in) */
void cgport_comb(int64_t n, int k, const double* X, int64_t ld, const double* coef, double beta, const double* in, double* out) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    double v = in ? beta * in[i] : 0.0;
    for (int j = 0; j < k; ++j) v += coef[j] * X[(int64_t)j * ld + i];
    out[i] = v;
  }
}
