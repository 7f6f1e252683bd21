// stokes.cu -- the operator-split Stokes step of code/StokesColor.py:537-575
// (== code/StokesFood.py:441-479), device resident.
//
//   u*  = solve(A_visc, u)            two RHS in one CG (A_visc = I + DT*nu*K, :471-475)
//   makePerBCU(u*); makeDirBCU(u*)    :546-547
//   p   = solve(A_pressure, -div(u*)/DT)          :551-555   (restated, see below)
//   u   = u* - DT*grad(p); BCs        :559-564
//   p2  = solve(A_pressure, -div(u)/DT)           :567-569
//   u[interior] -= DT*grad(p2)[interior]          :570-573
//
// Pressure restatement (DESIGN.md "pressure system"): the reference's
// A_pressure = diag(1/(M+1e-12)) K + 1e10 penalty on periodic pairs is singular
// and non-symmetric.  Here each periodic pair is merged into one dof (Z), the
// equations are multiplied by the lumped mass and the SPD system
//     (Z^T K Z) q = Z^T (M * b) - mean,   p = Z (q - mean q)
// is solved by CG.
#include <chrono>

#include "internal.cuh"
#include "recycle.cuh"

struct fs_stokes {
  fs_mesh* mesh = nullptr;
  double DT = 0, nu = 0;
  fs_csr a_visc;          // pattern borrowed from the mesh
  fs_csr k_red;           // periodic-merged stiffness, own pattern
  fs::Pattern pat_red;
  fs::DBuf<int> dof;      // (N) node -> merged dof
  int64_t nd = 0;
  fs::DBuf<int> rep;      // (nd) dof -> its representative (smallest) node
  fs::DBuf<int> ex_ptr, ex_dof, ex_node;   // dofs with merged-in nodes: CSR list of those nodes, ascending
  int64_t n_ex = 0;
  fs::DBuf<unsigned char> is_dir, is_interior;
  fs::DBuf<double> ustar, div, rhs_red, p_red, p2_red, p_full, p2_full;
  bool have_p = false;
  // history of one pressure solve: its solutions one and two steps back (q1, q2, nq of them valid) and
  // their images y = K q (ny valid; recomputed on demand, never part of a checkpoint)
  struct PressHist {
    fs::DBuf<double> q1, q2, y1, y2;
    int nq = 0, ny = 0;
  };
  PressHist h1, h2;                 // first / second projection of the step
  fs::DBuf<double> y0, q_try;       // K q of the current solve; the chosen extrapolated guess
  fs::Recycler rec1, rec2;          // solution-subspace projection of the two solves (large systems, recycle.cuh)
};

namespace fs {

__global__ void k_flag(const int* __restrict__ idx, int64_t n, unsigned char* __restrict__ flag) {
  int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k < n) flag[idx[k]] = 1;
}

// A_visc values on K's pattern: I + (DT*nu) K, Dirichlet rows+columns zeroed, diagonal 1
__global__ void k_visc_vals(CsrView K, double dtnu, const unsigned char* __restrict__ is_dir, double* __restrict__ out) {
  int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (row >= K.n) return;
  const bool rd = is_dir[row];
  for (int k = K.rowptr[row]; k < K.rowptr[row + 1]; ++k) {
    const int col = K.colidx[k];
    const bool diag = (col == row);
    double v = (diag ? 1.0 : 0.0) + dtnu * K.vals[k];
    if (rd || is_dir[col]) v = diag ? 1.0 : 0.0;
    out[k] = v;
  }
}

// rhs_red[d] = sum over the nodes n merged into dof d of M[n] * (-(1/DT) * div[n])
// (code/StokesColor.py:554 then mass-weighting).  No atomics: the dof's representative (smallest node
// id) is gathered first, the few merged-in nodes are added in ascending node order by one thread per
// dof that has any -- the same bits on every run however many nodes share a dof.
__global__ void k_pressure_rhs(int64_t nd, const int* __restrict__ rep, const double* __restrict__ mass,
                               const double* __restrict__ div, double s, double* __restrict__ rhs_red, int64_t bs_n = 0) {
  div += blockIdx.y * bs_n; rhs_red += blockIdx.y * nd;
  int64_t d = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (d >= nd) return;
  const int n = rep[d];
  rhs_red[d] = mass[n] * (s * div[n]);
}
__global__ void k_pressure_rhs_extra(int64_t n_ex, const int* __restrict__ ex_ptr, const int* __restrict__ ex_dof,
                                     const int* __restrict__ ex_node, const double* __restrict__ mass,
                                     const double* __restrict__ div, double s, double* __restrict__ rhs_red, int64_t bs_n = 0,
                                     int64_t bs_nd = 0) {
  div += blockIdx.y * bs_n; rhs_red += blockIdx.y * bs_nd;
  int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= n_ex) return;
  double v = rhs_red[ex_dof[j]];
  for (int k = ex_ptr[j]; k < ex_ptr[j + 1]; ++k) { const int n = ex_node[k]; v += mass[n] * (s * div[n]); }
  rhs_red[ex_dof[j]] = v;
}

__global__ void k_expand(int64_t N, const int* __restrict__ dof, const double* __restrict__ q, double* __restrict__ p,
                         int64_t bs_nd = 0) {
  q += blockIdx.y * bs_nd; p += blockIdx.y * N;            // blockIdx.y = configuration of a batched sweep
  int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (n < N) p[n] = q[dof[n]];
}

// Warm start: the previous step's solution q, or its extrapolation in time from the last two / three
// steps (2 q - q1, 3 q - 3 q1 + q2), whichever has the smallest residual |b - K q|.  K is linear, so
// with y_k = K q_k kept from the earlier steps the three residual norms cost ONE SpMV (y0 = K q) and
// one pass over four vectors: it buys 7-10 of ~27 PCG iterations while the flow evolves smoothly and
// loses nothing when it does not (every solve still runs to rtol).
// Systems of at least kRecycleMinRows rows (where AUTO also picks the multigrid preconditioner) start from the
// projection of the new solution onto the span of the previous ones instead (recycle.cuh): 542 instead of 921 PCG
// iterations over the bench's 20 timed steps.
constexpr int64_t kRecycleMinRows = 20000;

static void pressure_solve(fs_stokes* s, const double* d_vel, double* q, fs_stokes::PressHist& H, Recycler& rec, double* p_full,
                           const fs_stokes_opts& o, int* iters, double* relres, double* max_div) {
  static const int extrap = [] { const char* e = std::getenv("FS_STOKES_EXTRAP"); return e ? std::atoi(e) : 2; }();
  const bool recycle = o.warm_start && s->nd >= kRecycleMinRows && recycle_enabled();
  const Recycler::MatVec matvec = [s](const double* x, double* y) { spmv_best_dev(&s->k_red, x, y); };
  const Recycler::Reduce reduce = [](const double* part, int nblk, int nchunk, double* host) {
    std::vector<double> h((size_t)nblk * nchunk * 6);
    FS_CUDA(cudaMemcpyAsync(h.data(), part, h.size() * sizeof(double), cudaMemcpyDeviceToHost, stream()));
    FS_CUDA(cudaStreamSynchronize(stream()));
    for (int c = 0; c < nchunk; ++c)
      for (int j = 0; j < 6; ++j) {
        double t = 0.0;
        for (int b = 0; b < nblk; ++b) t += h[((size_t)c * nblk + b) * 6 + j];
        host[6 * c + j] = t;
      }
  };
  if (recycle && !rec.ready()) rec.init(s->nd, recycle_kmax(), recycle_keep());
  fs_mesh* m = s->mesh;
  cudaStream_t st = stream();
  divergence_dev(m, d_vel, s->div.p, nullptr);
  if (max_div) *max_div = max_abs_dev(s->div.p, m->N);
  k_pressure_rhs<<<div_up(s->nd, 256), 256, 0, st>>>(s->nd, s->rep.p, m->mass.p, s->div.p, -(1.0 / s->DT), s->rhs_red.p);
  FS_LAUNCH_CHECK();
  if (s->n_ex) {
    k_pressure_rhs_extra<<<div_up(s->n_ex, 128), 128, 0, st>>>(s->n_ex, s->ex_ptr.p, s->ex_dof.p, s->ex_node.p, m->mass.p,
                                                             s->div.p, -(1.0 / s->DT), s->rhs_red.p);
    FS_LAUNCH_CHECK();
  }
  if (!o.warm_start || !s->have_p) {
    FS_CUDA(cudaMemsetAsync(q, 0, s->nd * sizeof(double), st));
    H.nq = H.ny = 0;
    if (rec.ready()) rec.reset();
  } else if (recycle) {
    rec.guess(s->rhs_red.p, q, reduce);     // empty basis (after a restore without it): q keeps the previous solution
  } else if (extrap > 0) {
    const size_t bytes = s->nd * sizeof(double);
    int best = 0;   // 0: q, 1: linear, 2: quadratic
    bool have_y0 = false;
    if (H.nq >= 1) {
      const bool quad = H.nq >= 2 && extrap >= 2;
      spmv_best_dev(&s->k_red, q, s->y0.p);
      have_y0 = true;
      if (H.ny < 1) spmv_best_dev(&s->k_red, H.q1.p, H.y1.p);            // after a restore
      if (quad && H.ny < 2) spmv_best_dev(&s->k_red, H.q2.p, H.y2.p);
      double r3[3];
      cand_norms3_dev(&s->k_red, s->rhs_red.p, s->y0.p, H.y1.p, quad ? H.y2.p : nullptr, r3);
      if (r3[1] < r3[best]) best = 1;
      if (quad && r3[2] < r3[best]) best = 2;
      if (best == 1) lin3_dev(s->nd, 2.0, q, -1.0, H.q1.p, s->q_try.p);
      if (best == 2) {
        lin3_dev(s->nd, 3.0, q, -3.0, H.q1.p, s->q_try.p);
        lin3_dev(s->nd, 1.0, s->q_try.p, 1.0, H.q2.p, s->q_try.p);
      }
    }
    // shift the history: (q1, y1) -> (q2, y2) by swapping the buffers, then q -> q1, y0 -> y1
    std::swap(H.q1, H.q2);
    std::swap(H.y1, H.y2);
    FS_CUDA(cudaMemcpyAsync(H.q1.p, q, bytes, cudaMemcpyDeviceToDevice, st));
    if (have_y0) std::swap(H.y1, s->y0);
    H.ny = have_y0 ? 2 : 0;     // y1 = K q (just computed), y2 = the old y1 (valid or recomputed above)
    H.nq = std::min(H.nq + 1, 2);
    if (best) FS_CUDA(cudaMemcpyAsync(q, s->q_try.p, bytes, cudaMemcpyDeviceToDevice, st));
  }
  int it = cg_dev(&s->k_red, s->rhs_red.p, q, 1, o.rtol_pressure, o.maxit, o.precond, 1, relres);
  if (it < 0) throw Error(FS_ERR_NOCONV, "pressure CG did not converge within maxit");
  *iters = it;
  if (recycle) rec.update(q, matvec, reduce);
  k_expand<<<div_up(m->N, 256), 256, 0, st>>>(m->N, s->dof.p, q, p_full);
  FS_LAUNCH_CHECK();
}

// what the partitioned step (pstokes.cu) cuts its blocks out of
void stokes_global_view(fs_stokes* s, fs_mesh** mesh, fs_csr** a_visc, fs_csr** k_red, double* DT, double* nu, std::vector<int>& dof) {
  *mesh = s->mesh; *a_visc = &s->a_visc; *k_red = &s->k_red; *DT = s->DT; *nu = s->nu;
  dof = s->dof.to_host();
}

}  // namespace fs

using namespace fs;

extern "C" {

int fs_stokes_default_opts(fs_stokes_opts* o) {
  FS_API_BEGIN
  FS_REQUIRE(o, "opts is NULL");
  o->rtol_visc = 1e-12;
  o->rtol_pressure = 1e-10;
  o->maxit = 200000;
  o->precond = FS_PRECOND_AUTO;
  o->warm_start = 1;
  o->final_div = 0;
  o->bc_mode = 0;
  o->omega = 0.0;
  FS_API_END
}

int fs_stokes_create(fs_mesh* m, double DT, double nu, fs_stokes** out) {
  FS_API_BEGIN
  FS_REQUIRE(m && out, "NULL argument");
  *out = nullptr;
  FS_REQUIRE(m->bc_ready, "fs_bc_set must be called before fs_stokes_create");
  FS_REQUIRE(DT > 0, "DT must be positive");
  std::unique_ptr<fs_stokes> s(new fs_stokes());
  s->mesh = m; s->DT = DT; s->nu = nu;
  cudaStream_t st = stream();
  const int64_t N = m->N;
  ensure_geom(m);
  // K on the node pattern
  DBuf<double> kvals(m->pat.nnz);
  {
    int rc = fs_assemble_stiffness(m, kvals.p);
    if (rc != FS_OK) throw Error(rc, fs_last_error());
  }
  // Dirichlet / interior flags
  s->is_dir.alloc(N); s->is_dir.zero();
  s->is_interior.alloc(N); s->is_interior.zero();
  if (m->n_wall) { k_flag<<<div_up(m->n_wall, 256), 256, 0, st>>>(m->wall.p, m->n_wall, s->is_dir.p); FS_LAUNCH_CHECK(); }
  if (m->n_inner) { k_flag<<<div_up(m->n_inner, 256), 256, 0, st>>>(m->inner.p, m->n_inner, s->is_dir.p); FS_LAUNCH_CHECK(); }
  if (m->n_interior) { k_flag<<<div_up(m->n_interior, 256), 256, 0, st>>>(m->interior.p, m->n_interior, s->is_interior.p); FS_LAUNCH_CHECK(); }
  // A_visc
  s->a_visc.n = N; s->a_visc.nnz = m->pat.nnz;
  s->a_visc.rowptr = m->pat.rowptr.p; s->a_visc.colidx = m->pat.colidx.p;
  s->a_visc.vals.alloc(m->pat.nnz);
  CsrView Kv{(int)N, m->pat.nnz, m->pat.rowptr.p, m->pat.colidx.p, kvals.p};
  k_visc_vals<<<div_up(N, 256), 256, 0, st>>>(Kv, DT * nu, s->is_dir.p, s->a_visc.vals.p);
  FS_LAUNCH_CHECK();
  // periodic merge: union-find on the host (pairs are few), representative = smallest id,
  // dofs numbered in ascending representative order
  std::vector<int> parent(N);
  for (int64_t i = 0; i < N; ++i) parent[i] = (int)i;
  auto find = [&](int a) { while (parent[a] != a) { parent[a] = parent[parent[a]]; a = parent[a]; } return a; };
  for (int64_t k = 0; k < m->n_pairs; ++k) {
    int ra = find(m->pairs_host[2 * k]), rb = find(m->pairs_host[2 * k + 1]);
    if (ra != rb) parent[std::max(ra, rb)] = std::min(ra, rb);
  }
  std::vector<int> dof(N), newid(N, -1);
  int nd = 0;
  for (int64_t i = 0; i < N; ++i) if (find((int)i) == (int)i) newid[i] = nd++;
  for (int64_t i = 0; i < N; ++i) dof[i] = newid[find((int)i)];
  s->nd = nd;
  s->dof.alloc(N);
  s->dof.upload(dof.data(), N);
  {
    std::vector<int> rep(nd), cnt(nd, 0);
    for (int64_t i = 0; i < N; ++i) { if (newid[i] >= 0) rep[newid[i]] = (int)i; else ++cnt[dof[i]]; }
    std::vector<int> ex_dof, ex_ptr(1, 0), slot(nd, -1);
    for (int d = 0; d < nd; ++d) if (cnt[d]) { slot[d] = (int)ex_dof.size(); ex_dof.push_back(d); ex_ptr.push_back(ex_ptr.back() + cnt[d]); }
    std::vector<int> ex_node(ex_ptr.back()), fill(ex_ptr.begin(), ex_ptr.end() - 1);
    for (int64_t i = 0; i < N; ++i) if (newid[i] < 0) ex_node[fill[slot[dof[i]]]++] = (int)i;   // ascending node id
    s->rep.alloc(nd); s->rep.upload(rep.data(), nd);
    s->n_ex = (int64_t)ex_dof.size();
    if (s->n_ex) {
      s->ex_ptr.alloc(ex_ptr.size()); s->ex_ptr.upload(ex_ptr.data(), ex_ptr.size());
      s->ex_dof.alloc(ex_dof.size()); s->ex_dof.upload(ex_dof.data(), ex_dof.size());
      s->ex_node.alloc(ex_node.size()); s->ex_node.upload(ex_node.data(), ex_node.size());
    }
    fs::sync();   // the host vectors above go out of scope
  }
  fs::sync();
  build_pattern(m->tris.p, m->T, nd, s->dof.p, s->pat_red);
  s->k_red.n = nd; s->k_red.nnz = s->pat_red.nnz;
  s->k_red.rowptr = s->pat_red.rowptr.p; s->k_red.colidx = s->pat_red.colidx.p;
  s->k_red.vals.alloc(s->pat_red.nnz);
  assemble_on_pattern(s->pat_red, m->ke.p, s->k_red.vals.p);   // m->ke still holds the element matrices
  // the merged pattern's contribution lists are only needed for this assembly
  s->pat_red.contrib.release(); s->pat_red.seg_start.release(); s->pat_red.scatter.release();
  s->ustar.alloc(2 * N); s->div.alloc(N); s->rhs_red.alloc(nd);
  s->p_red.alloc(nd); s->p2_red.alloc(nd); s->p_full.alloc(N); s->p2_full.alloc(N);
  for (fs_stokes::PressHist* h : {&s->h1, &s->h2}) {
    h->q1.alloc(nd); h->q2.alloc(nd); h->y1.alloc(nd); h->y2.alloc(nd);
    h->q1.zero(); h->q2.zero();
  }
  s->y0.alloc(nd); s->q_try.alloc(nd);
  s->p_red.zero(); s->p2_red.zero(); s->p_full.zero(); s->p2_full.zero();
  fs::sync();
  *out = s.release();
  FS_API_END
}

int fs_stokes_destroy(fs_stokes* s) {
  FS_API_BEGIN
  if (s) { cudaStreamSynchronize(stream()); delete s; }
  FS_API_END
}

int fs_stokes_step(fs_stokes* s, double* u, double B1, double B2, const fs_stokes_opts* opts, fs_stokes_stats* stats) {
  FS_API_BEGIN
  FS_REQUIRE(s && u, "NULL argument");
  fs_stokes_opts o;
  fs_stokes_default_opts(&o);
  if (opts) o = *opts;
  fs_mesh* m = s->mesh;
  const int64_t N = m->N;
  cudaStream_t st = stream();
  Out<double> ou(u, 2 * N, true);
  double* du = ou.d;
  fs_stokes_stats sts;
  std::memset(&sts, 0, sizeof(sts));
  // FS_STEP_TIMING=1 (diagnostic): synchronise and print the wall time of each phase
  static const bool timing = std::getenv("FS_STEP_TIMING") != nullptr;
  double tmark[8];
  int nmark = 0;
  auto mark = [&]() {
    if (!timing) return;
    cudaStreamSynchronize(st);
    tmark[nmark++] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
  };
  mark();
  // Step 1: tentative velocity, both components in one 2-RHS CG started from u
  FS_CUDA(cudaMemcpyAsync(s->ustar.p, du, 2 * N * sizeof(double), cudaMemcpyDeviceToDevice, st));
  // A_visc = I + DT*nu*K has cond ~ 1: Jacobi is all it needs (the AMG option is for the pressure operator)
  const int pre_visc = (o.precond == FS_PRECOND_AMG || o.precond == FS_PRECOND_AUTO) ? FS_PRECOND_JACOBI : o.precond;
  int it = cg_dev(&s->a_visc, du, s->ustar.p, 2, o.rtol_visc, o.maxit, pre_visc, 0, &sts.relres_visc);
  if (it < 0) throw Error(FS_ERR_NOCONV, "viscous CG did not converge within maxit");
  sts.iters_visc = it;
  auto dirichlet = [&](double* v) {            // squirmer slip velocity, or the rotating-cylinder variant
    if (o.bc_mode == 1) rot_bcu_dev(m, v, o.omega, 0.5, 0.5);
    else dir_bcu_dev(m, v, B1, B2);
  };
  per_bcu_dev(m, s->ustar.p);
  dirichlet(s->ustar.p);
  mark();
  // Step 2+3: pressure correction and velocity update
  pressure_solve(s, s->ustar.p, s->p_red.p, s->h1, s->rec1, s->p_full.p, o, &sts.iters_p1, &sts.relres_p1,
                 o.final_div ? &sts.max_div_ustar : nullptr);
  mark();
  grad_update_dev(m, s->p_full.p, s->ustar.p, du, s->DT, nullptr);
  per_bcu_dev(m, du);
  dirichlet(du);
  mark();
  // second projection, interior nodes only, no BC re-imposition (:566-573)
  pressure_solve(s, du, s->p2_red.p, s->h2, s->rec2, s->p2_full.p, o, &sts.iters_p2, &sts.relres_p2, nullptr);
  mark();
  grad_update_dev(m, s->p2_full.p, du, du, s->DT, s->is_interior.p);
  s->have_p = true;
  if (o.final_div) {
    divergence_dev(m, du, s->div.p, nullptr);
    sts.max_final_div = max_abs_dev(s->div.p, N);
  }
  ou.commit();
  fs::sync();
  mark();
  if (timing && nmark == 6)
    std::fprintf(stderr, "[step] viscous %.0f us | pressure-1 %.0f us (%d it) | grad+bc %.0f us | pressure-2 %.0f us (%d it) | tail %.0f us\n",
                 tmark[1] - tmark[0], tmark[2] - tmark[1], sts.iters_p1, tmark[3] - tmark[2], tmark[4] - tmark[3], sts.iters_p2,
                 tmark[5] - tmark[4]);
  if (stats) *stats = sts;
  FS_API_END
}

int fs_stokes_pressure(fs_stokes* s, double* p, double* p2) {
  FS_API_BEGIN
  FS_REQUIRE(s, "NULL argument");
  const int64_t N = s->mesh->N;
  if (p) FS_CUDA(cudaMemcpyAsync(p, s->p_full.p, N * sizeof(double), cudaMemcpyDefault, stream()));
  if (p2) FS_CUDA(cudaMemcpyAsync(p2, s->p2_full.p, N * sizeof(double), cudaMemcpyDefault, stream()));
  fs::sync();
  FS_API_END
}

// ---- B configurations of a (B1, B2) sweep that share the mesh and therefore A_visc and Z^T K Z (BASELINE config 4):
// every kernel of the step runs once for all of them (grid.y = configuration; the single-CTA Krylov solvers as one
// CTA per configuration), instead of B launch-latency-bound step sequences.  B1, B2 enter only through makeDirBCU
// (code/StokesColor.py:419).  Small meshes only (the shipped ones): the solves use the one-CTA CG.
struct fs_stokes_batch {
  fs_stokes* s = nullptr;
  int B = 0;
  fs::DBuf<double> ustar, div, lump_a, lump_c, rhs, q1, q2, p_full, ws, out, b12;
  fs::DBuf<int> flags;
};

int fs_stokes_batch_create(fs_stokes* s, int32_t B, fs_stokes_batch** out) {
  FS_API_BEGIN
  FS_REQUIRE(s && out && B >= 1 && B <= 65535, "bad arguments");
  *out = nullptr;
  const int64_t N = s->mesh->N, T = s->mesh->T, nd = s->nd;
  FS_REQUIRE(N <= cg_small_limit(), "fs_stokes_batch: the mesh is too large for the batched single-CTA solvers; "
                                    "advance the configurations one by one (fs_stokes_step) or partition the mesh (fs_pstokes_*)");
  std::unique_ptr<fs_stokes_batch> b(new fs_stokes_batch());
  b->s = s; b->B = B;
  b->ustar.alloc((size_t)B * 2 * N); b->div.alloc((size_t)B * N); b->lump_a.alloc((size_t)B * T); b->lump_c.alloc((size_t)B * T);
  b->rhs.alloc((size_t)B * nd); b->q1.alloc((size_t)B * nd); b->q2.alloc((size_t)B * nd); b->p_full.alloc((size_t)B * N);
  b->ws.alloc((size_t)B * 4 * 2 * N); b->out.alloc((size_t)B * 4); b->flags.alloc((size_t)B * 2); b->b12.alloc((size_t)B * 2);
  b->q1.zero(); b->q2.zero();
  fs::sync();
  *out = b.release();
  FS_API_END
}

int fs_stokes_batch_destroy(fs_stokes_batch* b) {
  FS_API_BEGIN
  if (b) { cudaStreamSynchronize(stream()); delete b; }
  FS_API_END
}

int fs_stokes_step_batch(fs_stokes_batch* b, double* u, const double* b1b2, const fs_stokes_opts* opts, int32_t* iters) {
  FS_API_BEGIN
  FS_REQUIRE(b && u && b1b2, "NULL argument");
  fs_stokes_opts o;
  fs_stokes_default_opts(&o);
  if (opts) o = *opts;
  fs_stokes* s = b->s;
  fs_mesh* m = s->mesh;
  const int B = b->B;
  const int64_t N = m->N, nd = s->nd;
  cudaStream_t st = stream();
  Out<double> ou(u, (size_t)B * 2 * N, true);
  double* du = ou.d;
  FS_CUDA(cudaMemcpyAsync(b->b12.p, b1b2, (size_t)B * 2 * sizeof(double), cudaMemcpyDefault, st));
  const int pre_p = (o.precond == FS_PRECOND_NONE) ? FS_PRECOND_NONE : FS_PRECOND_JACOBI;
  std::vector<int> hf((size_t)B * 2);
  auto read_iters = [&](int col) {
    FS_CUDA(cudaMemcpyAsync(hf.data(), b->flags.p, hf.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
    FS_CUDA(cudaStreamSynchronize(st));
    for (int c = 0; c < B; ++c) {
      if (!hf[2 * c]) throw Error(FS_ERR_NOCONV, "fs_stokes_step_batch: a solve did not converge within maxit");
      if (iters) iters[3 * c + col] = hf[2 * c + 1];
    }
  };
  // viscous solve (2 right-hand sides per configuration), started from u
  FS_CUDA(cudaMemcpyAsync(b->ustar.p, du, (size_t)B * 2 * N * sizeof(double), cudaMemcpyDeviceToDevice, st));
  cg_small_batch_dev(&s->a_visc, du, b->ustar.p, 2, B, o.rtol_visc, o.maxit, pre_p, 0, b->ws.p, b->out.p, b->flags.p);
  if (iters) read_iters(0);
  bcu_batch_dev(m, B, b->ustar.p, b->b12.p);
  auto pressure = [&](const double* vel, double* q, int col) {
    divergence_batch_dev(m, B, vel, b->div.p, b->lump_a.p);
    k_pressure_rhs<<<dim3(div_up(nd, 256), B), 256, 0, st>>>(nd, s->rep.p, m->mass.p, b->div.p, -(1.0 / s->DT), b->rhs.p, N);
    FS_LAUNCH_CHECK();
    if (s->n_ex) {
      k_pressure_rhs_extra<<<dim3(div_up(s->n_ex, 128), B), 128, 0, st>>>(s->n_ex, s->ex_ptr.p, s->ex_dof.p, s->ex_node.p, m->mass.p,
                                                                        b->div.p, -(1.0 / s->DT), b->rhs.p, N, nd);
      FS_LAUNCH_CHECK();
    }
    if (!o.warm_start) FS_CUDA(cudaMemsetAsync(q, 0, (size_t)B * nd * sizeof(double), st));
    cg_small_batch_dev(&s->k_red, b->rhs.p, q, 1, B, o.rtol_pressure, o.maxit, pre_p, 1, b->ws.p, b->out.p, b->flags.p);
    if (iters) read_iters(col);
    k_expand<<<dim3(div_up(N, 256), B), 256, 0, st>>>(N, s->dof.p, q, b->p_full.p, nd);
    FS_LAUNCH_CHECK();
  };
  pressure(b->ustar.p, b->q1.p, 1);
  grad_update_batch_dev(m, B, b->p_full.p, b->ustar.p, du, s->DT, nullptr, b->lump_a.p, b->lump_c.p);
  bcu_batch_dev(m, B, du, b->b12.p);
  pressure(du, b->q2.p, 2);
  grad_update_batch_dev(m, B, b->p_full.p, du, du, s->DT, s->is_interior.p, b->lump_a.p, b->lump_c.p);
  ou.commit();
  if (!iters) {          // without per-solve reads: at least the last solve of every configuration must have converged
    FS_CUDA(cudaMemcpyAsync(hf.data(), b->flags.p, hf.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
    FS_CUDA(cudaStreamSynchronize(st));
    for (int c = 0; c < B; ++c)
      if (!hf[2 * c]) throw Error(FS_ERR_NOCONV, "fs_stokes_step_batch: a pressure solve did not converge within maxit");
  }
  fs::sync();
  FS_API_END
}

int fs_stokes_set_pressure(fs_stokes* s, const double* p, const double* p2) {
  FS_API_BEGIN
  FS_REQUIRE(s, "NULL argument");
  const int64_t N = s->mesh->N;
  if (p) FS_CUDA(cudaMemcpyAsync(s->p_full.p, p, N * sizeof(double), cudaMemcpyDefault, stream()));
  if (p2) FS_CUDA(cudaMemcpyAsync(s->p2_full.p, p2, N * sizeof(double), cudaMemcpyDefault, stream()));
  fs::sync();
  FS_API_END
}

int fs_stokes_matrices(fs_stokes* s, fs_csr** a_visc, fs_csr** k_pressure, int32_t* dof) {
  FS_API_BEGIN
  FS_REQUIRE(s, "NULL argument");
  if (a_visc) *a_visc = &s->a_visc;
  if (k_pressure) *k_pressure = &s->k_red;
  if (dof) {
    FS_CUDA(cudaMemcpyAsync(dof, s->dof.p, s->mesh->N * sizeof(int), cudaMemcpyDefault, stream()));
    fs::sync();
  }
  FS_API_END
}

int fs_stokes_warm_state(fs_stokes* s, double* q, int set) {
  FS_API_BEGIN
  FS_REQUIRE(s && q, "NULL argument");
  const size_t nd = (size_t)s->nd;
  cudaStream_t st = stream();
  double* bufs[6] = {s->p_red.p, s->p2_red.p, s->h1.q1.p, s->h2.q1.p, s->h1.q2.p, s->h2.q2.p};
  double hist[2] = {(double)s->h1.nq, (double)s->h2.nq};
  if (set) {
    for (int k = 0; k < 6; ++k) FS_CUDA(cudaMemcpyAsync(bufs[k], q + k * nd, nd * sizeof(double), cudaMemcpyDefault, st));
    FS_CUDA(cudaMemcpyAsync(hist, q + 6 * nd, sizeof(hist), cudaMemcpyDefault, st));
    fs::sync();
    s->h1.nq = std::max(0, std::min(2, (int)hist[0]));
    s->h2.nq = std::max(0, std::min(2, (int)hist[1]));
    s->h1.ny = s->h2.ny = 0;            // K q1, K q2 are recomputed by the next solve
    s->have_p = true;
  } else {
    for (int k = 0; k < 6; ++k) FS_CUDA(cudaMemcpyAsync(q + k * nd, bufs[k], nd * sizeof(double), cudaMemcpyDefault, st));
    FS_CUDA(cudaMemcpyAsync(q + 6 * nd, hist, sizeof(hist), cudaMemcpyDefault, st));
  }
  fs::sync();
  FS_API_END
}

int fs_stokes_recycle_state(fs_stokes* s, double* buf, int64_t cap, int set, int64_t* needed) {
  FS_API_BEGIN
  FS_REQUIRE(s, "NULL argument");
  Recycler* recs[2] = {&s->rec1, &s->rec2};
  if (!set) {
    int64_t need = 0;
    for (Recycler* r : recs) need += 1 + (r->ready() ? r->state_size() : 0);
    if (needed) *needed = need;
    if (buf && cap >= need) {
      double* p = buf;
      for (Recycler* r : recs) {
        const int64_t sz = r->ready() ? r->state_size() : 0;
        *p++ = (double)sz;
        if (sz) r->get_state(p);
        p += sz;
      }
    }
  } else {
    FS_REQUIRE(buf || cap == 0, "NULL argument");
    const double* p = buf;
    int64_t left = cap;
    for (Recycler* r : recs) {
      if (left <= 0) { if (r->ready()) r->reset(); continue; }       // a state without this part: start the basis afresh
      const int64_t sz = (int64_t)*p++;
      --left;
      FS_REQUIRE(sz >= 0 && sz <= left, "recycle state: truncated");
      if (sz) {
        if (!r->ready()) r->init(s->nd, recycle_kmax(), recycle_keep());
        r->set_state(p, sz);
      } else if (r->ready()) r->reset();
      p += sz;
      left -= sz;
    }
  }
  fs::sync();
  FS_API_END
}

}  // extern "C"
