// pstokes.cu -- the operator-split Stokes step of code/StokesColor.py:537-575 on a mesh that is cut into
// contiguous node blocks, one per GPU (BASELINE config 5).  Same sequence as stokes.cu:
//
//   u*  = solve(A_visc, u)                 2-RHS Jacobi-CG on this rank's rows           (:544-545)
//   makePerBCU(u*); makeDirBCU(u*)                                                        (:546-547)
//   p   = solve(Z^T K Z, Z^T M (-div u* / DT))   AMG-preconditioned CG, partitioned       (:551-555)
//   u   = u* - DT grad p; BCs                                                             (:559-564)
//   p2  = solve(..., -div u / DT);  u[interior] -= DT grad p2                             (:567-573)
//
// Every rank holds: its rows of A_visc and Z^T K Z (SELL-32, columns renumbered to [own | halo]), its
// rows of the AMG operators of the large levels (amg.cu), replicated copies of the small levels, and a
// sub-mesh made of the elements that touch its nodes (divergence / gradient / BC kernels run on that
// sub-mesh unchanged and evaluate the owned nodes only).  All vectors that are read across block
// boundaries live in a CUDA-IPC arena; halo values and dot-product partial sums travel as peer stores
// issued from inside the kernels (dist.cuh) -- there is no host, NCCL or copy-engine call inside a step.
// Setup is replicated: every rank builds the global operators and hierarchy once (so the partitioned
// solver has exactly the single-GPU hierarchy), keeps its blocks and frees the rest.
#include "dist.cuh"
#include "recycle.cuh"
#include "reduce.cuh"

namespace fs {

// ---- small kernels ---------------------------------------------------------------------------------
// sum the per-CTA partials (nblk x K) of this rank, then over all ranks; one CTA; bumps the sequence number
template <int K>
__global__ void __launch_bounds__(kBlock) k_allreduce(const double* __restrict__ part, int nblk, Comm c, double* __restrict__ out) {
  __shared__ double sm[K];
  double v[K];
  reduce_partials<K>(part, nblk, v, sm);
  const unsigned long long seq = dist_seq(c);
  rank_allreduce<K>(c, v, seq + 1);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) out[k] = v[k];
    dist_seq_bump(c);
  }
}

template <int K>
__global__ void __launch_bounds__(kBlock) k_dotk(int64_t n, const double* __restrict__ a0, const double* __restrict__ b0,
                                                const double* __restrict__ a1, const double* __restrict__ b1,
                                                double* __restrict__ part) {
  __shared__ double red[K * 32];
  double acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    acc[0] += a0[i] * (b0 ? b0[i] : 1.0);
    if (K > 1) acc[K - 1] += a1[i] * b1[i];
  }
  block_reduce<K>(acc, red);
  if (threadIdx.x == 0)
#pragma unroll
    for (int k = 0; k < K; ++k) part[(size_t)blockIdx.x * K + k] = acc[k];
}

// out = in - sum / n_global
__global__ void k_sub_mean_g(int64_t n, const double* in, double* out, const double* __restrict__ sum, double inv_ng) {
  const double mean = *sum * inv_ng;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = in[i] - mean;
}

// out = a x + b y + c z (y, z may be null); out may alias an input
__global__ void k_axpbypcz(int64_t n, double a, const double* x, double b, const double* y, double c, const double* z, double* out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double v = a * x[i];
    if (y) v += b * y[i];
    if (z) v += c * z[i];
    out[i] = v;
  }
}

// |b - y0|^2, |b - (2 y0 - y1)|^2, |b - (3 y0 - 3 y1 + y2)|^2 partials (see stokes.cu: warm start)
__global__ void __launch_bounds__(kBlock)
k_cand3p(int64_t n, const double* __restrict__ b, const double* __restrict__ y0, const double* __restrict__ y1,
         const double* __restrict__ y2, double* __restrict__ part) {
  __shared__ double red[3 * 32];
  double acc[3] = {0.0, 0.0, 0.0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double bv = b[i], a0 = y0[i], a1 = y1[i];
    const double r0 = bv - a0, r1 = bv - (2.0 * a0 - a1);
    acc[0] += r0 * r0;
    acc[1] += r1 * r1;
    if (y2) { const double r2 = bv - (3.0 * a0 - 3.0 * a1 + y2[i]); acc[2] += r2 * r2; }
  }
  block_reduce<3>(acc, red);
  if (threadIdx.x == 0) for (int k = 0; k < 3; ++k) part[3 * blockIdx.x + k] = acc[k];
}

// ---- AMG-PCG on the partitioned pressure operator (device-side scalars as in solver.cu) -------------
//   sc[0], sc[1]  r.z of the previous / current iteration (slot = iteration parity)
//   sc[2] b.b   sc[3] latest r.r   sc[4] tol^2      flags[0] converged   flags[1] iterations done
__global__ void __launch_bounds__(kBlock)
k_ppcg_init(const double* __restrict__ partRZ, int nrz, const double* __restrict__ bbrr, double tol2, double* __restrict__ sc,
            int* __restrict__ flags, Comm c) {
  __shared__ double sm[1];
  double rz[1];
  reduce_partials<1>(partRZ, nrz, rz, sm);
  const unsigned long long seq = dist_seq(c);
  rank_allreduce<1>(c, rz, seq + 1);
  if (threadIdx.x == 0) {
    sc[0] = rz[0]; sc[1] = 0.0; sc[2] = bbrr[0]; sc[3] = bbrr[1]; sc[4] = tol2;
    flags[0] = 0; flags[1] = 0;
    dist_seq_bump(c);
  }
}

// x += alpha p ; r -= alpha Ap with alpha = rz / (p.Ap summed over ranks); partial r.r out
__global__ void __launch_bounds__(kBlock)
k_ppcg_xr(int64_t n, const double* __restrict__ p, const double* __restrict__ Ap, double* __restrict__ x, double* __restrict__ r,
          const double* __restrict__ partA, int nblkA, const double* __restrict__ sc, int slot, const int* __restrict__ flags,
          double* __restrict__ partB, Comm c, PushSpec ps) {
  __shared__ double red[32];
  __shared__ double sm[1];
  if (block_done(flags)) return;
  dist_trace(c, 20);
  double pAp[1];
  reduce_partials<1>(partA, nblkA, pAp, sm);
  const unsigned long long seq = dist_seq(c);
  dist_trace(c, 21);
  rank_allreduce<1>(c, pAp, seq + 1);
  dist_trace(c, 22);
  const double rz = sc[slot];
  const double alpha = (pAp[0] != 0.0) ? rz / pAp[0] : 0.0;
  double acc[1] = {0.0};
  // one warp per 32-row slice; the slices with rows that other ranks read come first: their new residuals are stored
  // into the neighbours' halo slots and the flags released while the rest of the vector is still being updated
  // (the first nb CTAs do nothing else, the others share the rest: nobody carries boundary work on top of a full share)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kW = kBlock / 32;
  const int per_warp = max(1, (int)((n + 31) >> 5) / (2 * (int)gridDim.x * kW));      // half an interior warp's share
  const int nb = ps.enabled ? min((int)gridDim.x / 2, (ps.n_slist + kW * per_warp - 1) / (kW * per_warp)) : 0;
  const int nsl = (int)((n + 31) >> 5);
  auto upd = [&](int64_t i, bool send) -> bool {
    const double rn = r[i] - alpha * Ap[i];
    x[i] += alpha * p[i];
    r[i] = rn;
    acc[0] += rn * rn;
    return send ? push_row(ps, (int)i, rn) : false;
  };
  if ((int)blockIdx.x < nb) {
    bool pushed = false;
    for (int j = blockIdx.x * kW + warp; j < ps.n_slist; j += nb * kW) {
      const int64_t i = ((int64_t)__ldg(ps.slist + j) << 5) + lane;
      if (i < n) pushed |= upd(i, true);
    }
    push_finish(c, ps, pushed, seq + 1, nb);
  } else {
    for (int sl = ((int)blockIdx.x - nb) * kW + warp; sl < nsl; sl += ((int)gridDim.x - nb) * kW) {
      if (nb && __ldg(ps.smask + sl)) continue;
      const int64_t i = ((int64_t)sl << 5) + lane;
      if (i < n) upd(i, false);
    }
  }
  block_reduce<1>(acc, red);
  if (threadIdx.x == 0) partB[blockIdx.x] = acc[0];
  dist_trace(c, 23);
  dist_trace_last(c, 26);
  if (dist_last_block(c, 0)) dist_seq_bump(c);
}

// convergence test, beta, p = z + beta p.  Every CTA takes the same decision from the all-reduced sums; the
// CTA that finishes last publishes the scalars (so the flag cannot change under a CTA of the same launch).
__global__ void __launch_bounds__(kBlock)
k_ppcg_p(int64_t n, const double* __restrict__ z, double* __restrict__ p, const double* __restrict__ partB, int nB,
         const double* __restrict__ partRZ, int nRZ, double* __restrict__ sc, int slot, int* __restrict__ flags, Comm c,
         PushSpec ps) {
  __shared__ double sm[1];
  if (block_done(flags)) return;
  double v[2], t[1];
  reduce_partials<1>(partB, nB, t, sm);
  v[0] = t[0];
  reduce_partials<1>(partRZ, nRZ, t, sm);
  v[1] = t[0];
  const unsigned long long seq = dist_seq(c);
  dist_trace(c, 31);
  rank_allreduce<2>(c, v, seq + 1);
  dist_trace(c, 32);
  const double rz_old = sc[slot];
  const bool conv = v[0] <= sc[4] * sc[2];
  if (!conv) {
    const double beta = rz_old != 0.0 ? v[1] / rz_old : 0.0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kW = kBlock / 32;
    const int per_warp = max(1, (int)((n + 31) >> 5) / (2 * (int)gridDim.x * kW));      // half an interior warp's share
    const int nb = ps.enabled ? min((int)gridDim.x / 2, (ps.n_slist + kW * per_warp - 1) / (kW * per_warp)) : 0;
    const int nsl = (int)((n + 31) >> 5);
    if ((int)blockIdx.x < nb) {                                         // rows the neighbours read: dedicated CTAs, sent as produced
      bool pushed = false;
      for (int j = blockIdx.x * kW + warp; j < ps.n_slist; j += nb * kW) {
        const int64_t i = ((int64_t)__ldg(ps.slist + j) << 5) + lane;
        if (i < n) { const double pn = z[i] + beta * p[i]; p[i] = pn; pushed |= push_row(ps, (int)i, pn); }
      }
      push_finish(c, ps, pushed, seq + 1, nb);
    } else {
      for (int sl = ((int)blockIdx.x - nb) * kW + warp; sl < nsl; sl += ((int)gridDim.x - nb) * kW) {
        if (nb && __ldg(ps.smask + sl)) continue;
        const int64_t i = ((int64_t)sl << 5) + lane;
        if (i < n) p[i] = z[i] + beta * p[i];
      }
    }
  }
  dist_trace(c, 33);
  dist_trace_last(c, 36);
  if (dist_last_block(c, 1)) {
    sc[slot ^ 1] = v[1];
    sc[3] = v[0];
    flags[1] += 1;
    if (conv) flags[0] = 1;
    dist_seq_bump(c);
  }
}

// ---- 2-RHS Jacobi-CG on the partitioned viscous operator ---------------------------------------------
// scalars (doubles): rz[2] slot 0 at 0, slot 1 at 2 ; bb[2] at 4 ; rr[2] at 6
__global__ void __launch_bounds__(kBlock)
k_pv_init(int64_t n, const double* __restrict__ b, const double* __restrict__ Ap, const double* __restrict__ dinv,
          double* __restrict__ r, double* __restrict__ p, double* __restrict__ part) {
  __shared__ double red[6 * 32];
  double acc[6] = {0, 0, 0, 0, 0, 0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double bv[2], av[2], rv[2], zv[2];
    load_vec<2>(b, i, bv);
    load_vec<2>(Ap, i, av);
    const double di = dinv[i];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      rv[c] = bv[c] - av[c];
      zv[c] = di * rv[c];
      acc[c] += rv[c] * rv[c];
      acc[2 + c] += rv[c] * zv[c];
      acc[4 + c] += bv[c] * bv[c];
    }
    store_vec<2>(r, i, rv);
    store_vec<2>(p, i, zv);
  }
  block_reduce<6>(acc, red);
  if (threadIdx.x == 0)
#pragma unroll
    for (int k = 0; k < 6; ++k) part[(size_t)blockIdx.x * 6 + k] = acc[k];
}

__global__ void k_pv_init_fin(const double* __restrict__ s6 /* rr[2], rz[2], bb[2] summed over ranks */, double* __restrict__ sc,
                              double tol2, int* __restrict__ flags) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  bool all = true;
  for (int c = 0; c < 2; ++c) {
    sc[c] = s6[2 + c];
    sc[2 + c] = 0.0;
    sc[4 + c] = s6[4 + c];
    sc[6 + c] = s6[c];
    if (!(s6[c] <= tol2 * s6[4 + c])) all = false;
  }
  flags[0] = all ? 1 : 0;
  flags[1] = 0;
}

__global__ void __launch_bounds__(kBlock)
k_pv_xr(int64_t n, const double* __restrict__ p, const double* __restrict__ Ap, const double* __restrict__ dinv,
        double* __restrict__ x, double* __restrict__ r, const double* __restrict__ partA, int nblkA, const double* __restrict__ sc,
        int slot, const int* __restrict__ flags, double* __restrict__ partB, Comm c) {
  __shared__ double red[4 * 32];
  __shared__ double sm[2];
  if (block_done(flags)) return;
  double pAp[2], alpha[2];
  reduce_partials<2>(partA, nblkA, pAp, sm);
  const unsigned long long seq = dist_seq(c);
  rank_allreduce<2>(c, pAp, seq + 1);
#pragma unroll
  for (int k = 0; k < 2; ++k) alpha[k] = (pAp[k] != 0.0) ? sc[slot * 2 + k] / pAp[k] : 0.0;
  double acc[4] = {0, 0, 0, 0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double pv[2], av[2], xv[2], rv[2];
    load_vec<2>(p, i, pv);
    load_vec<2>(Ap, i, av);
    load_vec<2>(x, i, xv);
    load_vec<2>(r, i, rv);
    const double di = dinv[i];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      xv[k] += alpha[k] * pv[k];
      rv[k] -= alpha[k] * av[k];
      acc[k] += rv[k] * rv[k];
      acc[2 + k] += rv[k] * (di * rv[k]);
    }
    store_vec<2>(x, i, xv);
    store_vec<2>(r, i, rv);
  }
  block_reduce<4>(acc, red);
  if (threadIdx.x == 0)
#pragma unroll
    for (int k = 0; k < 4; ++k) partB[(size_t)blockIdx.x * 4 + k] = acc[k];
  if (dist_last_block(c, 2)) dist_seq_bump(c);
}

__global__ void __launch_bounds__(kBlock)
k_pv_p(int64_t n, const double* __restrict__ r, const double* __restrict__ dinv, double* __restrict__ p,
       const double* __restrict__ partB, int nblkB, double* __restrict__ sc, int slot, double tol2, int* __restrict__ flags, Comm c) {
  __shared__ double sm[4];
  if (block_done(flags)) return;
  double v[4], beta[2];
  reduce_partials<4>(partB, nblkB, v, sm);
  const unsigned long long seq = dist_seq(c);
  rank_allreduce<4>(c, v, seq + 1);
  bool all = true;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const double rz_old = sc[slot * 2 + k];
    beta[k] = (rz_old != 0.0) ? v[2 + k] / rz_old : 0.0;
    if (!(v[k] <= tol2 * sc[4 + k])) all = false;
  }
  if (!all)
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
      double rv[2], pv[2];
      load_vec<2>(r, i, rv);
      load_vec<2>(p, i, pv);
      const double di = dinv[i];
#pragma unroll
      for (int k = 0; k < 2; ++k) pv[k] = di * rv[k] + beta[k] * pv[k];
      store_vec<2>(p, i, pv);
    }
  if (dist_last_block(c, 3)) {
#pragma unroll
    for (int k = 0; k < 2; ++k) { sc[(slot ^ 1) * 2 + k] = v[2 + k]; sc[6 + k] = v[k]; }
    flags[1] += 1;
    if (all) flags[0] = 1;
    dist_seq_bump(c);
  }
}

__global__ void k_diag_inv_l(CsrView A, double* __restrict__ dinv) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= A.n) return;
  double d = 0.0;
  for (int k = A.rowptr[row]; k < A.rowptr[row + 1]; ++k)
    if (A.colidx[k] == row) d = A.vals[k];
  dinv[row] = (d != 0.0) ? 1.0 / d : 1.0;
}

// ---- pressure right-hand side / expansion on the local numbering --------------------------------------
__global__ void k_prhs(int64_t nd, const int* __restrict__ rep, const double* __restrict__ mass, const double* __restrict__ div,
                       double s, double* __restrict__ rhs) {
  const int64_t d = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (d >= nd) return;
  const int n = rep[d];
  rhs[d] = mass[n] * (s * div[n]);
}
__global__ void k_prhs_extra(int64_t n_ex, const int* __restrict__ ex_ptr, const int* __restrict__ ex_dof, const int* __restrict__ ex_node,
                             const double* __restrict__ mass, const double* __restrict__ div, double s, double* __restrict__ rhs) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= n_ex) return;
  double v = rhs[ex_dof[j]];
  for (int k = ex_ptr[j]; k < ex_ptr[j + 1]; ++k) { const int n = ex_node[k]; v += mass[n] * (s * div[n]); }
  rhs[ex_dof[j]] = v;
}
__global__ void k_expand_l(int64_t n, const int* __restrict__ ldof, const double* __restrict__ q, double* __restrict__ p) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) p[i] = q[ldof[i]];
}
__global__ void k_fill_pattern(int64_t n, int64_t g0, double* __restrict__ v) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) v[i] = (double)((g0 + i) % 13) - 6.0 + 0.25 * (double)((g0 + i) % 5);
}
__global__ void k_flag_l(const int* __restrict__ idx, int64_t n, unsigned char* __restrict__ flag) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k < n) flag[idx[k]] = 1;
}

static int vgrid_n(int64_t n) {
  const int64_t want = (n + kBlock - 1) / kBlock;
  const int64_t cap = std::min<int64_t>(kMaxBlocks, (int64_t)sm_count() * 6);
  return (int)std::max<int64_t>(1, std::min(want, cap));
}

}  // namespace fs

// =======================================================================================================
struct fs_pstokes {
  int rank = 0, world = 1;
  fs_mesh* lmesh = nullptr;         // sub-mesh of this rank (borrowed)
  double DT = 0, nu = 0;
  int64_t N_glob = 0, nd_glob = 0;
  fs::DistCtx ctx;
  fs::Space nodes, dofs;
  fs::fs_sell a_visc, k_red;        // own rows, fp64 SELL-32, local columns
  fs::DBuf<double> dinv_visc;
  fs::Amg* amg = nullptr;
  // node space (stride 2 vectors in the arena, plain work arrays)
  fs::DVec U, USTAR, PV;
  fs::DBuf<double> rv, apv, div, p_loc, p2_loc;
  fs::DBuf<int> ldof, rep, ex_ptr, ex_dof, ex_node;
  int64_t n_ex = 0;
  fs::DBuf<unsigned char> is_interior;
  // dof space
  fs::DVec P, R, Q1, Q2, QTRY;
  fs::DBuf<double> Ap, z, bproj, rhs, y0;
  struct Hist { fs::DBuf<double> q1, q2, y1, y2; int nq = 0, ny = 0; } h1, h2;
  fs::Recycler rec1, rec2;          // solution-subspace projection of the two pressure solves (recycle.cuh), own rows
  fs::DBuf<double> rec_red;         // its dot products, summed over the ranks
  bool have_p = false;
  fs::DBuf<double> partials, scal, red_out;
  int hint = 0, hint_prev = 0;
  size_t planned_bytes = 0;
  ~fs_pstokes() { if (amg) fs::amg_free(amg); }
};

namespace fs {

static void allreduce_k(fs_pstokes* s, const double* part, int nblk, int K, double* out) {
  cudaStream_t st = stream();
  switch (K) {
    case 1: k_allreduce<1><<<1, kBlock, 0, st>>>(part, nblk, s->ctx.comm, out); break;
    case 2: k_allreduce<2><<<1, kBlock, 0, st>>>(part, nblk, s->ctx.comm, out); break;
    case 3: k_allreduce<3><<<1, kBlock, 0, st>>>(part, nblk, s->ctx.comm, out); break;
    case 6: k_allreduce<6><<<1, kBlock, 0, st>>>(part, nblk, s->ctx.comm, out); break;
    default: throw Error(FS_ERR_INTERNAL, "allreduce_k: unsupported width");
  }
  FS_LAUNCH_CHECK();
}

// CG preconditioned by the partitioned V-cycle; X (arena vector: initial guess in, solution out, mean removed).
static int ppcg_amg(fs_pstokes* s, const double* b_in, DVec& X, double rtol, int maxit, double* relres) {
  DistCtx& ctx = s->ctx;
  const Comm& cm = ctx.comm;
  cudaStream_t st = stream();
  const int64_t n = s->dofs.n_own;
  const double inv_ng = 1.0 / (double)s->nd_glob;
  const int g = vgrid_n(n);
  double* part = s->partials.p;
  double* partA = part + kMaxBlocks * 4;
  double* partB = part + kMaxBlocks * 6;
  double* partRZ = part + kMaxBlocks * 8;
  double* sc = s->scal.p;
  int* flags = ctx.flags.p;
  double* red = s->red_out.p;          // [0] sum b, [1..2] bb rr, [3] sum x
  double *r = s->R.p, *p = s->P.p, *Ap = s->Ap.p, *z = s->z.p, *b = s->bproj.p;
  FS_CUDA(cudaMemsetAsync(flags, 0, 4 * sizeof(int), st));
  // b <- b - mean(b) over all ranks
  k_dotk<1><<<g, kBlock, 0, st>>>(n, b_in, nullptr, nullptr, nullptr, part); FS_LAUNCH_CHECK();
  allreduce_k(s, part, g, 1, red);
  k_sub_mean_g<<<g, kBlock, 0, st>>>(n, b_in, b, red, inv_ng); FS_LAUNCH_CHECK();
  // r = b - A x
  ctx.push(X);
  spmv_sell_dist(s->k_red, X.p, Ap, nullptr, nullptr, cm, ctx.wait_of(&X));
  k_axpbypcz<<<g, kBlock, 0, st>>>(n, 1.0, b, -1.0, Ap, 0.0, nullptr, r); FS_LAUNCH_CHECK();
  k_dotk<2><<<g, kBlock, 0, st>>>(n, b, b, r, r, part); FS_LAUNCH_CHECK();
  allreduce_k(s, part, g, 2, red + 1);
  double h2[2];
  FS_CUDA(cudaMemcpyAsync(h2, red + 1, sizeof(h2), cudaMemcpyDeviceToHost, st));
  FS_CUDA(cudaStreamSynchronize(st));
  const double bb = h2[0];
  double rr = h2[1];
  const double tol2 = rtol * rtol;
  int it = 0;
  bool done = (bb == 0.0) || rr <= tol2 * bb;
  if (bb == 0.0) FS_CUDA(cudaMemsetAsync(X.p, 0, n * sizeof(double), st));
  if (!done) {
    ctx.push(s->R);
    int nrz = amg_apply_dist(s->amg, s->R, z, partRZ);
    FS_CUDA(cudaMemcpyAsync(p, z, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    k_ppcg_init<<<1, kBlock, 0, st>>>(partRZ, nrz, red + 1, tol2, sc, flags, cm); FS_LAUNCH_CHECK();
    ctx.push(s->P);
    const int unchecked = std::max(0, std::min(std::min(s->hint, s->hint_prev) - 3, maxit));
    int queued = 0, hflags[2] = {0, 0};
    const PushSpec psR = ctx.push_spec(s->R), psP = ctx.push_spec(s->P);
    const HaloWait wP = ctx.wait_of(&s->P);
    while (queued < maxit) {
      const int slot = queued & 1;
      const int ga = spmv_sell_dist(s->k_red, p, Ap, nullptr, partA, cm, wP, nullptr, 1);
      k_ppcg_xr<<<g, kBlock, 0, st>>>(n, p, Ap, X.p, r, partA, ga, sc, slot, flags, partB, cm, psR); FS_LAUNCH_CHECK();
      nrz = amg_apply_dist(s->amg, s->R, z, partRZ);
      k_ppcg_p<<<g, kBlock, 0, st>>>(n, z, p, partB, g, partRZ, nrz, sc, slot, flags, cm, psP); FS_LAUNCH_CHECK();
      ++queued;
      if (queued > unchecked) {
        int e = 0;   // a time-out ends the solve instead of queueing maxit empty iterations
        FS_CUDA(cudaMemcpyAsync(hflags, flags, sizeof(hflags), cudaMemcpyDeviceToHost, st));
        FS_CUDA(cudaMemcpyAsync(&e, ctx.err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        FS_CUDA(cudaStreamSynchronize(st));
        if (hflags[0] || e) break;
      }
    }
    double hsc[5];
    FS_CUDA(cudaMemcpyAsync(hflags, flags, sizeof(hflags), cudaMemcpyDeviceToHost, st));
    FS_CUDA(cudaMemcpyAsync(hsc, sc, sizeof(hsc), cudaMemcpyDeviceToHost, st));
    FS_CUDA(cudaStreamSynchronize(st));
    done = hflags[0] != 0;
    it = hflags[1];
    rr = hsc[3];
    if (done) { s->hint_prev = s->hint; s->hint = it; }
    FS_CUDA(cudaMemsetAsync(flags, 0, 4 * sizeof(int), st));   // the kernels below must not see "converged"
    if (!done) ctx.check("partitioned pressure CG");             // (a time-out is reported; otherwise the step's final check suffices)
  }
  // x <- x - mean(x); publish its boundary values (the caller expands it to the nodes)
  k_dotk<1><<<g, kBlock, 0, st>>>(n, X.p, nullptr, nullptr, nullptr, part); FS_LAUNCH_CHECK();
  allreduce_k(s, part, g, 1, red + 3);
  k_sub_mean_g<<<g, kBlock, 0, st>>>(n, X.p, X.p, red + 3, inv_ng); FS_LAUNCH_CHECK();
  ctx.push(X);
  if (relres) *relres = bb > 0.0 ? std::sqrt(rr / bb) : 0.0;
  return done ? it : -it - 1;
}

// 2-RHS Jacobi-CG: USTAR = solve(A_visc, U), started from U.  U's halo has been pushed by the caller.
static int visc_solve(fs_pstokes* s, double rtol, int maxit, double* relres) {
  DistCtx& ctx = s->ctx;
  const Comm& cm = ctx.comm;
  cudaStream_t st = stream();
  const int64_t n = s->nodes.n_own;
  const int g = vgrid_n(n);
  double* part0 = s->partials.p;
  double* partA = part0 + kMaxBlocks * 6;
  double* partB = part0 + kMaxBlocks * 8;
  double* sc = s->scal.p + 16;
  int* flags = ctx.flags.p;
  double* x = s->USTAR.p;
  const double tol2 = rtol * rtol;
  FS_CUDA(cudaMemsetAsync(flags, 0, 4 * sizeof(int), st));
  FS_CUDA(cudaMemcpyAsync(x, s->U.p, 2 * n * sizeof(double), cudaMemcpyDeviceToDevice, st));
  spmv_sell2_dist(s->a_visc, s->U.p, s->apv.p, nullptr, nullptr, cm, ctx.wait_of(&s->U));
  k_pv_init<<<g, kBlock, 0, st>>>(n, s->U.p, s->apv.p, s->dinv_visc.p, s->rv.p, s->PV.p, part0); FS_LAUNCH_CHECK();
  allreduce_k(s, part0, g, 6, s->red_out.p + 8);
  k_pv_init_fin<<<1, 32, 0, st>>>(s->red_out.p + 8, sc, tol2, flags); FS_LAUNCH_CHECK();
  struct { double d[8]; int f[2]; } hs;
  auto poll = [&]() {
    FS_CUDA(cudaMemcpyAsync(hs.d, sc, sizeof(hs.d), cudaMemcpyDeviceToHost, st));
    FS_CUDA(cudaMemcpyAsync(hs.f, flags, sizeof(hs.f), cudaMemcpyDeviceToHost, st));
    FS_CUDA(cudaStreamSynchronize(st));
  };
  poll();
  if (hs.d[4] == 0.0 && hs.d[5] == 0.0) {
    FS_CUDA(cudaMemsetAsync(x, 0, 2 * n * sizeof(double), st));
    if (relres) *relres = 0.0;
    return 0;
  }
  ctx.push(s->PV);
  const int ga = spmv_sell_grid(s->a_visc);
  int launched = 0, slot = 0;
  while (!hs.f[0] && launched < maxit) {
    const int todo = std::min(6, maxit - launched);
    for (int k = 0; k < todo; ++k) {
      spmv_sell2_dist(s->a_visc, s->PV.p, s->apv.p, partA, flags, cm, ctx.wait_of(&s->PV));
      k_pv_xr<<<g, kBlock, 0, st>>>(n, s->PV.p, s->apv.p, s->dinv_visc.p, x, s->rv.p, partA, ga, sc, slot, flags, partB, cm);
      FS_LAUNCH_CHECK();
      k_pv_p<<<g, kBlock, 0, st>>>(n, s->rv.p, s->dinv_visc.p, s->PV.p, partB, g, sc, slot, tol2, flags, cm);
      FS_LAUNCH_CHECK();
      ctx.push(s->PV);
      slot ^= 1;
    }
    launched += todo;
    int e = 0;
    FS_CUDA(cudaMemcpyAsync(&e, ctx.err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    poll();
    if (e) break;
  }
  if (!hs.f[0]) ctx.check("partitioned viscous CG");
  FS_CUDA(cudaMemsetAsync(flags, 0, 4 * sizeof(int), st));
  double worst = 0.0;
  for (int c = 0; c < 2; ++c) if (hs.d[4 + c] > 0.0) worst = std::max(worst, std::sqrt(hs.d[6 + c] / hs.d[4 + c]));
  if (relres) *relres = worst;
  return hs.f[0] ? hs.f[1] : -hs.f[1] - 1;
}

// y = (Z^T K Z) x for an arena vector x (own rows out)
static void kmul(fs_pstokes* s, DVec& X, double* y) {
  s->ctx.push(X);
  spmv_sell_dist(s->k_red, X.p, y, nullptr, nullptr, s->ctx.comm, s->ctx.wait_of(&X));
}
static void barrier_dev(fs_pstokes* s) {   // a reduction of nothing: separates two pushes of the same channel
  FS_CUDA(cudaMemsetAsync(s->partials.p, 0, sizeof(double), stream()));
  allreduce_k(s, s->partials.p, 1, 1, s->red_out.p + 7);
}

// divergence -> mass-weighted merged right-hand side -> warm start -> PCG -> nodal pressure on [own | halo] nodes
static void pressure_solve(fs_pstokes* s, DVec& VEL, DVec& Q, fs_pstokes::Hist& H, Recycler& rec, double* p_loc,
                           const fs_stokes_opts& o, int* iters, double* relres) {
  static const int extrap = [] { const char* e = std::getenv("FS_STOKES_EXTRAP"); return e ? std::atoi(e) : 2; }();
  // the same rule as the single-GPU step (stokes.cu), on the GLOBAL system size: every rank decides alike
  const bool recycle = o.warm_start && s->nd_glob >= 20000 && recycle_enabled();
  // K d: through the scratch arena vector (halo exchange inside); a reduction separates two pushes of its channel
  const Recycler::MatVec matvec = [s](const double* x, double* y) {
    barrier_dev(s);
    FS_CUDA(cudaMemcpyAsync(s->QTRY.p, x, (size_t)s->dofs.n_own * sizeof(double), cudaMemcpyDeviceToDevice, stream()));
    kmul(s, s->QTRY, y);
  };
  // sums over this rank's CTAs and over the ranks, six at a time: identical on every rank, so the bases stay consistent
  const Recycler::Reduce reduce = [s](const double* part, int nblk, int nchunk, double* host) {
    if (!s->rec_red.p) s->rec_red.alloc(kRecCap);
    for (int c = 0; c < nchunk; ++c) allreduce_k(s, part + (size_t)c * nblk * 6, nblk, 6, s->rec_red.p + 6 * c);
    FS_CUDA(cudaMemcpyAsync(host, s->rec_red.p, (size_t)6 * nchunk * sizeof(double), cudaMemcpyDeviceToHost, stream()));
    FS_CUDA(cudaStreamSynchronize(stream()));
  };
  if (recycle && !rec.ready()) rec.init(s->dofs.n_own, recycle_kmax(), recycle_keep());
  DistCtx& ctx = s->ctx;
  fs_mesh* m = s->lmesh;
  cudaStream_t st = stream();
  const int64_t nd = s->dofs.n_own;
  const int g = vgrid_n(nd);
  ctx.wait(VEL);                                        // halo velocities (pushed by the caller)
  divergence_dev(m, VEL.p, s->div.p, nullptr);
  k_prhs<<<div_up(nd, 256), 256, 0, st>>>(nd, s->rep.p, m->mass.p, s->div.p, -(1.0 / s->DT), s->rhs.p); FS_LAUNCH_CHECK();
  if (s->n_ex) {
    k_prhs_extra<<<div_up(s->n_ex, 128), 128, 0, st>>>(s->n_ex, s->ex_ptr.p, s->ex_dof.p, s->ex_node.p, m->mass.p, s->div.p,
                                                     -(1.0 / s->DT), s->rhs.p);
    FS_LAUNCH_CHECK();
  }
  if (!o.warm_start || !s->have_p) {
    FS_CUDA(cudaMemsetAsync(Q.p, 0, nd * sizeof(double), st));
    H.nq = H.ny = 0;
    if (rec.ready()) rec.reset();
  } else if (recycle) {
    rec.guess(s->rhs.p, Q.p, reduce);
  } else if (extrap > 0) {
    const size_t bytes = nd * sizeof(double);
    int best = 0;
    bool have_y0 = false;
    if (H.nq >= 1) {
      const bool quad = H.nq >= 2 && extrap >= 2;
      kmul(s, Q, s->y0.p);
      have_y0 = true;
      if (H.ny < 1) {          // after a restore: K q1 (and K q2) through the scratch arena vector
        barrier_dev(s);
        FS_CUDA(cudaMemcpyAsync(s->QTRY.p, H.q1.p, bytes, cudaMemcpyDeviceToDevice, st));
        kmul(s, s->QTRY, H.y1.p);
      }
      if (quad && H.ny < 2) {
        barrier_dev(s);
        FS_CUDA(cudaMemcpyAsync(s->QTRY.p, H.q2.p, bytes, cudaMemcpyDeviceToDevice, st));
        kmul(s, s->QTRY, H.y2.p);
      }
      // residual norms of the three candidates, summed over ranks (the mean of the right-hand side adds the same
      // constant to all three: the images K q are mean-free)
      k_cand3p<<<g, kBlock, 0, st>>>(nd, s->rhs.p, s->y0.p, H.y1.p, quad ? H.y2.p : nullptr, s->partials.p);
      FS_LAUNCH_CHECK();
      allreduce_k(s, s->partials.p, g, 3, s->red_out.p + 4);
      double r3[3];
      FS_CUDA(cudaMemcpyAsync(r3, s->red_out.p + 4, sizeof(r3), cudaMemcpyDeviceToHost, st));
      FS_CUDA(cudaStreamSynchronize(st));
      if (r3[1] < r3[best]) best = 1;
      if (quad && r3[2] < r3[best]) best = 2;
      if (best == 1) { k_axpbypcz<<<g, kBlock, 0, st>>>(nd, 2.0, Q.p, -1.0, H.q1.p, 0.0, nullptr, s->QTRY.p); FS_LAUNCH_CHECK(); }
      if (best == 2) { k_axpbypcz<<<g, kBlock, 0, st>>>(nd, 3.0, Q.p, -3.0, H.q1.p, 1.0, H.q2.p, s->QTRY.p); FS_LAUNCH_CHECK(); }
    }
    std::swap(H.q1, H.q2);
    std::swap(H.y1, H.y2);
    FS_CUDA(cudaMemcpyAsync(H.q1.p, Q.p, bytes, cudaMemcpyDeviceToDevice, st));
    if (have_y0) std::swap(H.y1, s->y0);
    H.ny = have_y0 ? 2 : 0;
    H.nq = std::min(H.nq + 1, 2);
    if (best) FS_CUDA(cudaMemcpyAsync(Q.p, s->QTRY.p, bytes, cudaMemcpyDeviceToDevice, st));
  }
  const int it = ppcg_amg(s, s->rhs.p, Q, o.rtol_pressure, o.maxit, relres);
  if (it < 0) throw Error(FS_ERR_NOCONV, "partitioned pressure CG did not converge within maxit");
  *iters = it;
  ctx.wait(Q);                                          // ppcg_amg pushed the final solution
  // (a halo flag carries the sequence number of its push: it has to be consumed before the next reduction bumps it)
  if (recycle) rec.update(Q.p, matvec, reduce);
  k_expand_l<<<div_up(m->N, 256), 256, 0, st>>>(m->N, s->ldof.p, Q.p, p_loc); FS_LAUNCH_CHECK();
}

}  // namespace fs

using namespace fs;

extern "C" {

int fs_pstokes_create(fs_stokes* glob, fs_mesh* lmesh, int rank, int world, const int64_t* node_split, const int32_t* l2g,
                      int gather_rows, fs_pstokes** out) {
  FS_API_BEGIN
  FS_REQUIRE(glob && lmesh && node_split && l2g && out, "NULL argument");
  *out = nullptr;
  FS_REQUIRE(world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world, "bad rank / world (at most 8 ranks)");
  FS_REQUIRE(lmesh->bc_ready, "fs_bc_set must have been called on the local mesh");
  fs_mesh* gm = nullptr;
  fs_csr *gav = nullptr, *gk = nullptr;
  double DT = 0, nu = 0;
  std::vector<int> dof;
  stokes_global_view(glob, &gm, &gav, &gk, &DT, &nu, dof);
  const int64_t N = gm->N, nd = gk->n;
  std::unique_ptr<fs_pstokes> s(new fs_pstokes());
  s->rank = rank; s->world = world; s->lmesh = lmesh; s->DT = DT; s->nu = nu; s->N_glob = N; s->nd_glob = nd;
  cudaStream_t st = stream();
  // ---- node space: halo = the neighbours of the owned nodes in the mesh graph (pattern of K)
  s->nodes.split.assign(node_split, node_split + world + 1);
  FS_REQUIRE(s->nodes.split.front() == 0 && s->nodes.split.back() == N, "node split must cover [0, N]");
  for (int r = 0; r < world; ++r) FS_REQUIRE(s->nodes.split[r] < s->nodes.split[r + 1], "every rank needs at least one node");
  {
    HaloCollector c;
    c.add_matrix(gav->view(), s->nodes.split, s->nodes.split, 0, N);
    c.finish(s->nodes, rank, world);
  }
  const int64_t n_own = s->nodes.n_own, n_loc = n_own + s->nodes.n_halo, lo = s->nodes.own_lo;
  FS_REQUIRE(lmesh->N == n_loc, "local mesh does not have n_own + n_halo nodes (own nodes first, then the halo nodes in ascending global id)");
  for (int64_t i = 0; i < n_own; ++i) FS_REQUIRE(l2g[i] == lo + i, "local mesh: own nodes must come first, in global order");
  for (int64_t i = 0; i < s->nodes.n_halo; ++i)
    FS_REQUIRE(l2g[n_own + i] == s->nodes.halo_all[rank][i], "local mesh: halo nodes must follow in ascending global id");
  lmesh->n_active = n_own;
  // ---- dof space: dofs are numbered by ascending representative node, so the node split induces a dof split;
  // every node merged into a dof must sit on the rank that owns the dof (periodic pairs are not cut)
  std::vector<int64_t> dsplit(world + 1, 0);
  {
    std::vector<int> rep(nd, -1);
    for (int64_t i = N - 1; i >= 0; --i) rep[dof[i]] = (int)i;           // smallest node of each dof
    for (int r = 1; r < world; ++r) {
      const int64_t nsr = s->nodes.split[r];
      dsplit[r] = std::lower_bound(rep.begin(), rep.end(), (int)nsr) - rep.begin();   // rep is ascending in the dof id
    }
    dsplit[world] = nd;
    for (int64_t i = 0; i < N; ++i) {
      const int rn = s->nodes.rank_of(i);
      const int rd = (int)(std::upper_bound(dsplit.begin(), dsplit.end(), (int64_t)dof[i]) - dsplit.begin()) - 1;
      FS_REQUIRE(rn == rd, "the node split cuts a periodic pair: both nodes of a merged pair must belong to one rank");
    }
    // local lists for the right-hand side: representative node of each own dof, merged-in nodes in ascending order
    const int64_t d0 = dsplit[rank], d1 = dsplit[rank + 1];
    std::vector<int> lrep(d1 - d0), cnt(d1 - d0, 0);
    for (int64_t d = d0; d < d1; ++d) lrep[d - d0] = rep[d] - (int)lo;
    for (int64_t i = lo; i < lo + n_own; ++i) if (rep[dof[i]] != (int)i) ++cnt[dof[i] - d0];
    std::vector<int> ex_dof, ex_ptr(1, 0), slot(d1 - d0, -1);
    for (int64_t d = 0; d < d1 - d0; ++d) if (cnt[d]) { slot[d] = (int)ex_dof.size(); ex_dof.push_back((int)d); ex_ptr.push_back(ex_ptr.back() + cnt[d]); }
    std::vector<int> ex_node(ex_ptr.back()), fill(ex_ptr.begin(), ex_ptr.end() - 1);
    for (int64_t i = lo; i < lo + n_own; ++i) if (rep[dof[i]] != (int)i) ex_node[fill[slot[dof[i] - d0]]++] = (int)(i - lo);
    s->rep.alloc(lrep.size()); s->rep.upload(lrep.data(), lrep.size());
    s->n_ex = (int64_t)ex_dof.size();
    if (s->n_ex) {
      s->ex_ptr.alloc(ex_ptr.size()); s->ex_ptr.upload(ex_ptr.data(), ex_ptr.size());
      s->ex_dof.alloc(ex_dof.size()); s->ex_dof.upload(ex_dof.data(), ex_dof.size());
      s->ex_node.alloc(ex_node.size()); s->ex_node.upload(ex_node.data(), ex_node.size());
    }
    FS_CUDA(cudaStreamSynchronize(st));
  }
  s->dofs.split = dsplit;
  // ---- the hierarchy (replicated setup), then the halo lists of the dof space
  AmgPartSpec spec;
  spec.rank = rank; spec.world = world; spec.split0 = dsplit; spec.gather_rows = gather_rows > 0 ? gather_rows : 100000;
  s->amg = amg_setup(gk, &spec);
  {
    HaloCollector c0;
    c0.add_matrix(gk->view(), dsplit, dsplit, 0, nd);
    s->planned_bytes = amg_part_collect(s->amg, c0);
    for (int q = 0; q < world; ++q)                               // the dofs of the halo nodes (pressure expansion)
      for (int h : s->nodes.halo_all[q]) c0.add(q, dof[h]);
    c0.finish(s->dofs, rank, world);
  }
  // ---- arena and its vectors (identical layout on every rank)
  auto vbytes = [](const Space& sp, int stride) { return ((size_t)sp.cap * stride * sizeof(double) + 255) / 256 * 256; };
  s->planned_bytes += 3 * vbytes(s->nodes, 2) + 5 * vbytes(s->dofs, 1);
  s->ctx.init(rank, world, s->planned_bytes);
  s->U = s->ctx.carve(s->nodes, 2);
  s->USTAR = s->ctx.carve(s->nodes, 2);
  s->PV = s->ctx.carve(s->nodes, 2);
  s->P = s->ctx.carve(s->dofs, 1);
  s->R = s->ctx.carve(s->dofs, 1);
  s->Q1 = s->ctx.carve(s->dofs, 1);
  s->Q2 = s->ctx.carve(s->dofs, 1);
  s->QTRY = s->ctx.carve(s->dofs, 1);
  amg_part_finalize(s->amg, s->dofs, s->ctx);
  // ---- this rank's rows of the two fine operators
  {
    fs_csr loc;
    extract_rows(gav->view(), lo, lo + n_own, s->nodes, N, nullptr, loc);
    sell_build(loc, false, s->a_visc);
    s->dinv_visc.alloc(n_own);
    k_diag_inv_l<<<div_up(n_own, 256), 256, 0, st>>>(loc.view(), s->dinv_visc.p);
    FS_LAUNCH_CHECK();
    FS_CUDA(cudaStreamSynchronize(st));
  }
  {
    fs_csr loc;
    extract_rows(gk->view(), dsplit[rank], dsplit[rank + 1], s->dofs, nd, nullptr, loc);
    sell_build(loc, false, s->k_red);
    sell_mark_boundary(s->k_red, loc, (int)s->dofs.n_own, 0x7fffffff, -1, nullptr);
  }
  // ---- local node -> local dof (own dofs first, then the dof halo list)
  {
    std::vector<int> ldof(n_loc);
    const std::vector<int>& hd = s->dofs.halo_all[rank];
    for (int64_t i = 0; i < n_loc; ++i) {
      const int d = dof[l2g[i]];
      if (d >= dsplit[rank] && d < dsplit[rank + 1]) ldof[i] = d - (int)dsplit[rank];
      else {
        auto it = std::lower_bound(hd.begin(), hd.end(), d);
        FS_REQUIRE(it != hd.end() && *it == d, "internal: dof of a halo node missing from the dof halo list");
        ldof[i] = (int)(s->dofs.n_own + (it - hd.begin()));
      }
    }
    s->ldof.alloc(n_loc); s->ldof.upload(ldof.data(), n_loc);
    FS_CUDA(cudaStreamSynchronize(st));
  }
  s->is_interior.alloc(n_loc); s->is_interior.zero();
  if (lmesh->n_interior) { k_flag_l<<<div_up(lmesh->n_interior, 256), 256, 0, st>>>(lmesh->interior.p, lmesh->n_interior, s->is_interior.p); FS_LAUNCH_CHECK(); }
  ensure_geom(lmesh);
  const int64_t ndl = s->dofs.n_own;
  s->rv.alloc(2 * n_own); s->apv.alloc(2 * n_own); s->div.alloc(n_loc); s->p_loc.alloc(n_loc); s->p2_loc.alloc(n_loc);
  s->p_loc.zero(); s->p2_loc.zero();
  s->Ap.alloc(ndl); s->z.alloc(ndl); s->bproj.alloc(ndl); s->rhs.alloc(ndl); s->y0.alloc(ndl);
  for (fs_pstokes::Hist* h : {&s->h1, &s->h2}) { h->q1.alloc(ndl); h->q2.alloc(ndl); h->y1.alloc(ndl); h->y2.alloc(ndl); h->q1.zero(); h->q2.zero(); }
  s->partials.alloc((size_t)kMaxBlocks * 16); s->scal.alloc(64); s->red_out.alloc(32);
  s->scal.zero(); s->red_out.zero();
  fs::sync();
  *out = s.release();
  FS_API_END
}

int fs_pstokes_destroy(fs_pstokes* s) {
  FS_API_BEGIN
  if (s) { cudaStreamSynchronize(stream()); delete s; }
  FS_API_END
}

int fs_pstokes_sizes(const fs_pstokes* s, int64_t* n_own_nodes, int64_t* n_halo_nodes, int64_t* n_own_dofs, int64_t* n_halo_dofs,
                     int32_t* levels_partitioned) {
  FS_API_BEGIN
  FS_REQUIRE(s, "NULL argument");
  if (n_own_nodes) *n_own_nodes = s->nodes.n_own;
  if (n_halo_nodes) *n_halo_nodes = s->nodes.n_halo;
  if (n_own_dofs) *n_own_dofs = s->dofs.n_own;
  if (n_halo_dofs) *n_halo_dofs = s->dofs.n_halo;
  if (levels_partitioned) *levels_partitioned = amg_part_levels(s->amg);
  FS_API_END
}

int fs_pstokes_ipc_handle(fs_pstokes* s, void* handle64) {
  FS_API_BEGIN
  FS_REQUIRE(s && handle64, "NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  FS_CUDA(cudaIpcGetMemHandle(&h, s->ctx.arena.p));
  std::memcpy(handle64, &h, 64);
  FS_API_END
}

int fs_pstokes_connect(fs_pstokes* s, const void* all_handles) {
  FS_API_BEGIN
  FS_REQUIRE(s, "NULL argument");
  s->ctx.connect(all_handles);
  FS_API_END
}

int fs_pstokes_step(fs_pstokes* s, double* u_own, double B1, double B2, const fs_stokes_opts* opts, fs_stokes_stats* stats) {
  FS_API_BEGIN
  FS_REQUIRE(s && u_own, "NULL argument");
  FS_REQUIRE(s->ctx.connected, "fs_pstokes_connect has not been called");
  fs_stokes_opts o;
  fs_stokes_default_opts(&o);
  if (opts) o = *opts;
  fs_mesh* m = s->lmesh;
  const int64_t n = s->nodes.n_own;
  cudaStream_t st = stream();
  DistCtx& ctx = s->ctx;
  fs_stokes_stats sts;
  std::memset(&sts, 0, sizeof(sts));
  FS_CUDA(cudaMemcpyAsync(s->U.p, u_own, 2 * n * sizeof(double), cudaMemcpyDefault, st));
  ctx.push(s->U);
  int it = visc_solve(s, o.rtol_visc, o.maxit, &sts.relres_visc);
  if (it < 0) throw Error(FS_ERR_NOCONV, "partitioned viscous CG did not converge within maxit");
  sts.iters_visc = it;
  auto dirichlet = [&](double* v) {
    if (o.bc_mode == 1) rot_bcu_dev(m, v, o.omega, 0.5, 0.5);
    else dir_bcu_dev(m, v, B1, B2);
  };
  per_bcu_dev(m, s->USTAR.p);
  dirichlet(s->USTAR.p);
  ctx.push(s->USTAR);
  pressure_solve(s, s->USTAR, s->Q1, s->h1, s->rec1, s->p_loc.p, o, &sts.iters_p1, &sts.relres_p1);
  grad_update_dev(m, s->p_loc.p, s->USTAR.p, s->U.p, s->DT, nullptr);
  per_bcu_dev(m, s->U.p);
  dirichlet(s->U.p);
  ctx.push(s->U);
  pressure_solve(s, s->U, s->Q2, s->h2, s->rec2, s->p2_loc.p, o, &sts.iters_p2, &sts.relres_p2);
  grad_update_dev(m, s->p2_loc.p, s->U.p, s->U.p, s->DT, s->is_interior.p);
  s->have_p = true;
  // the next step pushes U again: separate the two pushes of that channel by a reduction
  barrier_dev(s);
  FS_CUDA(cudaMemcpyAsync(u_own, s->U.p, 2 * n * sizeof(double), cudaMemcpyDefault, st));
  fs::sync();
  ctx.check("fs_pstokes_step");
  if (stats) *stats = sts;
  FS_API_END
}

// Profiling aid: `iters` PCG iterations on the right-hand side of the last pressure solve (x0 = 0, no convergence
// test), timed with a CUDA event pair on the library stream.  Collective.
int fs_pstokes_profile_pcg(fs_pstokes* s, int iters, double* us_per_iter) {
  FS_API_BEGIN
  FS_REQUIRE(s && us_per_iter && iters > 0, "bad arguments");
  cudaStream_t st = stream();
  cudaEvent_t e0, e1;
  FS_CUDA(cudaEventCreate(&e0));
  FS_CUDA(cudaEventCreate(&e1));
  const int hint = s->hint, hint_prev = s->hint_prev;
  k_fill_pattern<<<div_up(s->dofs.n_own, 256), 256, 0, st>>>(s->dofs.n_own, s->dofs.own_lo, s->rhs.p);   // synthetic right-hand side
  FS_LAUNCH_CHECK();
  for (int pass = 0; pass < 3; ++pass) {       // passes 0, 1 warm up (graph capture), pass 2 is timed
    FS_CUDA(cudaMemsetAsync(s->QTRY.p, 0, s->dofs.n_own * sizeof(double), st));
    barrier_dev(s);
    s->hint = s->hint_prev = iters + 3;
    FS_CUDA(cudaEventRecord(e0, st));
    double rr = 0.0;
    ppcg_amg(s, s->rhs.p, s->QTRY, 0.0, iters, &rr);
    FS_CUDA(cudaEventRecord(e1, st));
    FS_CUDA(cudaEventSynchronize(e1));
  }
  float ms = 0.f;
  FS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  *us_per_iter = 1e3 * ms / iters;
  s->hint = hint; s->hint_prev = hint_prev;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  barrier_dev(s);
  fs::sync();
  FS_API_END
}

// FS_DIST_TRACE=1: copy out (and reset) the event trace: pairs {tag, ns}.  Returns the number of events in *n.
int fs_pstokes_trace(fs_pstokes* s, uint64_t* out, int64_t cap, int64_t* n) {
  FS_API_BEGIN
  FS_REQUIRE(s && n, "NULL argument");
  *n = 0;
  if (!s->ctx.comm.trace) return FS_OK;
  fs::sync();
  unsigned cnt = s->ctx.trace_n.to_host()[0];
  cnt = std::min<unsigned>(cnt, s->ctx.comm.trace_cap);
  const int64_t m = std::min<int64_t>(cnt, cap);
  if (out && m) FS_CUDA(cudaMemcpy(out, s->ctx.trace.p, (size_t)m * 16, cudaMemcpyDeviceToHost));
  s->ctx.trace_n.zero();
  fs::sync();
  *n = m;
  FS_API_END
}

int fs_pstokes_pressure(fs_pstokes* s, double* p_own, double* p2_own) {
  FS_API_BEGIN
  FS_REQUIRE(s, "NULL argument");
  const int64_t n = s->nodes.n_own;
  if (p_own) FS_CUDA(cudaMemcpyAsync(p_own, s->p_loc.p, n * sizeof(double), cudaMemcpyDefault, stream()));
  if (p2_own) FS_CUDA(cudaMemcpyAsync(p2_own, s->p2_loc.p, n * sizeof(double), cudaMemcpyDefault, stream()));
  fs::sync();
  FS_API_END
}

// warm-start state: q, q1, q2, y1, y2 of both pressure solves (10 n_own_dofs doubles) + {nq1, ny1, nq2, ny2}
int fs_pstokes_state(fs_pstokes* s, double* buf, int set) {
  FS_API_BEGIN
  FS_REQUIRE(s && buf, "NULL argument");
  const size_t nd = (size_t)s->dofs.n_own;
  cudaStream_t st = stream();
  double* bufs[10] = {s->Q1.p, s->h1.q1.p, s->h1.q2.p, s->h1.y1.p, s->h1.y2.p, s->Q2.p, s->h2.q1.p, s->h2.q2.p, s->h2.y1.p, s->h2.y2.p};
  double meta[4] = {(double)s->h1.nq, (double)s->h1.ny, (double)s->h2.nq, (double)s->h2.ny};
  if (set) {
    for (int k = 0; k < 10; ++k) FS_CUDA(cudaMemcpyAsync(bufs[k], buf + k * nd, nd * sizeof(double), cudaMemcpyDefault, st));
    FS_CUDA(cudaMemcpyAsync(meta, buf + 10 * nd, sizeof(meta), cudaMemcpyDefault, st));
    fs::sync();
    s->h1.nq = (int)meta[0]; s->h1.ny = (int)meta[1]; s->h2.nq = (int)meta[2]; s->h2.ny = (int)meta[3];
    s->have_p = true;
  } else {
    for (int k = 0; k < 10; ++k) FS_CUDA(cudaMemcpyAsync(buf + k * nd, bufs[k], nd * sizeof(double), cudaMemcpyDefault, st));
    FS_CUDA(cudaMemcpyAsync(buf + 10 * nd, meta, sizeof(meta), cudaMemcpyDefault, st));
  }
  fs::sync();
  FS_API_END
}

// the projection bases of the two pressure solves (own rows); same protocol as fs_stokes_recycle_state
int fs_pstokes_recycle_state(fs_pstokes* s, double* buf, int64_t cap, int set, int64_t* needed) {
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
      if (left <= 0) { if (r->ready()) r->reset(); continue; }
      const int64_t sz = (int64_t)*p++;
      --left;
      FS_REQUIRE(sz >= 0 && sz <= left, "recycle state: truncated");
      if (sz) {
        if (!r->ready()) r->init(s->dofs.n_own, recycle_kmax(), recycle_keep());
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
