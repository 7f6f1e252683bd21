// amg.cu -- smoothed-aggregation algebraic multigrid as a CG preconditioner for the
// pressure operator (the iteration count of the Jacobi-CG grows
// like 1/h -- ~8 000 iterations per solve at 4M triangles -- and is what limits steps/s).
//
// Setup (device only, once per matrix):
//   aggregates of ~4 rows from two passes of pairwise "handshake" matching along the
//   strongest negative coupling (deterministic: ties go to the smaller index);
//   prolongator P = (I - w D_F^-1 A_F) P_tent with piecewise-constant P_tent and A_F = A with the
//   couplings weaker than 0.25 x the row's strongest lumped onto the diagonal (the constants
//   stay in the range of P, so every level keeps the null space of the Neumann operator);
//   coarse operator P^T (A P).  All sparse products are expand / radix-sort / reduce-by-key
//   on the GPU.  Levels are added until <= 2048 rows; that last operator is inverted densely.
// Application (one symmetric V(1,1) cycle, damped Jacobi, fixed => a valid CG
//   preconditioner): x = w D^-1 b; r = b - A x; b_c = P^T r; recurse; x += P x_c;
//   x += w D^-1 (b - A x).  Coarsest level: dense (pseudo-)inverse GEMV.
//   Folded form (default): with S = I - w D^-1 A the same cycle is
//       b_c = R~ b,            R~ = P^T S^T = (S P)^T          (one SpMV down)
//       x   = G b + P~ x_c,    P~ = S P,  G = w D^-1 (2I - w A D^-1)   (one SpMV up, on [G | P~])
//   i.e. two dependent kernels per level instead of four and no intermediate vectors; P~ comes
//   for free from the A P product of the Galerkin step.  The unfolded form (FS_AMG_FOLD=0) fuses
//   smoother / residual into SpMV epilogues instead.  Matrices are streamed as fp32 copies
//   (vectors and sums fp64).
#include <cub/cub.cuh>

#include "dist.cuh"

namespace fs {

struct AmgLevel {
  int part_nsplit = 0;      // partitioned cycle: width of the [own | halo] right-hand side this rank's rows of U refer to
  fs_csr A;                 // operator of this level (level 0 borrows the fine matrix)
  const fs_csr* Aref = nullptr;
  int n = 0;
  fs_csr P, PT;             // smoothed prolongator (n x n_coarse) and its transpose (restriction)
  fs_csr U, Rt;             // folded cycle: U = [G | S P] (n x (n + n_coarse)),  Rt = (S P)^T
  fs_sell Us, Rts;          // ... and their SELL-32 copies (what the cycle streams; CSR kept on small levels)
  DBuf<double> x, b, r, t;  // work vectors of this level (level 0 uses caller buffers for b/x)
  DBuf<float> x32, b32;     // fp32 mirrors of x / b, written by the kernels that produce x / b for the fp32 gathers of the
                            // packed SELL kernels (spmv_sell.cu, FMT 4); allocated when the caller keeps a mirror of r
  const fs_csr& mat() const { return Aref ? *Aref : A; }
};

struct Amg {
  std::vector<std::unique_ptr<AmgLevel>> L;
  double omega = 2.0 / 3.0;     // damped-Jacobi smoother
  double omega_p = 2.0 / 3.0;   // prolongator smoothing
  double theta = 0.25;          // strength threshold of the filtered prolongator smoothing (0 = unfiltered)
  int coarse_sweeps = 40;
  DBuf<double> coarse_inv;      // dense (pseudo-)inverse of the coarsest operator (n <= 2048), row-major
  int coarse_n = 0;
  // the V-cycle as a CUDA graph (captured on its second application; one launch per cycle)
  cudaGraphExec_t graph = nullptr;
  const double* graph_r = nullptr;
  const float* graph_r32 = nullptr;
  const float* r32 = nullptr;   // transient: fp32 mirror of the finest right-hand side (amg_apply's r32)
  double* graph_z = nullptr;
  bool graph_x0 = false;
  bool graph_failed = false;
  double* graph_part = nullptr;
  int graph_nparts = 0;
  bool folded = true;           // two-kernel-per-level form of the cycle
  bool sell = true;             // folded operators in SELL-32 layout
  cudaEvent_t* top_ev = nullptr; // transient: events around the finest up-sweep (amg_apply's top_ev)
  int sub_rows = 0;             // matrices with at most this many rows use the lanes-per-row CSR kernel
  int applications = 0;
  // Partitioned application (dist.cuh, pstokes.cu): levels [0, Lp) are cut into contiguous row blocks, one per
  // rank -- each rank keeps only its rows of R~ and [G | P~], columns renumbered to [own | halo] -- and level Lp
  // and below are replicated: the restriction into level Lp is computed block-wise and stored into every rank.
  bool part_f32 = true;
  struct Part {
    DistCtx* ctx = nullptr;
    int rank = 0, world = 1, Lp = 0;
    std::vector<std::vector<int64_t>> split;     // [0..Lp] row split of each level
    std::vector<std::unique_ptr<Space>> space;   // [1..Lp] (entry 0 unused: level 0's space belongs to the caller)
    const Space* space0 = nullptr;
    std::vector<DVec> vb, vx;                    // [1..Lp] right-hand side / solution of each level inside the arena
  } part;
  ~Amg() { if (graph) cudaGraphExecDestroy(graph); }
};

void amg_free(Amg* a) { delete a; }

// ---------------------------------------------------------------- matching kernels
// Edge priority: a symmetric pseudo-random hash of the edge.  Picking "the strongest"
// neighbour with index tie-breaks degenerates on structured meshes (every node of a chain
// prefers the same side and only one pair per chain matches per round); instead every
// STRONG neighbour (coupling >= 0.5 x the row's largest) is a candidate and the edge with the
// highest hash wins, so a constant fraction of the rows finds a mutual partner every round.
__device__ __forceinline__ unsigned long long edge_hash(int i, int j) {
  unsigned long long z = ((unsigned long long)(unsigned)min(i, j) << 32) | (unsigned)max(i, j);
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

__global__ void k_pick(CsrView A, const int* __restrict__ state, int* __restrict__ best) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  int bj = -1;
  if (state[i] < 0) {
    double wmax = 0.0;
    for (int k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k)
      if (A.colidx[k] != i) wmax = fmax(wmax, -A.vals[k]);
    const double thr = 0.5 * wmax;
    unsigned long long bh = 0;
    for (int k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) {
      const int j = A.colidx[k];
      if (j == i || state[j] >= 0) continue;
      const double w = -A.vals[k];
      if (!(w > 0.0) || w < thr) continue;
      const unsigned long long h = edge_hash(i, j);
      if (bj < 0 || h > bh || (h == bh && j < bj)) { bh = h; bj = j; }
    }
  }
  best[i] = bj;
}

__global__ void k_match(int n, const int* __restrict__ best, int* __restrict__ state, int* __restrict__ partner) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int j = best[i];
  if (j >= 0 && best[j] == i) { partner[i] = j; state[i] = 1; }
}

// leftovers join the aggregate of their strongest already-matched neighbour (or stay alone)
__global__ void k_leftover(CsrView A, const int* __restrict__ state, const int* __restrict__ partner,
                           int* __restrict__ leader) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  if (state[i] >= 0) { const int j = partner[i]; leader[i] = (j >= 0 && j < i) ? j : i; return; }
  int bj = -1;
  double bw = 0.0;
  for (int k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) {
    const int j = A.colidx[k];
    if (j == i || state[j] < 0) continue;
    const double w = -A.vals[k];
    if (w > bw || (w == bw && w > 0.0 && j < bj)) { bw = w; bj = j; }
  }
  if (bj < 0) leader[i] = i;
  else { const int pj = partner[bj]; leader[i] = (pj >= 0 && pj < bj) ? pj : bj; }
}

__global__ void k_is_leader(int n, const int* __restrict__ leader, int* __restrict__ flag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = (leader[i] == i) ? 1 : 0;
}

__global__ void k_agg_id(int n, const int* __restrict__ leader, const int* __restrict__ scan_excl, int* __restrict__ agg) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) agg[i] = scan_excl[leader[i]];
}

__global__ void k_compose(int n, const int* a1, const int* __restrict__ a2, int* out /* may be a1 */) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a2[a1[i]];
}

// one pairwise pass: agg (n) and the number of aggregates
static int pairwise(const CsrView& A, DBuf<int>& agg) {
  cudaStream_t st = stream();
  const int n = A.n, B = 256, g = div_up(n, B);
  DBuf<int> state(n), best(n), partner(n), leader(n), flag(n), scan(n);
  FS_CUDA(cudaMemsetAsync(state.p, 0xff, n * sizeof(int), st));
  FS_CUDA(cudaMemsetAsync(partner.p, 0xff, n * sizeof(int), st));
  for (int round = 0; round < 8; ++round) {
    k_pick<<<g, B, 0, st>>>(A, state.p, best.p);
    FS_LAUNCH_CHECK();
    k_match<<<g, B, 0, st>>>(n, best.p, state.p, partner.p);
    FS_LAUNCH_CHECK();
  }
  k_leftover<<<g, B, 0, st>>>(A, state.p, partner.p, leader.p);
  FS_LAUNCH_CHECK();
  k_is_leader<<<g, B, 0, st>>>(n, leader.p, flag.p);
  FS_LAUNCH_CHECK();
  size_t bytes = 0;
  FS_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, flag.p, scan.p, n, st));
  DBuf<char> tmp(bytes);
  FS_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, flag.p, scan.p, n, st));
  count_launch(2);
  agg.alloc(n);
  k_agg_id<<<g, B, 0, st>>>(n, leader.p, scan.p, agg.p);
  FS_LAUNCH_CHECK();
  int last_flag = 0, last_scan = 0;
  FS_CUDA(cudaMemcpyAsync(&last_flag, flag.p + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
  FS_CUDA(cudaMemcpyAsync(&last_scan, scan.p + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
  FS_CUDA(cudaStreamSynchronize(st));
  return last_flag + last_scan;
}

// ---------------------------------------------------------------- sparse products (expand / sort / compress)
// coo (key = row<<32 | col, value) -> CSR with duplicates summed.  n_rows rows.
__global__ void k_split_keys(const unsigned long long* __restrict__ keys, int m, int* __restrict__ rowof, int* __restrict__ col) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= m) return;
  rowof[k] = (int)(keys[k] >> 32);
  col[k] = (int)(keys[k] & 0xffffffffu);
}

__global__ void k_rowptr_from_rowof(const int* __restrict__ rowof, int nnz, int n, int* __restrict__ rowptr) {
  int z = blockIdx.x * blockDim.x + threadIdx.x;
  if (z > nnz) return;
  int lo = (z == 0) ? -1 : rowof[z - 1];
  int hi = (z == nnz) ? n : rowof[z];
  for (int r = lo + 1; r <= hi; ++r) rowptr[r] = z;
}

static int nbits(uint64_t v) { int b = 1; while (b < 64 && (v >> b)) ++b; return b; }

static void coo_to_csr(DBuf<unsigned long long>& keys, DBuf<double>& v, size_t m, int n_rows, int n_cols, fs_csr& out) {
  cudaStream_t st = stream();
  FS_REQUIRE(m < ((size_t)1 << 31), "sparse product too large");
  DBuf<unsigned long long> keys_alt(m), ukeys(m);
  DBuf<double> v_alt(m), uv(m);
  DBuf<int> nruns(1);
  cub::DoubleBuffer<unsigned long long> kb(keys.p, keys_alt.p);
  cub::DoubleBuffer<double> vb(v.p, v_alt.p);
  size_t bytes = 0;
  const int end_bit = 32 + nbits((uint64_t)std::max(n_rows - 1, 1));
  FS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kb, vb, (int)m, 0, end_bit, st));
  DBuf<char> tmp(bytes);
  FS_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, kb, vb, (int)m, 0, end_bit, st));
  size_t bytes2 = 0;
  FS_CUDA(cub::DeviceReduce::ReduceByKey(nullptr, bytes2, kb.Current(), ukeys.p, vb.Current(), uv.p, nruns.p, cub::Sum(), (int)m, st));
  DBuf<char> tmp2(bytes2);
  FS_CUDA(cub::DeviceReduce::ReduceByKey(tmp2.p, bytes2, kb.Current(), ukeys.p, vb.Current(), uv.p, nruns.p, cub::Sum(), (int)m, st));
  count_launch(10);
  const int nnzc = nruns.to_host()[0];
  out.n = n_rows;
  out.nnz = nnzc;
  out.rowptr_own.alloc(n_rows + 1);
  out.colidx_own.alloc(nnzc);
  out.vals.alloc(nnzc);
  DBuf<int> rowof(nnzc);
  k_split_keys<<<div_up(nnzc, 256), 256, 0, st>>>(ukeys.p, nnzc, rowof.p, out.colidx_own.p);
  FS_LAUNCH_CHECK();
  k_rowptr_from_rowof<<<div_up(nnzc + 1, 256), 256, 0, st>>>(rowof.p, nnzc, n_rows, out.rowptr_own.p);
  FS_LAUNCH_CHECK();
  FS_CUDA(cudaMemcpyAsync(out.vals.p, uv.p, (size_t)nnzc * sizeof(double), cudaMemcpyDeviceToDevice, st));
  out.rowptr = out.rowptr_own.p;
  out.colidx = out.colidx_own.p;
  (void)n_cols;
  FS_CUDA(cudaStreamSynchronize(st));
}

// plain-aggregation Galerkin product (used between the two pairwise passes)
__global__ void k_coarse_keys(CsrView A, const int* __restrict__ agg, unsigned long long* __restrict__ keys,
                              double* __restrict__ v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  const unsigned long long ri = (unsigned long long)(unsigned)agg[i] << 32;
  for (int k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) {
    keys[k] = ri | (unsigned)agg[A.colidx[k]];
    v[k] = A.vals[k];
  }
}

static void galerkin(const CsrView& A, const int* agg, int nc, fs_csr& out) {
  const size_t m = (size_t)A.nnz;
  DBuf<unsigned long long> keys(m);
  DBuf<double> v(m);
  k_coarse_keys<<<div_up(A.n, 256), 256, 0, stream()>>>(A, agg, keys.p, v.p);
  FS_LAUNCH_CHECK();
  coo_to_csr(keys, v, m, nc, nc, out);
}

// smoothed prolongator P = (I - w D^-1 A) P_tent,  P_tent(i, agg[i]) = 1
// Filtered smoothing (theta > 0): couplings weaker than theta x the row's strongest one are
// lumped onto the diagonal before P is smoothed, so P only spreads along strong couplings and
// the Galerkin operator stays sparse on anisotropic meshes.  theta = 0: plain smoothing.
__global__ void k_prolongator_coo(CsrView A, const int* __restrict__ agg, const double* __restrict__ dinv, double w,
                                  double theta, unsigned long long* __restrict__ keys, double* __restrict__ v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  const unsigned long long ri = (unsigned long long)(unsigned)i << 32;
  double wmax = 0.0, diag = 0.0;
  for (int k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) {
    if (A.colidx[k] == i) diag = A.vals[k];
    else wmax = fmax(wmax, -A.vals[k]);
  }
  const double thr = theta * wmax;
  double dF = diag;
  if (theta > 0.0)
    for (int k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k)
      if (A.colidx[k] != i && -A.vals[k] < thr) dF += A.vals[k];
  const double s = (theta > 0.0) ? ((dF != 0.0) ? -w / dF : 0.0) : -w * dinv[i];
  const unsigned self = (unsigned)agg[i];
  for (int k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) {
    const int j = A.colidx[k];
    double val;
    if (j == i) val = s * ((theta > 0.0) ? dF : A.vals[k]);
    else val = (theta > 0.0 && -A.vals[k] < thr) ? 0.0 : s * A.vals[k];
    // dropped (weak) entries are parked on the row's own aggregate with value 0: they vanish in the reduce
    keys[k + i] = ri | ((val == 0.0 && j != i) ? self : (unsigned)agg[j]);
    v[k + i] = val;
  }
  const int e = A.rowptr[i + 1] + i;          // one extra slot per row for the tentative entry
  keys[e] = ri | self;
  v[e] = 1.0;
}

__global__ void k_transpose_coo(CsrView P, unsigned long long* __restrict__ keys, double* __restrict__ v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n) return;
  for (int k = P.rowptr[i]; k < P.rowptr[i + 1]; ++k) {
    keys[k] = ((unsigned long long)(unsigned)P.colidx[k] << 32) | (unsigned)i;
    v[k] = P.vals[k];
  }
}

// Folded cycle operators from A, Q = A P and P (all n rows):
//   U  = [G | P~]  with G_ij = 2 w d_i delta_ij - w^2 (d_i d_j) A_ij,  P~ = P - w D^-1 Q  (columns shifted by n)
//   R~ = P~^T
// as COO lists (duplicates are summed by coo_to_csr).
__global__ void k_fold_coo(CsrView A, CsrView Q, CsrView P, const double* __restrict__ dinv, double w,
                           unsigned long long* __restrict__ ukeys, double* __restrict__ uv,
                           unsigned long long* __restrict__ rkeys, double* __restrict__ rv) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  const double di = dinv[i];
  const unsigned long long ri = (unsigned long long)(unsigned)i << 32;
  for (int k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) {
    const int j = A.colidx[k];
    double g = -(w * w) * ((di * dinv[j]) * A.vals[k]);
    if (j == i) g += 2.0 * w * di;
    ukeys[k] = ri | (unsigned)j;
    uv[k] = g;
  }
  const size_t oq = (size_t)A.nnz, op = (size_t)A.nnz + (size_t)Q.nnz;
  for (int q = Q.rowptr[i]; q < Q.rowptr[i + 1]; ++q) {
    const int c = Q.colidx[q];
    const double val = -w * di * Q.vals[q];
    ukeys[oq + q] = ri | (unsigned)(A.n + c);
    uv[oq + q] = val;
    rkeys[q] = ((unsigned long long)(unsigned)c << 32) | (unsigned)i;
    rv[q] = val;
  }
  for (int q = P.rowptr[i]; q < P.rowptr[i + 1]; ++q) {
    const int c = P.colidx[q];
    const double val = P.vals[q];
    ukeys[op + q] = ri | (unsigned)(A.n + c);
    uv[op + q] = val;
    rkeys[(size_t)Q.nnz + q] = ((unsigned long long)(unsigned)c << 32) | (unsigned)i;
    rv[(size_t)Q.nnz + q] = val;
  }
}

// C = A * B : per nonzero (i,k) of A the whole row k of B
__global__ void k_spgemm_count(CsrView A, const int* __restrict__ Browptr, int* __restrict__ cnt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  int c = 0;
  for (int k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) { const int j = A.colidx[k]; c += Browptr[j + 1] - Browptr[j]; }
  cnt[i] = c;
}
__global__ void k_spgemm_expand(CsrView A, CsrView B, const long long* __restrict__ off, unsigned long long* __restrict__ keys,
                                double* __restrict__ v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  long long o = off[i];
  const unsigned long long ri = (unsigned long long)(unsigned)i << 32;
  for (int k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) {
    const int j = A.colidx[k];
    const double a = A.vals[k];
    for (int q = B.rowptr[j]; q < B.rowptr[j + 1]; ++q) { keys[o] = ri | (unsigned)B.colidx[q]; v[o] = a * B.vals[q]; ++o; }
  }
}

static void spgemm(const CsrView& A, const CsrView& B, int n_cols, fs_csr& out) {
  cudaStream_t st = stream();
  DBuf<int> cnt(A.n);
  DBuf<long long> off((size_t)A.n + 1);
  k_spgemm_count<<<div_up(A.n, 256), 256, 0, st>>>(A, B.rowptr, cnt.p);
  FS_LAUNCH_CHECK();
  size_t bytes = 0;
  FS_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, cnt.p, off.p, A.n, st));
  DBuf<char> tmp(bytes);
  FS_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, cnt.p, off.p, A.n, st));
  count_launch(2);
  long long last_off = 0;
  int last_cnt = 0;
  FS_CUDA(cudaMemcpyAsync(&last_off, off.p + (A.n - 1), sizeof(long long), cudaMemcpyDeviceToHost, st));
  FS_CUDA(cudaMemcpyAsync(&last_cnt, cnt.p + (A.n - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
  FS_CUDA(cudaStreamSynchronize(st));
  const size_t m = (size_t)(last_off + last_cnt);
  DBuf<unsigned long long> keys(m);
  DBuf<double> v(m);
  k_spgemm_expand<<<div_up(A.n, 256), 256, 0, st>>>(A, B, off.p, keys.p, v.p);
  FS_LAUNCH_CHECK();
  coo_to_csr(keys, v, m, A.n, n_cols, out);
}

// dense copy of the coarsest operator; a Neumann (singular) operator gets the rank-one term
// sigma/n * 1 1^T so that its inverse is the pseudo-inverse on mean-free vectors
__global__ void k_dense_from_csr(CsrView A, double shift, double* __restrict__ M) {
  const int n = A.n;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n * n; e += gridDim.x * blockDim.x) M[e] = shift;
}
__global__ void k_dense_add_csr(CsrView A, double* __restrict__ M) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  for (int k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) M[(size_t)i * A.n + A.colidx[k]] += A.vals[k];
}
__global__ void k_rowsum_max(CsrView A, double* __restrict__ out2) {   // out2 = {max |rowsum|, max diag}
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  double s = 0.0, d = 0.0;
  for (int k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) { s += A.vals[k]; if (A.colidx[k] == i) d = A.vals[k]; }
  atomicMax(reinterpret_cast<unsigned long long*>(out2), (unsigned long long)__double_as_longlong(fabs(s)));
  atomicMax(reinterpret_cast<unsigned long long*>(out2 + 1), (unsigned long long)__double_as_longlong(fabs(d)));
}
// in-place Gauss-Jordan inversion of an SPD n x n matrix (no pivoting needed), one CTA
__global__ void __launch_bounds__(1024) k_dense_invert(int n, double* __restrict__ M) {
  __shared__ double colk[2048];
  __shared__ double piv;
  for (int k = 0; k < n; ++k) {
    if (threadIdx.x == 0) piv = 1.0 / M[(size_t)k * n + k];
    __syncthreads();
    const double ip = piv;
    for (int i = threadIdx.x; i < n; i += blockDim.x) colk[i] = M[(size_t)i * n + k];
    __syncthreads();
    // row k: scale; pivot entry becomes 1/pivot
    for (int j = threadIdx.x; j < n; j += blockDim.x) M[(size_t)k * n + j] = (j == k) ? ip : M[(size_t)k * n + j] * ip;
    __syncthreads();
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
      const int i = e / n, j = e - i * n;
      if (i == k) continue;
      const double f = colk[i];
      M[e] = (j == k) ? -f * ip : M[e] - f * M[(size_t)k * n + j];
    }
    __syncthreads();
  }
}
// The same elimination with the whole GPU: two launches per pivot (save the pivot column and scale the pivot row; update
// all other rows).  2 n small launches (~6 ms at n = 960) instead of one CTA sweeping n^2 entries n times (256 ms).
__global__ void __launch_bounds__(1024) k_gj_pivot(int n, int k, double* __restrict__ M, double* __restrict__ colk) {
  __shared__ double piv;
  if (threadIdx.x == 0) piv = 1.0 / M[(size_t)k * n + k];
  __syncthreads();
  const double ip = piv;
  for (int i = threadIdx.x; i < n; i += blockDim.x) colk[i] = M[(size_t)i * n + k];
  __syncthreads();
  for (int j = threadIdx.x; j < n; j += blockDim.x) M[(size_t)k * n + j] = (j == k) ? ip : M[(size_t)k * n + j] * ip;
}
__global__ void k_gj_update(int n, int k, double* __restrict__ M, const double* __restrict__ colk) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= n || i == k) return;
  const double f = colk[i];
  const double rkj = M[(size_t)k * n + j];            // already scaled; entry (k,k) holds 1/pivot
  M[(size_t)i * n + j] = (j == k) ? -f * rkj : M[(size_t)i * n + j] - f * rkj;
}
static void dense_invert(int n, double* M) {
  cudaStream_t st = stream();
  if (n <= 64) {
    k_dense_invert<<<1, 1024, 0, st>>>(n, M);
    FS_LAUNCH_CHECK();
    return;
  }
  DBuf<double> colk(n);
  const dim3 grid(div_up(n, 256), n);
  for (int k = 0; k < n; ++k) {
    k_gj_pivot<<<1, 1024, 0, st>>>(n, k, M, colk.p);
    k_gj_update<<<grid, 256, 0, st>>>(n, k, M, colk.p);
  }
  FS_CUDA(cudaGetLastError());
  count_launch(2 * n);
  FS_CUDA(cudaStreamSynchronize(st));
}

// x = Minv * b, one warp per row, four independent partial sums per lane (the row is a dependent
// chain of ~n/32 loads otherwise: this kernel is pure latency)
__global__ void k_dense_gemv(int n, const double* __restrict__ Minv, const double* __restrict__ b, double* __restrict__ x) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  pdl_launch();
  if (row >= n) return;
  const double* __restrict__ m = Minv + (size_t)row * n;
  for (int j = lane * 16; j < n; j += 512) prefetch_l2(m + j);     // the row is on its way while the producer of b finishes
  pdl_wait();
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int j = lane;
  for (; j + 96 < n; j += 128) {
    const double m0 = __ldg(m + j), m1 = __ldg(m + j + 32), m2 = __ldg(m + j + 64), m3 = __ldg(m + j + 96);
    s0 += m0 * __ldg(b + j); s1 += m1 * __ldg(b + j + 32); s2 += m2 * __ldg(b + j + 64); s3 += m3 * __ldg(b + j + 96);
  }
  for (; j < n; j += 32) s0 += __ldg(m + j) * __ldg(b + j);
  double s = (s0 + s1) + (s2 + s3);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) x[row] = s;
}

__global__ void k_to_f32(const double* __restrict__ in, int64_t n, float* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)in[i];
}

static double env_num(const char* name, double dflt) {
  const char* e = std::getenv(name);
  return e ? std::atof(e) : dflt;
}

// first-row split of the coarse level from the fine one: coarse ids follow their aggregates' leaders, so
// the aggregates touched by the rows below split[r] are (up to stragglers) the ids below max(agg[0..split[r])) + 1
static std::vector<int64_t> coarse_split(const int* agg, const std::vector<int64_t>& split, int nc) {
  cudaStream_t st = stream();
  const int world = (int)split.size() - 1;
  std::vector<int64_t> cs(world + 1, 0);
  DBuf<int> out(1);
  for (int r = 1; r < world; ++r) {
    const int k = (int)split[r];
    if (k <= 0) { cs[r] = 0; continue; }
    size_t bytes = 0;
    FS_CUDA(cub::DeviceReduce::Max(nullptr, bytes, agg, out.p, k, st));
    DBuf<char> tmp(bytes);
    FS_CUDA(cub::DeviceReduce::Max(tmp.p, bytes, agg, out.p, k, st));
    count_launch();
    cs[r] = std::min<int64_t>(nc, (int64_t)out.to_host()[0] + 1);
  }
  cs[world] = nc;
  for (int r = 1; r <= world; ++r) cs[r] = std::max(cs[r], cs[r - 1]);
  return cs;
}

Amg* amg_setup(fs_csr* fine, const AmgPartSpec* ps) {
  std::unique_ptr<Amg> amg(new Amg());
  if (ps) {
    amg->part.rank = ps->rank;
    amg->part.world = ps->world;
    amg->part.split.push_back(ps->split0);
    FS_REQUIRE((int)ps->split0.size() == ps->world + 1 && ps->split0.back() == fine->n, "bad fine-level split");
  }
  amg->omega = env_num("FS_AMG_OMEGA", amg->omega);
  amg->omega_p = env_num("FS_AMG_OMEGA_P", amg->omega_p);
  amg->theta = env_num("FS_AMG_THETA", amg->theta);
  amg->coarse_sweeps = (int)env_num("FS_AMG_COARSE_SWEEPS", amg->coarse_sweeps);
  const int min_rows = (int)env_num("FS_AMG_MIN_ROWS", 2048);   // coarsest level: dense inverse
  const size_t max_levels = (size_t)env_num("FS_AMG_MAX_LEVELS", 16);
  const int passes0 = std::max(1, (int)env_num("FS_AMG_PASSES0", 2));       // pairwise passes on the finest level
  const int passes_rest = std::max(1, (int)env_num("FS_AMG_PASSES", 2));    // ... and on the coarser ones
  amg->folded = env_num("FS_AMG_FOLD", 1) != 0;
  amg->sell = env_num("FS_AMG_SELL", 1) != 0;
  amg->sub_rows = (int)env_num("FS_AMG_SUB_ROWS", 100000);
  const bool fp32 = env_num("FS_AMG_FP32", 1) != 0;
  ensure_tiles(fine);
  jacobi_prepare(fine);
  {
    std::unique_ptr<AmgLevel> l0(new AmgLevel());
    l0->Aref = fine;
    l0->n = (int)fine->n;
    amg->L.push_back(std::move(l0));
  }
  while (amg->L.back()->n > min_rows && amg->L.size() < max_levels) {
    AmgLevel& cur = *amg->L.back();
    const fs_csr& A = cur.mat();
    const CsrView Av = A.view();
    cudaStream_t st = stream();
    // aggregates from repeated pairwise passes along the strongest couplings (each pass roughly
    // halves the row count; the plain Galerkin product of a pass is the next pass's graph)
    const int passes = (amg->L.size() == 1) ? passes0 : passes_rest;
    DBuf<int> agg(cur.n);
    int nc = 0;
    {
      DBuf<int> a1;
      nc = pairwise(Av, a1);
      FS_CUDA(cudaMemcpyAsync(agg.p, a1.p, cur.n * sizeof(int), cudaMemcpyDeviceToDevice, st));
      std::unique_ptr<fs_csr> Ak;
      for (int ps = 1; ps < passes; ++ps) {
        std::unique_ptr<fs_csr> An(new fs_csr());
        galerkin(Ak ? Ak->view() : Av, Ak ? a1.p : agg.p, nc, *An);   // operator on the current aggregates
        DBuf<int> a2;
        const int n2 = pairwise(An->view(), a2);
        k_compose<<<div_up(cur.n, 256), 256, 0, st>>>(cur.n, agg.p, a2.p, agg.p);
        FS_LAUNCH_CHECK();
        a1 = std::move(a2);
        Ak = std::move(An);
        nc = n2;
      }
    }
    if (nc >= cur.n * 0.8) break;                      // coarsening stalled
    // partitioned cycle: this level is cut into row blocks iff it is the finest or larger than gather_rows
    const size_t lidx = amg->L.size() - 1;
    const bool part_l = ps && amg->part.split.size() == lidx + 1 && (lidx == 0 || cur.n > ps->gather_rows);
    if (part_l) amg->part.split.push_back(coarse_split(agg.p, amg->part.split[lidx], nc));
    // smoothed prolongator, its transpose, and the Galerkin operator P^T (A P)
    {
      const size_t m = (size_t)Av.nnz + cur.n;
      DBuf<unsigned long long> keys(m);
      DBuf<double> v(m);
      k_prolongator_coo<<<div_up(cur.n, 256), 256, 0, st>>>(Av, agg.p, A.dinv.p, amg->omega_p, amg->theta, keys.p, v.p);
      FS_LAUNCH_CHECK();
      coo_to_csr(keys, v, m, cur.n, nc, cur.P);
    }
    {
      const size_t m = (size_t)cur.P.nnz;
      DBuf<unsigned long long> keys(m);
      DBuf<double> v(m);
      k_transpose_coo<<<div_up(cur.n, 256), 256, 0, st>>>(cur.P.view(), keys.p, v.p);
      FS_LAUNCH_CHECK();
      coo_to_csr(keys, v, m, nc, cur.n, cur.PT);
    }
    std::unique_ptr<AmgLevel> nxt(new AmgLevel());
    {
      fs_csr Q;
      spgemm(Av, cur.P.view(), nc, Q);
      spgemm(cur.PT.view(), Q.view(), nc, nxt->A);
      if (amg->folded) {
        const size_t mu = (size_t)Av.nnz + (size_t)Q.nnz + (size_t)cur.P.nnz, mr = (size_t)Q.nnz + (size_t)cur.P.nnz;
        DBuf<unsigned long long> ukeys(mu), rkeys(mr);
        DBuf<double> uv(mu), rv(mr);
        k_fold_coo<<<div_up(cur.n, 256), 256, 0, st>>>(Av, Q.view(), cur.P.view(), A.dinv.p, amg->omega, ukeys.p, uv.p, rkeys.p, rv.p);
        FS_LAUNCH_CHECK();
        coo_to_csr(ukeys, uv, mu, cur.n, cur.n + nc, cur.U);
        coo_to_csr(rkeys, rv, mr, nc, cur.n, cur.Rt);
        if (part_l) {
          // the global CSR forms stay until amg_part_finalize has cut this rank's rows out of them
          FS_REQUIRE(amg->sell, "the partitioned cycle needs the SELL layout (FS_AMG_SELL=1)");
        } else {
          if (amg->sell) {
            // SELL-C-sigma (rows sorted by length inside windows of sigma rows) removes the padding -- 35 % for R~, 10 % for
            // [G | P~] -- but was measured slower or equal on the B200 (V-cycle 153.3 -> 155.6 us with sigma = 1024 on R~,
            // 162.5 us with U sorted too, profiles/r02_ab_l2hint_sigma.txt): these kernels are bound by the latency of the
            // dependent stream -> gather chain, not by bytes, and sorting costs gather locality.  Off by default.
            static const int sigma = (int)env_num("FS_SELL_SIGMA", 0);
            static const int sigma_u = (int)env_num("FS_SELL_SIGMA_U", 0);
            sell_build(cur.U, fp32, cur.Us, cur.n, sigma_u);
            sell_build(cur.Rt, fp32, cur.Rts, -1, sigma);
          }
          auto drop = [](fs_csr& M) { M.vals.release(); M.colidx_own.release(); M.rowptr_own.release(); M.rowptr = M.colidx = nullptr; };
          if (amg->sell && cur.n > amg->sub_rows) drop(cur.U); else ensure_tiles(&cur.U);
          if (amg->sell && nc > amg->sub_rows) drop(cur.Rt); else ensure_tiles(&cur.Rt);
        }
      }
    }
    if (amg->folded) {   // the folded cycle does not touch P / P^T again
      cur.P.vals.release(); cur.P.colidx_own.release(); cur.P.rowptr_own.release(); cur.P.nnz = 0;
      cur.PT.vals.release(); cur.PT.colidx_own.release(); cur.PT.rowptr_own.release(); cur.PT.nnz = 0;
    } else {
      ensure_tiles(&cur.P);
      ensure_tiles(&cur.PT);
    }
    nxt->n = nc;
    ensure_tiles(&nxt->A);
    jacobi_prepare(&nxt->A);
    amg->L.push_back(std::move(nxt));
  }
  if (ps) {
    FS_REQUIRE(amg->folded, "the partitioned cycle needs the folded form (FS_AMG_FOLD=1)");
    amg->part.Lp = (int)amg->part.split.size() - 1;
    FS_REQUIRE(amg->part.Lp >= 1 && amg->part.Lp < (int)amg->L.size(),
               "partitioned AMG: the coarsest level must be replicated (raise gather_rows or refine the mesh)");
    amg->part_f32 = fp32;
  }
  if (fp32) {
    // mixed precision: the cycle streams fp32 copies of all its matrices (vectors stay fp64)
    auto to32 = [&](const fs_csr& M) {
      fs_csr& W = const_cast<fs_csr&>(M);
      if (W.vals32.n == (size_t)W.nnz || W.nnz == 0 || !W.vals.p) return;
      W.vals32.alloc(W.nnz);
      k_to_f32<<<div_up(W.nnz, 256), 256, 0, stream()>>>(W.vals.p, W.nnz, W.vals32.p);
      FS_LAUNCH_CHECK();
    };
    for (size_t li = 0; li < amg->L.size(); ++li) {
      auto& l = amg->L[li];
      if (ps && (int)li < amg->part.Lp) continue;   // fp64 CSR kept for the row-block extraction
      if (!amg->folded) to32(l->mat());
      if (l->P.nnz) { to32(l->P); to32(l->PT); }
      if (l->U.nnz) {   // only the fp32 copies of the folded operators are ever read
        to32(l->U); to32(l->Rt);
        FS_CUDA(cudaStreamSynchronize(stream()));
        if (l->U.vals32.n) l->U.vals.release();
        if (l->Rt.vals32.n) l->Rt.vals.release();
      }
    }
  }
  for (size_t l = 0; l < amg->L.size(); ++l) {
    AmgLevel& lv = *amg->L[l];
    if (!amg->folded || l + 1 == amg->L.size()) { lv.r.alloc(lv.n); lv.t.alloc(lv.n); }
    if (l > 0) { lv.x.alloc(lv.n); lv.b.alloc(lv.n); }
  }
  {
    AmgLevel& last = *amg->L.back();
    if (amg->L.size() > 1 && last.n <= 2048 && env_num("FS_AMG_DENSE_COARSE", 1) != 0) {
      const CsrView Ac = last.mat().view();
      const int n = last.n;
      DBuf<double> stat(2);
      stat.zero();
      k_rowsum_max<<<div_up(n, 256), 256, 0, stream()>>>(Ac, stat.p);
      FS_LAUNCH_CHECK();
      std::vector<double> hs = stat.to_host();
      const bool singular = hs[0] <= 1e-9 * hs[1];
      amg->coarse_inv.alloc((size_t)n * n);
      k_dense_from_csr<<<div_up(n * n, 256), 256, 0, stream()>>>(Ac, singular ? hs[1] / n : 0.0, amg->coarse_inv.p);
      FS_LAUNCH_CHECK();
      k_dense_add_csr<<<div_up(n, 256), 256, 0, stream()>>>(Ac, amg->coarse_inv.p);
      FS_LAUNCH_CHECK();
      dense_invert(n, amg->coarse_inv.p);
      amg->coarse_n = n;
    }
  }
  FS_CUDA(cudaStreamSynchronize(stream()));
  if (std::getenv("FS_AMG_VERBOSE")) {
    std::fprintf(stderr, "[amg] levels:");
    for (auto& l : amg->L)
      std::fprintf(stderr, " %d(nnz %lld, U %lld [sell %lld%s, %lld slices unpacked], Rt %lld [sell %lld%s, %lld slices unpacked])", l->n,
                   (long long)l->mat().nnz, (long long)l->U.nnz, l->Us.padded, l->Us.pk.p ? " packed" : "", l->Us.pk_unpacked,
                   (long long)l->Rt.nnz, l->Rts.padded, l->Rts.pk.p ? " packed" : "", l->Rts.pk_unpacked);
    std::fprintf(stderr, "  folded %d\n", (int)amg->folded);
  }
  return amg.release();
}

// ---------------------------------------------------------------- cycle kernels
// x = w * dinv * b
__global__ void k_jac0(int n, double w, const double* __restrict__ dinv, const double* __restrict__ b, double* __restrict__ x) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[i] = w * dinv[i] * b[i];
}
// r = b - Ax (Ax given)
__global__ void k_resid(int n, const double* __restrict__ b, const double* Ax, double* r /* may be Ax */) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) r[i] = b[i] - Ax[i];
}
// x += w * dinv * (b - Ax)
__global__ void k_jac_update(int n, double w, const double* __restrict__ dinv, const double* __restrict__ b,
                             const double* __restrict__ Ax, double* __restrict__ x) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[i] += w * dinv[i] * (b[i] - Ax[i]);
}
// x += y
__global__ void k_add(int n, const double* __restrict__ y, double* __restrict__ x) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[i] += y[i];
}
// coarsest level: `sweeps` damped-Jacobi sweeps from x = 0 inside one CTA (n <= 1024)
__global__ void __launch_bounds__(1024)
k_coarse_jacobi(CsrView A, const double* __restrict__ dinv, const double* __restrict__ b, double* __restrict__ x, double w, int sweeps) {
  __shared__ double xs[1024];
  const int i = threadIdx.x;
  double xi = (i < A.n) ? w * dinv[i] * b[i] : 0.0;
  for (int s = 1; s < sweeps; ++s) {
    xs[i] = xi;
    __syncthreads();
    if (i < A.n) {
      double ax = 0.0;
      for (int k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) ax += A.vals[k] * xs[A.colidx[k]];
      xi += w * dinv[i] * (b[i] - ax);
    }
    __syncthreads();
  }
  if (i < A.n) x[i] = xi;
}

static int vgrid(int n) { return std::max(1, std::min(div_up(n, 256), sm_count() * 8)); }

static void coarse_solve(Amg& amg, size_t l, const double* b, double* x) {
  cudaStream_t st = stream();
  AmgLevel& lv = *amg.L[l];
  const fs_csr& A = lv.mat();
  const CsrView Av = A.view();
  const int n = lv.n, g = vgrid(n);
  const double w = amg.omega;
  if (amg.coarse_n == n && l > 0) {
    launch_pdl(k_dense_gemv, div_up(n * 32, 128), 128, 0, n, amg.coarse_inv.p, b, x);
    FS_LAUNCH_CHECK();
  } else if (n <= 1024) {
    k_coarse_jacobi<<<1, 1024, 0, st>>>(Av, A.dinv.p, b, x, w, amg.coarse_sweeps);
    FS_LAUNCH_CHECK();
  } else {
    k_jac0<<<g, 256, 0, st>>>(n, w, A.dinv.p, b, x);
    FS_LAUNCH_CHECK();
    for (int s = 1; s < amg.coarse_sweeps; ++s) {
      spmv_dev(Av, x, lv.r.p);
      k_jac_update<<<g, 256, 0, st>>>(n, w, A.dinv.p, b, lv.r.p, x);
      FS_LAUNCH_CHECK();
    }
  }
}

// The folded cycle: b_c = R~ b ; recurse ; x = [G | P~] [b; x_c].  With dot_part the up-sweep of
// this level also leaves the per-CTA partials of b.x (the CG's r.z); returns their count.
// b32: fp32 mirror of b (null: none).  x32: where to leave an fp32 mirror of x (null: not wanted); *x32_done says whether
// this level's up-sweep kernel wrote it.
static int vcycle_folded(Amg& amg, size_t l, const double* b, double* x, double* dot_part, const float* b32 = nullptr,
                         float* x32 = nullptr, bool* x32_done = nullptr) {
  if (x32_done) *x32_done = false;
  if (l + 1 == amg.L.size()) {
    coarse_solve(amg, l, b, x);
    return 0;
  }
  AmgLevel& lv = *amg.L[l];
  AmgLevel& nx = *amg.L[l + 1];
  const int sub_rows = amg.sub_rows;
  // down: b_c = R~ b
  const float* nb32 = nullptr;
  if (nx.n <= sub_rows && lv.Rt.rowptr) {
    spmv_sub(lv.Rt.view32(), b, nx.b.p, nullptr, 0, nx.b32.p);
    nb32 = nx.b32.p;
  } else if (lv.Rts.nslices) {
    SellF32 f;
    f.xf = b32;
    f.yf = nx.b32.p;
    spmv_sell(lv.Rts, b, nx.b.p, nullptr, nullptr, &f);
    nb32 = nx.b32.p;
  } else {
    const CsrView Rt = lv.Rt.view32();
    if (!spmv_warp(Rt, EPI_AX, b, nx.b.p, nullptr, nullptr, 0.0, nullptr, nullptr)) spmv_sub(Rt, b, nx.b.p, nullptr, 0);
  }
  bool nx32 = false;
  vcycle_folded(amg, l + 1, nx.b.p, nx.x.p, nullptr, nb32, nx.x32.p, &nx32);
  // up: x = [G | P~] [b; x_c]
  int g = 0;
  if (lv.n <= sub_rows && lv.U.rowptr) {
    spmv_sub(lv.U.view32(), b, x, nx.x.p, lv.n, x32);
    if (x32_done) *x32_done = x32 != nullptr;
  } else if (lv.Us.nslices) {
    if (l == 0 && amg.top_ev) cudaEventRecord(amg.top_ev[0], stream());
    SellF32 f;
    if (b32 && nx32) { f.xf = b32; f.x2f = nx.x32.p; }      // both gather sources have their mirror: fp32 gathers
    f.yf = x32;
    g = spmv_sell(lv.Us, b, x, nx.x.p, dot_part, &f);
    if (x32_done) *x32_done = x32 != nullptr;
    if (l == 0 && amg.top_ev) cudaEventRecord(amg.top_ev[1], stream());
  }
  else {
    const CsrView U = lv.U.view32();
    g = spmv_warp(U, EPI_AXS, b, x, nullptr, nullptr, 0.0, nullptr, dot_part, nx.x.p, lv.n);
    if (!g) spmv_sub(U, b, x, nx.x.p, lv.n);
  }
  return dot_part ? g : 0;
}

// x0_ready: lv.t already holds the pre-smoothed iterate w D^-1 b (written by the producer of b:
// the CG update kernel on level 0, the restriction's epilogue below it)
static void vcycle_level(Amg& amg, size_t l, const double* b, double* x, bool x0_ready) {
  cudaStream_t st = stream();
  AmgLevel& lv = *amg.L[l];
  const fs_csr& A = lv.mat();
  const CsrView Av = A.view();
  const int n = lv.n, g = vgrid(n);
  const double w = amg.omega;
  if (l + 1 == amg.L.size()) {
    coarse_solve(amg, l, b, x);
    return;
  }
  AmgLevel& nx = *amg.L[l + 1];
  double* xt = lv.t.p;                                                   // iterate before the post-smoothing
  // pre-smooth from a zero guess and residual: xt = w D^-1 b, r = b - A xt
  const CsrView Av32 = A.view32();
  if (x0_ready && spmv_warp(Av32, EPI_RESID, xt, lv.r.p, b, nullptr, 0.0, nullptr, nullptr)) {
    // xt came with b: one gather per nonzero instead of two
  } else if (!spmv_warp(Av32, EPI_PRESM, nullptr, lv.r.p, b, A.dinv.p, w, xt, nullptr)) {
    k_jac0<<<g, 256, 0, st>>>(n, w, A.dinv.p, b, xt);
    FS_LAUNCH_CHECK();
    spmv_dev(Av, xt, lv.r.p);
    k_resid<<<g, 256, 0, st>>>(n, b, lv.r.p, lv.r.p);
    FS_LAUNCH_CHECK();
  }
  // restrict; unless the next level is the dense one, the epilogue also emits its pre-smoothed iterate
  const bool nx_last = (l + 2 == amg.L.size());
  bool nx_ready = false;
  if (!nx_last && spmv_warp(lv.PT.view32(), EPI_AX2, lv.r.p, nx.b.p, nullptr, nx.mat().dinv.p, w, nx.t.p, nullptr)) nx_ready = true;
  else if (!spmv_warp(lv.PT.view32(), EPI_AX, lv.r.p, nx.b.p, nullptr, nullptr, 0.0, nullptr, nullptr))
    spmv_dev(lv.PT.view(), lv.r.p, nx.b.p);
  vcycle_level(amg, l + 1, nx.b.p, nx.x.p, nx_ready);
  if (!spmv_warp(lv.P.view32(), EPI_ADD, nx.x.p, xt, nullptr, nullptr, 0.0, nullptr, nullptr)) {       // xt += P x_c
    spmv_dev(lv.P.view(), nx.x.p, lv.r.p);
    k_add<<<g, 256, 0, st>>>(n, lv.r.p, xt);
    FS_LAUNCH_CHECK();
  }
  // post-smooth into the caller's buffer: x = xt + w D^-1 (b - A xt)
  if (!spmv_warp(Av32, EPI_JACOBI, xt, x, b, A.dinv.p, w, nullptr, nullptr)) {
    spmv_dev(Av, xt, lv.r.p);
    FS_CUDA(cudaMemcpyAsync(x, xt, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    k_jac_update<<<g, 256, 0, st>>>(n, w, A.dinv.p, b, lv.r.p, x);
    FS_LAUNCH_CHECK();
  }
}

void amg_presmooth_target(Amg* amg, double** x0, const double** dinv, double* omega) {
  AmgLevel& l0 = *amg->L[0];
  *x0 = (amg->L.size() > 1 && !amg->folded) ? l0.t.p : nullptr;
  *dinv = l0.mat().dinv.p;
  *omega = amg->omega;
}

double amg_top_bytes(const Amg* amg) {
  if (!amg->folded || amg->L.size() < 2 || !amg->L[0]->Us.nslices) return 0.0;
  const AmgLevel& l0 = *amg->L[0];
  // bytes per true nonzero: fp64 / fp32 value + 32-bit column, or one packed 32-bit word (fp16 value | 16-bit column offset)
  const double entry_bytes = l0.Us.pk.n ? 4.0 : (l0.Us.v32.n ? 8.0 : 12.0);
  // true nonzeros, slice pointers (+ packed: base columns), b gathered + re-read for the dot, x_c gathered, y written
  return (double)l0.Us.nnz * entry_bytes + (l0.Us.pk.n ? 16.0 : 8.0) * (l0.Us.nslices + 1) + 8.0 * l0.n + 8.0 * amg->L[1]->n + 8.0 * l0.n;
}

// Algorithmic bytes of one application of the folded cycle: every operator entry once (packed word 4 B, fp32 value +
// column 8 B, fp64 value + column 12 B; CSR levels + 4 B per row), the dense coarse inverse, and per operator its gather
// sources and its result once (fp32 where a mirror is gathered, + the mirror writes, + r re-read for the fused dot).
double amg_cycle_bytes(const Amg* amg, bool mirrors) {
  if (!amg->folded || amg->L.size() < 2) return 0.0;
  auto entry = [](const fs_sell& S) { return S.pk.n ? 4.0 : (S.v32.n ? 8.0 : 12.0); };
  double bytes = 0.0;
  const size_t nl = amg->L.size();
  for (size_t l = 0; l + 1 < nl; ++l) {
    const AmgLevel& lv = *amg->L[l];
    const AmgLevel& nx = *amg->L[l + 1];
    const bool rt_sell = !(nx.n <= amg->sub_rows && lv.Rt.rowptr) && lv.Rts.nslices;
    const bool u_sell = !(lv.n <= amg->sub_rows && lv.U.rowptr) && lv.Us.nslices;
    const double g_in = (mirrors && rt_sell && lv.Rts.pk.n) ? 4.0 : 8.0;       // gather width of b_l in the restriction
    bytes += rt_sell ? (double)lv.Rts.nnz * entry(lv.Rts) + 16.0 * lv.Rts.nslices : (double)lv.Rt.nnz * 8.0 + 4.0 * nx.n;
    bytes += g_in * lv.n + (mirrors ? 12.0 : 8.0) * nx.n;
    const double g_up = (mirrors && u_sell && lv.Us.pk.n) ? 4.0 : 8.0;
    bytes += u_sell ? (double)lv.Us.nnz * entry(lv.Us) + 16.0 * lv.Us.nslices : (double)lv.U.nnz * 8.0 + 4.0 * lv.n;
    bytes += g_up * (lv.n + nx.n) + ((mirrors && l > 0) ? 12.0 : 8.0) * lv.n + (l == 0 ? 8.0 * lv.n : 0.0);
  }
  const double nc = (double)amg->L.back()->n;
  bytes += amg->coarse_n ? nc * nc * 8.0 + 16.0 * nc : 0.0;
  return bytes;
}

int amg_apply(Amg* amg, const double* r, double* z, bool x0_ready, double* rz_part, cudaEvent_t* top_ev, const float* r32) {
  static const bool use_graph = env_num("FS_AMG_GRAPH", 1) != 0;
  cudaStream_t st = stream();
  ++amg->applications;
  amg->r32 = r32;
  static const bool mirror_all = env_num("FS_AMG_MIRROR_ALL", 1) != 0;   // 0: only the finest level gathers in fp32
  if (r32 && amg->folded)     // mirrors of the level vectors (a few MB), written by the kernels that produce them
    for (size_t l = 1; l < (mirror_all ? amg->L.size() : std::min<size_t>(2, amg->L.size())); ++l)
      if (!amg->L[l]->x32.p) { amg->L[l]->x32.alloc((size_t)amg->L[l]->n); amg->L[l]->b32.alloc((size_t)amg->L[l]->n); }
  const bool fold = amg->folded && amg->L.size() > 1;
  auto cycle = [&]() -> int {
    if (fold) return vcycle_folded(*amg, 0, r, z, rz_part, amg->r32);
    vcycle_level(*amg, 0, r, z, x0_ready);
    return 0;
  };
  if (top_ev && fold) {   // sampled timing: eager, with the event pair around the finest up-sweep
    amg->top_ev = top_ev;
    const int np = vcycle_folded(*amg, 0, r, z, rz_part, amg->r32);
    amg->top_ev = nullptr;
    return np;
  }
  if (use_graph && amg->graph && amg->graph_r == r && amg->graph_z == z && amg->graph_x0 == x0_ready &&
      amg->graph_part == rz_part && amg->graph_r32 == r32) {
    FS_CUDA(cudaGraphLaunch(amg->graph, st));
    count_launch();
    return amg->graph_nparts;
  }
  if (use_graph && amg->applications >= 2 && !amg->graph_failed) {
    // second application with these buffers: capture the cycle (all kernels go to `st`)
    if (amg->graph) { cudaGraphExecDestroy(amg->graph); amg->graph = nullptr; }
    cudaGraph_t g = nullptr;
    FS_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    bool ok = true;
    int nparts = 0;
    try { nparts = cycle(); } catch (...) { ok = false; }
    cudaError_t e = cudaStreamEndCapture(st, &g);
    if (ok && e == cudaSuccess && g) {
      cudaGraphExec_t ex = nullptr;
      if (cudaGraphInstantiate(&ex, g, 0) == cudaSuccess) {
        amg->graph = ex; amg->graph_r = r; amg->graph_z = z; amg->graph_x0 = x0_ready; amg->graph_r32 = r32;
        amg->graph_part = rz_part; amg->graph_nparts = nparts;
        if (std::getenv("FS_AMG_VERBOSE")) std::fprintf(stderr, "[amg] V-cycle captured as a graph\n");
        cudaGraphDestroy(g);
        FS_CUDA(cudaGraphLaunch(amg->graph, st));
        count_launch();
        return nparts;
      }
    }
    if (g) cudaGraphDestroy(g);
    cudaGetLastError();
    if (std::getenv("FS_AMG_VERBOSE")) std::fprintf(stderr, "[amg] graph capture failed (%s)\n", cudaGetErrorString(e));
    amg->graph_failed = true;   // run the cycle eagerly from now on
  }
  return cycle();
}

// ---------------------------------------------------------------- partitioned cycle
int amg_part_levels(const Amg* amg) { return amg->part.Lp; }
const std::vector<int64_t>& amg_part_split(const Amg* amg, int level) { return amg->part.split.at(level); }

static Space make_gather_space(const std::vector<int64_t>& split, int rank, int world) {
  Space sp;
  sp.split = split;
  sp.gather = true;
  sp.own_lo = split[rank];
  sp.n_own = split[rank + 1] - split[rank];
  sp.n_halo = 0;
  sp.cap = (split[world] + 31) / 32 * 32;
  for (int q = 0; q < world; ++q) {
    if (q == rank) continue;
    if (sp.n_own > 0) sp.to[sp.n_to++] = (signed char)q;                       // my block goes to everybody
    if (split[q + 1] > split[q]) sp.from[sp.n_from++] = (signed char)q;        // and I wait for every non-empty block
  }
  return sp;
}

size_t amg_part_collect(Amg* amg, HaloCollector& c0) {
  Amg::Part& P = amg->part;
  const int Lp = P.Lp;
  FS_REQUIRE(Lp >= 1, "amg_part_collect: the hierarchy was not set up for a partitioned cycle");
  std::vector<HaloCollector> col(Lp);   // col[l] for 1 <= l < Lp
  for (int l = 0; l < Lp; ++l) {
    AmgLevel& lv = *amg->L[l];
    const int n = lv.n;
    const CsrView U = lv.U.view(), Rt = lv.Rt.view();
    HaloCollector& cl = (l == 0) ? c0 : col[l];
    cl.add_matrix(U, P.split[l], P.split[l], 0, n);                             // G part reads b_l
    cl.add_matrix(Rt, P.split[l + 1], P.split[l], 0, n);                        // R~ reads b_l, rows = next level's blocks
    if (l + 1 < Lp) col[l + 1].add_matrix(U, P.split[l], P.split[l + 1], n, (int64_t)n + amg->L[l + 1]->n);   // P~ part reads x_{l+1}
  }
  P.space.clear();
  P.space.resize(Lp + 1);
  size_t bytes = 0;
  for (int l = 1; l < Lp; ++l) {
    P.space[l].reset(new Space());
    P.space[l]->split = P.split[l];
    col[l].finish(*P.space[l], P.rank, P.world);
    bytes += 2 * (((size_t)P.space[l]->cap * sizeof(double) + 255) / 256 * 256);
  }
  P.space[Lp].reset(new Space(make_gather_space(P.split[Lp], P.rank, P.world)));
  bytes += ((size_t)P.space[Lp]->cap * sizeof(double) + 255) / 256 * 256;
  return bytes;
}

void amg_part_finalize(Amg* amg, const Space& space0, DistCtx& ctx) {
  Amg::Part& P = amg->part;
  const int Lp = P.Lp, rank = P.rank;
  P.ctx = &ctx;
  P.space0 = &space0;
  P.vb.assign(Lp + 1, DVec{});
  P.vx.assign(Lp + 1, DVec{});
  for (int l = 1; l < Lp; ++l) { P.vb[l] = ctx.carve(*P.space[l], 1); P.vx[l] = ctx.carve(*P.space[l], 1); }
  P.vb[Lp] = ctx.carve(*P.space[Lp], 1);
  auto drop = [](fs_csr& M) {
    M.vals.release(); M.vals32.release(); M.colidx_own.release(); M.rowptr_own.release();
    M.rowptr = M.colidx = nullptr; M.nnz = 0;
  };
  // Operators with at most sub_rows local rows stay in CSR form for the lanes-per-row kernel (these levels are latency
  // bound: one wave of short chains; their halo exchange runs as separate push / wait kernels), the large ones become
  // SELL-32 with the exchange fused into the kernel (spmv_sell.cu).
  auto keep_csr = [&](fs_csr& dst, fs_csr& src) {
    dst.n = src.n; dst.nnz = src.nnz;
    dst.rowptr_own = std::move(src.rowptr_own); dst.colidx_own = std::move(src.colidx_own);
    dst.rowptr = dst.rowptr_own.p; dst.colidx = dst.colidx_own.p;
    dst.vals32.alloc(std::max<int64_t>(dst.nnz, 1));
    if (dst.nnz) { k_to_f32<<<div_up(dst.nnz, 256), 256, 0, stream()>>>(src.vals.p, dst.nnz, dst.vals32.p); FS_LAUNCH_CHECK(); }
    FS_CUDA(cudaStreamSynchronize(stream()));
    ensure_tiles(&dst);
  };
  for (int l = 0; l < Lp; ++l) {
    AmgLevel& lv = *amg->L[l];
    const Space& sl = (l == 0) ? space0 : *P.space[l];
    const Space* sn = P.space[l + 1].get();           // next level's space (gather space for l + 1 == Lp)
    {
      fs_csr loc;
      extract_rows(lv.U.view(), P.split[l][rank], P.split[l][rank + 1], sl, lv.n, sn, loc);
      // the packed entries are rounded at the GLOBAL operator's scale: same bits as the single-GPU hierarchy
      const double mu = (amg->part_f32 && sell_pack_enabled()) ? csr_maxabs(lv.U) : 0.0;
      drop(lv.U);
      const int nsplit = (int)(sl.n_own + sl.n_halo);
      if (l >= 1 && loc.n <= amg->sub_rows) keep_csr(lv.U, loc);
      else {
        sell_build(loc, amg->part_f32, lv.Us, nsplit, 0, mu);
        sell_mark_boundary(lv.Us, loc, (int)sl.n_own, nsplit, sn->gather ? -1 : (int)sn->n_own, l >= 1 ? &sl : nullptr);
      }
      lv.part_nsplit = nsplit;
    }
    {
      fs_csr loc;
      const int64_t r0 = P.split[l + 1][rank], r1 = P.split[l + 1][rank + 1];
      const bool have = r1 > r0;
      if (have) extract_rows(lv.Rt.view(), r0, r1, sl, lv.n, nullptr, loc);
      const double mr = (amg->part_f32 && sell_pack_enabled()) ? csr_maxabs(lv.Rt) : 0.0;
      drop(lv.Rt);
      if (have) {
        if (loc.n <= amg->sub_rows) keep_csr(lv.Rt, loc);
        else {
          sell_build(loc, amg->part_f32, lv.Rts, -1, 0, mr);
          sell_mark_boundary(lv.Rts, loc, (int)sl.n_own, 0x7fffffff, -1, sn);
        }
      }
    }
  }
  // level 0 borrowed the caller's global fine matrix: nothing in the folded cycle reads it again
  amg->L[0]->Aref = nullptr;
  FS_CUDA(cudaStreamSynchronize(stream()));
  if (std::getenv("FS_AMG_VERBOSE")) {
    std::fprintf(stderr, "[amg rank %d] partitioned levels %d:", rank, Lp);
    for (int l = 0; l < Lp; ++l) {
      const Space& sl = (l == 0) ? space0 : *P.space[l];
      std::fprintf(stderr, " L%d own %lld halo %lld send %d |", l, (long long)sl.n_own, (long long)sl.n_halo, sl.n_send);
    }
    std::fprintf(stderr, " gathered level: %lld rows (own %lld)\n", (long long)P.split[Lp].back(), (long long)P.space[Lp]->n_own);
  }
}

// b_c = R~ b (own block; pushed) ; recurse ; x = [G | P~] [b; x_c] (own rows; pushed for the level above)
static int vcycle_folded_dist(Amg& amg, int l, const DVec& bv, double* x, double* dot_part) {
  Amg::Part& P = amg.part;
  DistCtx& ctx = *P.ctx;
  const int Lp = P.Lp;
  if (l == Lp) {                       // replicated tail: wait for every rank's block of b, then the usual cycle
    ctx.wait(P.vb[l]);
    vcycle_folded(amg, (size_t)l, P.vb[l].p, x, nullptr);
    return 0;
  }
  AmgLevel& lv = *amg.L[l];
  const bool next_part = l + 1 < Lp;
  const DVec& nb = P.vb[l + 1];
  // the restriction stores the rows of b_{l+1} that other ranks read (all of them for the replicated level) into
  // their copies itself; a rank without coarse rows has nothing to compute or to send
  double* y = next_part ? nb.p : nb.p + P.split[l + 1][P.rank];
  if (lv.Rts.nslices) {
    const PushSpec ps = ctx.push_spec(nb);
    spmv_sell_dist(lv.Rts, bv.p, y, nullptr, nullptr, ctx.comm, ctx.wait_of(&bv), &ps, 10 + 2 * l);
  } else if (lv.Rt.rowptr) {          // small block: lanes-per-row CSR kernel with the wait and the push inside
    spmv_sub_dist(lv.Rt.view32(), bv.p, y, nullptr, 0, ctx.comm, ctx.wait_of(&bv), ctx.push_spec(nb));
  }
  double* xc = next_part ? P.vx[l + 1].p : amg.L[l + 1]->x.p;
  vcycle_folded_dist(amg, l + 1, nb, xc, nullptr);
  if (lv.Us.nslices) {
    PushSpec ps;
    if (l >= 1) ps = ctx.push_spec(P.vx[l]);
    const int g = spmv_sell_dist(lv.Us, bv.p, x, xc, dot_part, ctx.comm, ctx.wait_of(&bv, next_part ? &P.vx[l + 1] : nullptr), &ps, 11 + 2 * l);
    return dot_part ? g : 0;
  }
  spmv_sub_dist(lv.U.view32(), bv.p, x, xc, lv.part_nsplit, ctx.comm, ctx.wait_of(&bv, next_part ? &P.vx[l + 1] : nullptr),
                ctx.push_spec(P.vx[l]));
  return 0;
}

int amg_apply_dist(Amg* amg, const DVec& r, double* z, double* rz_part) {
  static const bool use_graph = env_num("FS_AMG_GRAPH", 1) != 0;
  cudaStream_t st = stream();
  ++amg->applications;
  if (use_graph && amg->graph && amg->graph_r == r.p && amg->graph_z == z && amg->graph_part == rz_part) {
    FS_CUDA(cudaGraphLaunch(amg->graph, st));
    count_launch();
    return amg->graph_nparts;
  }
  if (use_graph && amg->applications >= 2 && !amg->graph_failed) {
    if (amg->graph) { cudaGraphExecDestroy(amg->graph); amg->graph = nullptr; }
    cudaGraph_t g = nullptr;
    FS_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    bool ok = true;
    int nparts = 0;
    try { nparts = vcycle_folded_dist(*amg, 0, r, z, rz_part); } catch (...) { ok = false; }
    cudaError_t e = cudaStreamEndCapture(st, &g);
    if (ok && e == cudaSuccess && g) {
      cudaGraphExec_t ex = nullptr;
      if (cudaGraphInstantiate(&ex, g, 0) == cudaSuccess) {
        amg->graph = ex; amg->graph_r = r.p; amg->graph_z = z; amg->graph_x0 = false;
        amg->graph_part = rz_part; amg->graph_nparts = nparts;
        cudaGraphDestroy(g);
        FS_CUDA(cudaGraphLaunch(amg->graph, st));
        count_launch();
        return nparts;
      }
    }
    if (g) cudaGraphDestroy(g);
    cudaGetLastError();
    amg->graph_failed = true;
  }
  return vcycle_folded_dist(*amg, 0, r, z, rz_part);
}

int amg_levels(const Amg* amg, int* sizes, int cap) {
  int k = 0;
  for (auto& l : amg->L) { if (k < cap) sizes[k] = l->n; ++k; }
  return k;
}

}  // namespace fs
