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
//   x += w D^-1 (b - A x), with the smoother / residual fused into the SpMV epilogues and the
//   matrices streamed as fp32 copies (vectors fp64).  Coarsest level: dense (pseudo-)inverse GEMV.
#include <cub/cub.cuh>

#include "internal.cuh"

namespace fs {

struct AmgLevel {
  fs_csr A;                 // operator of this level (level 0 borrows the fine matrix)
  const fs_csr* Aref = nullptr;
  int n = 0;
  fs_csr P, PT;             // smoothed prolongator (n x n_coarse) and its transpose (restriction)
  DBuf<double> x, b, r, t;  // work vectors of this level (level 0 uses caller buffers for b/x)
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
  double* graph_z = nullptr;
  bool graph_x0 = false;
  int applications = 0;
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

__global__ void k_compose(int n, const int* __restrict__ a1, const int* __restrict__ a2, int* __restrict__ out) {
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
// x = Minv * b, one warp per row
__global__ void k_dense_gemv(int n, const double* __restrict__ Minv, const double* __restrict__ b, double* __restrict__ x) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= n) return;
  double s = 0.0;
  for (int j = lane; j < n; j += 32) s += Minv[(size_t)row * n + j] * b[j];
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

Amg* amg_setup(fs_csr* fine) {
  std::unique_ptr<Amg> amg(new Amg());
  amg->omega = env_num("FS_AMG_OMEGA", amg->omega);
  amg->omega_p = env_num("FS_AMG_OMEGA_P", amg->omega_p);
  amg->theta = env_num("FS_AMG_THETA", amg->theta);
  amg->coarse_sweeps = (int)env_num("FS_AMG_COARSE_SWEEPS", amg->coarse_sweeps);
  const int min_rows = (int)env_num("FS_AMG_MIN_ROWS", 2048);   // coarsest level: dense inverse
  const size_t max_levels = (size_t)env_num("FS_AMG_MAX_LEVELS", 16);
  const int passes0 = std::max(1, (int)env_num("FS_AMG_PASSES0", 2));       // pairwise passes on the finest level
  const int passes_rest = std::max(1, (int)env_num("FS_AMG_PASSES", 2));    // ... and on the coarser ones
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
    }
    ensure_tiles(&cur.P);
    ensure_tiles(&cur.PT);
    nxt->n = nc;
    ensure_tiles(&nxt->A);
    jacobi_prepare(&nxt->A);
    amg->L.push_back(std::move(nxt));
  }
  if (env_num("FS_AMG_FP32", 1) != 0) {
    // mixed precision: the cycle streams fp32 copies of all its matrices (vectors stay fp64)
    auto to32 = [&](const fs_csr& M) {
      fs_csr& W = const_cast<fs_csr&>(M);
      if (W.vals32.n == (size_t)W.nnz || W.nnz == 0) return;
      W.vals32.alloc(W.nnz);
      k_to_f32<<<div_up(W.nnz, 256), 256, 0, stream()>>>(W.vals.p, W.nnz, W.vals32.p);
      FS_LAUNCH_CHECK();
    };
    for (auto& l : amg->L) {
      to32(l->mat());
      if (l->P.nnz) { to32(l->P); to32(l->PT); }
    }
  }
  for (size_t l = 0; l < amg->L.size(); ++l) {
    AmgLevel& lv = *amg->L[l];
    lv.r.alloc(lv.n);
    lv.t.alloc(lv.n);
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
      k_dense_invert<<<1, 1024, 0, stream()>>>(n, amg->coarse_inv.p);
      FS_LAUNCH_CHECK();
      amg->coarse_n = n;
    }
  }
  FS_CUDA(cudaStreamSynchronize(stream()));
  if (std::getenv("FS_AMG_VERBOSE")) {
    std::fprintf(stderr, "[amg] levels:");
    for (auto& l : amg->L) std::fprintf(stderr, " %d(nnz %lld)", l->n, (long long)l->mat().nnz);
    std::fprintf(stderr, "\n");
  }
  return amg.release();
}

// ---------------------------------------------------------------- cycle kernels
// x = w * dinv * b
__global__ void k_jac0(int n, double w, const double* __restrict__ dinv, const double* __restrict__ b, double* __restrict__ x) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[i] = w * dinv[i] * b[i];
}
// r = b - Ax (Ax given)
__global__ void k_resid(int n, const double* __restrict__ b, const double* __restrict__ Ax, double* __restrict__ r) {
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
    if (amg.coarse_n == n && l > 0) {
      k_dense_gemv<<<div_up(n * 32, 256), 256, 0, st>>>(n, amg.coarse_inv.p, b, x);
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
  *x0 = (amg->L.size() > 1) ? l0.t.p : nullptr;
  *dinv = l0.mat().dinv.p;
  *omega = amg->omega;
}

void amg_apply(Amg* amg, const double* r, double* z, bool x0_ready) {
  static const bool use_graph = env_num("FS_AMG_GRAPH", 1) != 0;
  cudaStream_t st = stream();
  ++amg->applications;
  if (use_graph && amg->graph && amg->graph_r == r && amg->graph_z == z && amg->graph_x0 == x0_ready) {
    FS_CUDA(cudaGraphLaunch(amg->graph, st));
    count_launch();
    return;
  }
  if (use_graph && amg->applications >= 2) {
    // second application with these buffers: capture the cycle (all kernels go to `st`)
    if (amg->graph) { cudaGraphExecDestroy(amg->graph); amg->graph = nullptr; }
    cudaGraph_t g = nullptr;
    FS_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    bool ok = true;
    try { vcycle_level(*amg, 0, r, z, x0_ready); } catch (...) { ok = false; }
    cudaError_t e = cudaStreamEndCapture(st, &g);
    if (ok && e == cudaSuccess && g) {
      cudaGraphExec_t ex = nullptr;
      if (cudaGraphInstantiate(&ex, g, 0) == cudaSuccess) {
        amg->graph = ex; amg->graph_r = r; amg->graph_z = z; amg->graph_x0 = x0_ready;
        cudaGraphDestroy(g);
        FS_CUDA(cudaGraphLaunch(amg->graph, st));
        count_launch();
        return;
      }
    }
    if (g) cudaGraphDestroy(g);
    cudaGetLastError();
  }
  vcycle_level(*amg, 0, r, z, x0_ready);
}

int amg_levels(const Amg* amg, int* sizes, int cap) {
  int k = 0;
  for (auto& l : amg->L) { if (k < cap) sizes[k] = l->n; ++k; }
  return k;
}

}  // namespace fs
