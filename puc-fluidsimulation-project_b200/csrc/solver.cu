// solver.cu -- fp64 CSR SpMV and the Krylov solvers that replace the reference's
// dense np.linalg.solve (code/StokesColor.py:544-545,555,569; code/heatEq.py:323;
// code/poisson.py:285).
//
// Paths, chosen by size and preconditioner (cg_dev / fs_bicgstab):
//   * n <= 8192: the whole CG (1 or 2 RHS) or BiCGStab solve in one CTA (k_cg_small, k_bicgstab_small)
//   * Jacobi, 1 RHS: the persistent cooperative kernel of cg_persistent.cu
//   * AMG: pcg_amg_impl below (device-side scalars; V-cycle of amg.cu; A*p and the 2-RHS SpMV of spmv_sell.cu)
//   * 2 RHS / fallback (FS_CG_MODE=multi): 3 kernels per iteration with device-side scalars:
//       A: Ap = A p            + partial p.Ap          (matrix stream + p gather)
//       B: x += a p; r -= a Ap + partial r.r, r.z      (z = Dinv r never stored)
//       C: p = z + b p
//     each pass re-reduces the previous pass's per-block partials in a fixed order
//     (deterministic, no atomics); the host polls a "done" flag every chunk.
#include <algorithm>

#include "reduce.cuh"

namespace fs {

// ---- SpMV: LPR lanes per row, R interleaved right-hand sides -------------------
template <int R, int LPR, bool DOT>
__global__ void __launch_bounds__(kBlock)
k_spmv(CsrView A, const double* __restrict__ x, double* __restrict__ y, double* __restrict__ part,
       const int* __restrict__ done) {
  __shared__ double red[R * 32];
  if (done && *done) return;
  const int64_t gt = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int sub = threadIdx.x % LPR;
  const int64_t nrow_threads = ((int64_t)gridDim.x * blockDim.x) / LPR;
  double acc[R];
#pragma unroll
  for (int c = 0; c < R; ++c) acc[c] = 0.0;
  const int64_t niter = (A.n + nrow_threads - 1) / nrow_threads;
  for (int64_t it = 0; it < niter; ++it) {
    const int64_t row = it * nrow_threads + gt / LPR;
    double s[R];
#pragma unroll
    for (int c = 0; c < R; ++c) s[c] = 0.0;
    if (row < A.n) {
      const int rs = __ldg(A.rowptr + row), re = __ldg(A.rowptr + row + 1);
      for (int k = rs + sub; k < re; k += LPR) {
        const double a = __ldg(A.vals + k);
        const int col = __ldg(A.colidx + k);
        double xv[R];
        load_vec_ldg<R>(x, col, xv);
#pragma unroll
        for (int c = 0; c < R; ++c) s[c] += a * xv[c];
      }
    }
#pragma unroll
    for (int c = 0; c < R; ++c)
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) s[c] += __shfl_xor_sync(0xffffffffu, s[c], o);
    if (row < A.n && sub == 0) {
      store_vec<R>(y, row, s);
      if (DOT) {
        double xv[R];
        load_vec_ldg<R>(x, row, xv);
#pragma unroll
        for (int c = 0; c < R; ++c) acc[c] += xv[c] * s[c];
      }
    }
  }
  if (DOT) {
    block_reduce<R>(acc, red);
    if (threadIdx.x == 0)
#pragma unroll
      for (int c = 0; c < R; ++c) part[(size_t)blockIdx.x * R + c] = acc[c];
  }
}

// ---- L2 cache-policy hints --------------------------------------------------------
// The matrix (12 B/nnz) is streamed once per iteration: evict-first, so that the CG
// vectors (5 x 8N bytes, < 126 MB L2) stay resident between passes: evict-last.
__device__ __forceinline__ uint64_t pol_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t pol_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ double ldg_f64_hint(const double* a, uint64_t pol) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(a), "l"(pol));
  return v;
}
__device__ __forceinline__ int ldg_s32_hint(const int* a, uint64_t pol) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(pol));
  return v;
}
__device__ __forceinline__ double ld_f64_keep(const double* a, uint64_t pol) {
  double v;
  asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(a), "l"(pol));
  return v;
}
__device__ __forceinline__ void st_f64_keep(double* a, double v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(a), "d"(v), "l"(pol) : "memory");
}

// ---- tile SpMV (CSR-stream): a CTA owns kTileRows consecutive rows.  Phase 1 streams
// the tile's values/columns with fully coalesced, independent loads (8 per thread in
// flight), gathers x and parks the products in shared memory; phase 2 sums each row
// from shared memory in ascending column order (deterministic) and writes y coalesced.
constexpr int kTileRows = 256;
constexpr int kTileUnroll = 8;

template <bool DOT>
__global__ void __launch_bounds__(kTileRows)
k_spmv_tile(CsrView A, const double* __restrict__ x, double* __restrict__ y, double* __restrict__ part,
            const int* __restrict__ done, int ntiles) {
  extern __shared__ double prod[];
  __shared__ int s_rp[kTileRows + 1];
  __shared__ double red[32];
  if (done && *done) return;
  const uint64_t pf = pol_evict_first(), pl = pol_evict_last();
  const int t = threadIdx.x;
  double acc[1] = {0.0};
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int r0 = tile * kTileRows;
    const int nr = min(kTileRows, A.n - r0);
    if (t <= nr) s_rp[t] = __ldg(A.rowptr + r0 + t);
    if (t == 0 && nr == kTileRows) s_rp[kTileRows] = __ldg(A.rowptr + r0 + kTileRows);
    __syncthreads();
    const int base = s_rp[0];
    const int cnt = s_rp[nr] - base;
    const double* __restrict__ va = A.vals + base;
    const int* __restrict__ ca = A.colidx + base;
    for (int k0 = 0; k0 < cnt; k0 += kTileRows * kTileUnroll) {
      double a[kTileUnroll];
      int c[kTileUnroll];
#pragma unroll
      for (int j = 0; j < kTileUnroll; ++j) {
        const int k = k0 + j * kTileRows + t;
        const bool ok = k < cnt;
        a[j] = ok ? ldg_f64_hint(va + k, pf) : 0.0;
        c[j] = ok ? ldg_s32_hint(ca + k, pf) : -1;
      }
#pragma unroll
      for (int j = 0; j < kTileUnroll; ++j) {
        const int k = k0 + j * kTileRows + t;
        if (c[j] >= 0) prod[k] = a[j] * ld_f64_keep(x + c[j], pl);
      }
    }
    __syncthreads();
    if (t < nr) {
      double s = 0.0;
      const int ke = s_rp[t + 1] - base;
      for (int k = s_rp[t] - base; k < ke; ++k) s += prod[k];
      st_f64_keep(y + r0 + t, s, pl);
      if (DOT) acc[0] += ld_f64_keep(x + r0 + t, pl) * s;
    }
    __syncthreads();
  }
  if (DOT) {
    block_reduce<1>(acc, red);
    if (t == 0) part[blockIdx.x] = acc[0];
  }
}

__global__ void k_tile_nnz_max(const int* __restrict__ rowptr, int n, int rows_per_tile, int ntiles, int* __restrict__ out) {
  int tile = blockIdx.x * blockDim.x + threadIdx.x;
  if (tile >= ntiles) return;
  int r0 = tile * rows_per_tile, r1 = min(n, r0 + rows_per_tile);
  atomicMax(out, rowptr[r1] - rowptr[r0]);
}

void ensure_tiles(fs_csr* a) {
  if (a->tile_nnz_max >= 0) return;
  DBuf<int> mx(2);
  mx.zero();
  const int nt256 = div_up(a->n, kTileRows), nt32 = div_up(a->n, 32);
  k_tile_nnz_max<<<div_up(nt256, 256), 256, 0, stream()>>>(a->rowptr, (int)a->n, kTileRows, nt256, mx.p);
  FS_LAUNCH_CHECK();
  k_tile_nnz_max<<<div_up(nt32, 256), 256, 0, stream()>>>(a->rowptr, (int)a->n, 32, nt32, mx.p + 1);
  FS_LAUNCH_CHECK();
  std::vector<int> h = mx.to_host();
  a->tile_nnz_max = h[0];
  a->wtile_nnz_max = h[1];
}

constexpr int kTileSmemMax = 96 * 1024;   // products buffer cap (bytes); beyond it use the vector kernel

static bool tile_ok(const CsrView& A) { return A.tile_nnz_max > 0 && (size_t)A.tile_nnz_max * 8 <= kTileSmemMax; }

static int tile_grid(const CsrView& A) {
  const int ntiles = div_up(A.n, kTileRows);
  const size_t smem = (size_t)A.tile_nnz_max * 8;
  int per_sm = (int)std::min<size_t>(8, (200 * 1024) / (smem + 2048));
  per_sm = std::max(per_sm, 1);
  return std::max(1, std::min(ntiles, std::min(kMaxBlocks, sm_count() * per_sm)));
}

template <bool DOT>
static void launch_spmv_tile(const CsrView& A, const double* x, double* y, double* part, const int* done, int grid) {
  const size_t smem = (size_t)A.tile_nnz_max * 8;
  // raise the limit on first use whatever the size: the kernel's static shared memory (row
  // pointers, reduction scratch) counts against the default 48 KB too
  static bool attr_set[2] = {false, false};
  if (!attr_set[DOT]) {
    FS_CUDA(cudaFuncSetAttribute(k_spmv_tile<DOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileSmemMax));
    attr_set[DOT] = true;
  }
  k_spmv_tile<DOT><<<grid, kTileRows, smem, stream()>>>(A, x, y, part, done, div_up(A.n, kTileRows));
  FS_LAUNCH_CHECK();
}

static int spmv_grid(int64_t n, int lpr) {
  int64_t want = (n * lpr + kBlock - 1) / kBlock;
  int64_t cap = std::min<int64_t>(kMaxBlocks, (int64_t)sm_count() * 6);
  return (int)std::max<int64_t>(1, std::min(want, cap));
}

static int pick_lpr(const CsrView& A) {
  double avg = A.n ? (double)A.nnz / (double)A.n : 1.0;
  if (avg <= 2.5) return 2;
  if (avg <= 5.0) return 4;
  if (avg <= 12.0) return 8;
  if (avg <= 24.0) return 16;
  return 32;
}

template <int R>
static int spmv_launch_grid(const CsrView& A) {
  if (R == 1 && tile_ok(A)) return tile_grid(A);
  return spmv_grid(A.n, pick_lpr(A));
}

template <int R, bool DOT>
static void launch_spmv(const CsrView& A, const double* x, double* y, double* part, const int* done, int grid_override = 0) {
  if (R == 1 && tile_ok(A)) {
    launch_spmv_tile<DOT>(A, x, y, part, done, grid_override ? grid_override : tile_grid(A));
    return;
  }
  int lpr = pick_lpr(A);
  int grid = grid_override ? grid_override : spmv_grid(A.n, lpr);
  cudaStream_t st = stream();
  switch (lpr) {
    case 2: k_spmv<R, 2, DOT><<<grid, kBlock, 0, st>>>(A, x, y, part, done); break;
    case 4: k_spmv<R, 4, DOT><<<grid, kBlock, 0, st>>>(A, x, y, part, done); break;
    case 8: k_spmv<R, 8, DOT><<<grid, kBlock, 0, st>>>(A, x, y, part, done); break;
    case 16: k_spmv<R, 16, DOT><<<grid, kBlock, 0, st>>>(A, x, y, part, done); break;
    default: k_spmv<R, 32, DOT><<<grid, kBlock, 0, st>>>(A, x, y, part, done); break;
  }
  FS_LAUNCH_CHECK();
}

void spmv_dev(const CsrView& A, const double* d_x, double* d_y) { launch_spmv<1, false>(A, d_x, d_y, nullptr, nullptr); }

// ---- Jacobi ---------------------------------------------------------------------
__global__ void k_diag_inv(CsrView A, double* __restrict__ dinv) {
  int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (row >= A.n) return;
  double d = 0.0;
  for (int k = A.rowptr[row]; k < A.rowptr[row + 1]; ++k)
    if (A.colidx[k] == row) d = A.vals[k];
  dinv[row] = (d != 0.0) ? 1.0 / d : 1.0;
}

void jacobi_prepare(fs_csr* a) {
  if (a->dinv_ready) return;
  a->dinv.alloc(a->n);
  k_diag_inv<<<div_up(a->n, 256), 256, 0, stream()>>>(a->view(), a->dinv.p);
  FS_LAUNCH_CHECK();
  a->dinv_ready = true;
}

// ---- CG passes --------------------------------------------------------------------
// scalar block (doubles): slot s in {0,1}: rz[R] at s*R ; then bb[R] at 2R ; rr[R] at 3R
// ints live in a separate array: done, iters
struct CgScal {
  double* d;     // 4R doubles
  int* flags;    // [0]=done [1]=iters
};

// r = b - Ap (Ap holds A x0) ; z = Dinv r ; p = z ; partials {rr, rz, bb}
template <int R>
__global__ void __launch_bounds__(kBlock)
k_cg_init(int64_t n, const double* __restrict__ b, const double* __restrict__ Ap, const double* __restrict__ dinv,
          double* __restrict__ r, double* __restrict__ p, double* __restrict__ part) {
  __shared__ double red[3 * R * 32];
  double acc[3 * R];
#pragma unroll
  for (int k = 0; k < 3 * R; ++k) acc[k] = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double bv[R], av[R], rv[R], zv[R];
    load_vec<R>(b, i, bv);
    load_vec<R>(Ap, i, av);
    const double di = dinv ? dinv[i] : 1.0;
#pragma unroll
    for (int c = 0; c < R; ++c) {
      rv[c] = bv[c] - av[c];
      zv[c] = di * rv[c];
      acc[c] += rv[c] * rv[c];
      acc[R + c] += rv[c] * zv[c];
      acc[2 * R + c] += bv[c] * bv[c];
    }
    store_vec<R>(r, i, rv);
    store_vec<R>(p, i, zv);
  }
  block_reduce<3 * R>(acc, red);
  if (threadIdx.x == 0)
#pragma unroll
    for (int k = 0; k < 3 * R; ++k) part[(size_t)blockIdx.x * 3 * R + k] = acc[k];
}

template <int R>
__global__ void k_cg_init_fin(const double* __restrict__ part, int nblk, CgScal sc, double tol2) {
  __shared__ double sm[3 * R];
  double v[3 * R];
  reduce_partials<3 * R>(part, nblk, v, sm);
  if (threadIdx.x == 0) {
    bool all = true;
#pragma unroll
    for (int c = 0; c < R; ++c) {
      sc.d[c] = v[R + c];          // rz slot 0
      sc.d[2 * R + c] = v[2 * R + c];  // bb
      sc.d[3 * R + c] = v[c];      // rr
      if (!(v[c] <= tol2 * v[2 * R + c])) all = false;
    }
    sc.flags[0] = all ? 1 : 0;
    sc.flags[1] = 0;
  }
}

// pass B
template <int R>
__global__ void __launch_bounds__(kBlock)
k_cg_update_xr(int64_t n, const double* __restrict__ p, const double* __restrict__ Ap, const double* __restrict__ dinv,
               double* __restrict__ x, double* __restrict__ r, const double* __restrict__ part_pAp, int nblk_a,
               CgScal sc, int slot, double* __restrict__ part_out) {
  __shared__ double red[2 * R * 32];
  __shared__ double sm[R];
  if (sc.flags[0]) return;
  double pAp[R], alpha[R];
  reduce_partials<R>(part_pAp, nblk_a, pAp, sm);
#pragma unroll
  for (int c = 0; c < R; ++c) {
    double rz = sc.d[slot * R + c];
    alpha[c] = (pAp[c] != 0.0) ? rz / pAp[c] : 0.0;
  }
  double acc[2 * R];
#pragma unroll
  for (int k = 0; k < 2 * R; ++k) acc[k] = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double pv[R], av[R], xv[R], rv[R];
    load_vec<R>(p, i, pv);
    load_vec<R>(Ap, i, av);
    load_vec<R>(x, i, xv);
    load_vec<R>(r, i, rv);
    const double di = dinv ? dinv[i] : 1.0;
#pragma unroll
    for (int c = 0; c < R; ++c) {
      xv[c] += alpha[c] * pv[c];
      rv[c] -= alpha[c] * av[c];
      acc[c] += rv[c] * rv[c];
      acc[R + c] += rv[c] * (di * rv[c]);
    }
    store_vec<R>(x, i, xv);
    store_vec<R>(r, i, rv);
  }
  block_reduce<2 * R>(acc, red);
  if (threadIdx.x == 0)
#pragma unroll
    for (int k = 0; k < 2 * R; ++k) part_out[(size_t)blockIdx.x * 2 * R + k] = acc[k];
}

// pass C
template <int R>
__global__ void __launch_bounds__(kBlock)
k_cg_update_p(int64_t n, const double* __restrict__ r, const double* __restrict__ dinv, double* __restrict__ p,
              const double* __restrict__ part_b, int nblk_b, CgScal sc, int slot, double tol2) {
  __shared__ double sm[2 * R];
  if (block_done(sc.flags)) return;   // block 0 of THIS launch may set the flag: one read per CTA
  double v[2 * R], beta[R];
  reduce_partials<2 * R>(part_b, nblk_b, v, sm);
#pragma unroll
  for (int c = 0; c < R; ++c) {
    double rz_old = sc.d[slot * R + c];
    beta[c] = (rz_old != 0.0) ? v[R + c] / rz_old : 0.0;
  }
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double rv[R], pv[R];
    load_vec<R>(r, i, rv);
    load_vec<R>(p, i, pv);
    const double di = dinv ? dinv[i] : 1.0;
#pragma unroll
    for (int c = 0; c < R; ++c) pv[c] = di * rv[c] + beta[c] * pv[c];
    store_vec<R>(p, i, pv);
  }
  // block 0 publishes the scalars of the finished iteration into the other slot;
  // nobody reads that slot, rr or the counters during this launch.  A block that
  // starts after "done" is set skips its p update, which is no longer needed.
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    bool all = true;
#pragma unroll
    for (int c = 0; c < R; ++c) {
      sc.d[(slot ^ 1) * R + c] = v[R + c];
      sc.d[3 * R + c] = v[c];
      if (!(v[c] <= tol2 * sc.d[2 * R + c])) all = false;
    }
    sc.flags[1] += 1;
    if (all) sc.flags[0] = 1;
  }
}

// mean handling for the singular pressure operator
__global__ void __launch_bounds__(kBlock)
k_sum(int64_t n, const double* __restrict__ v, double* __restrict__ part) {
  __shared__ double red[32];
  double acc[1] = {0.0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) acc[0] += v[i];
  block_reduce<1>(acc, red);
  if (threadIdx.x == 0) part[blockIdx.x] = acc[0];
}

__global__ void __launch_bounds__(kBlock)
k_sub_mean(int64_t n, const double* in, double* out /* may be `in` */, const double* __restrict__ part, int nblk) {
  __shared__ double sm[1];
  double s[1];
  reduce_partials<1>(part, nblk, s, sm);
  const double mean = s[0] / (double)n;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = in[i] - mean;
}

__global__ void __launch_bounds__(kBlock)
k_maxabs(int64_t n, const double* __restrict__ v, double* __restrict__ part) {
  __shared__ double red[32];
  double m = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = fmax(m, fabs(v[i]));
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) part[blockIdx.x] = m;
  }
}

// ---- sampled per-pass timing (bench.py roofline): CUDA events around the three
// passes of every `every`-th CG iteration, harvested at the convergence polls.
struct PassProf {
  int every = 0;
  std::vector<cudaEvent_t> ev;
  int used = 0;
  double ms[3] = {0, 0, 0};
  long long samples = 0;
  long long iters = 0;   // CG iterations launched while enabled
  // AMG-PCG: the V-cycle's largest kernel (finest up-sweep), timed on every 8th iteration
  double top_ms = 0.0, top_bytes = 0.0;
  long long top_samples = 0;
};
static PassProf g_prof;

static void prof_harvest() {
  for (int k = 0; k + 3 < g_prof.used; k += 4) {
    for (int j = 0; j < 3; ++j) {
      float t = 0.f;
      if (cudaEventElapsedTime(&t, g_prof.ev[k + j], g_prof.ev[k + j + 1]) == cudaSuccess) g_prof.ms[j] += t;
    }
    g_prof.samples += 1;
  }
  g_prof.used = 0;
}

static int vec_grid(int64_t n) {
  int64_t want = (n + kBlock - 1) / kBlock;
  int64_t cap = std::min<int64_t>(kMaxBlocks, (int64_t)sm_count() * 6);
  return (int)std::max<int64_t>(1, std::min(want, cap));
}

static void ensure_ws(fs_csr* a, size_t doubles) {
  if (a->ws_n < doubles) { a->ws.alloc(doubles); a->ws_n = doubles; }
  if (!a->partials.n) a->partials.alloc((size_t)kMaxBlocks * 16);
  if (!a->scal.n) a->scal.alloc(64);
}

double max_abs_dev(const double* d_x, int64_t n) {
  int g = vec_grid(n);
  DBuf<double> part(g);
  k_maxabs<<<g, kBlock, 0, stream()>>>(n, d_x, part.p);
  FS_LAUNCH_CHECK();
  std::vector<double> h = part.to_host();
  double m = 0.0;
  for (double v : h) m = std::max(m, v);
  return m;
}

// ---- small systems: the whole (Jacobi-)PCG solve in ONE CTA (shipped meshes: a few hundred rows).
// Same recurrence and stopping rule as the large paths; R interleaved right-hand sides;
// project_mean handled inside.  No grid barrier, no host read until the end.
constexpr int kSmallCgN = 8192;

template <int R>
__global__ void __launch_bounds__(1024)
k_cg_small(CsrView A, const double* __restrict__ dinv, const double* __restrict__ b_in, double* __restrict__ x,
           double* __restrict__ ws, double tol2, int maxit, int project_mean, double* __restrict__ out /* rr[R], bb[R] */,
           int* __restrict__ flags) {
  __shared__ double red[32];
  const int n = A.n, t = threadIdx.x, nt = blockDim.x;
  // batch: CTA c solves system c (same matrix; right-hand side, solution, workspace and results c-th in their arrays)
  b_in += (size_t)blockIdx.x * n * R;
  x += (size_t)blockIdx.x * n * R;
  ws += (size_t)blockIdx.x * 4 * n * R;
  out += (size_t)blockIdx.x * 2 * R;
  flags += (size_t)blockIdx.x * 2;
  double *r = ws, *p = r + (size_t)n * R, *Ap = p + (size_t)n * R, *b = Ap + (size_t)n * R;
  auto csum = [&](double v) -> double {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((t & 31) == 0) red[t >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < (nt >> 5); ++w) s += red[w];
    return s;
  };
  // b (mean-free if requested)
  double mean = 0.0;
  if (project_mean) {
    double l = 0.0;
    for (int i = t; i < n; i += nt) l += b_in[i];
    mean = csum(l) / (double)n;
  }
  for (int i = t; i < n * R; i += nt) b[i] = b_in[i] - mean;
  __syncthreads();
  auto spmv = [&](const double* in, double* outv) {
    for (int i = t; i < n; i += nt) {
      double acc[R];
#pragma unroll
      for (int c = 0; c < R; ++c) acc[c] = 0.0;
      for (int k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) {
        const double av = A.vals[k];
        const int j = A.colidx[k];
#pragma unroll
        for (int c = 0; c < R; ++c) acc[c] += av * in[(size_t)j * R + c];
      }
#pragma unroll
      for (int c = 0; c < R; ++c) outv[(size_t)i * R + c] = acc[c];
    }
    __syncthreads();
  };
  spmv(x, Ap);
  double rr[R], rz[R], bb[R];
  {
    double l[3 * R];
#pragma unroll
    for (int k = 0; k < 3 * R; ++k) l[k] = 0.0;
    for (int i = t; i < n; i += nt) {
      const double di = dinv ? dinv[i] : 1.0;
#pragma unroll
      for (int c = 0; c < R; ++c) {
        const double bv = b[(size_t)i * R + c], rv = bv - Ap[(size_t)i * R + c], zv = di * rv;
        r[(size_t)i * R + c] = rv;
        p[(size_t)i * R + c] = zv;
        l[c] += rv * rv; l[R + c] += rv * zv; l[2 * R + c] += bv * bv;
      }
    }
#pragma unroll
    for (int c = 0; c < R; ++c) { rr[c] = csum(l[c]); rz[c] = csum(l[R + c]); bb[c] = csum(l[2 * R + c]); }
  }
  auto all_done = [&]() { bool d = true;
#pragma unroll
    for (int c = 0; c < R; ++c) d = d && (rr[c] <= tol2 * bb[c]);
    return d; };
  bool zero_rhs = true;
#pragma unroll
  for (int c = 0; c < R; ++c) zero_rhs = zero_rhs && (bb[c] == 0.0);
  int it = 0;
  if (zero_rhs) {
    for (int i = t; i < n * R; i += nt) x[i] = 0.0;
#pragma unroll
    for (int c = 0; c < R; ++c) rr[c] = 0.0;
  } else {
    while (!all_done() && it < maxit) {
      spmv(p, Ap);
      double l[R];
#pragma unroll
      for (int c = 0; c < R; ++c) l[c] = 0.0;
      for (int i = t; i < n; i += nt)
#pragma unroll
        for (int c = 0; c < R; ++c) l[c] += p[(size_t)i * R + c] * Ap[(size_t)i * R + c];
      double alpha[R];
#pragma unroll
      for (int c = 0; c < R; ++c) { const double pAp = csum(l[c]); alpha[c] = pAp != 0.0 ? rz[c] / pAp : 0.0; }
      double l2[2 * R];
#pragma unroll
      for (int k = 0; k < 2 * R; ++k) l2[k] = 0.0;
      for (int i = t; i < n; i += nt) {
        const double di = dinv ? dinv[i] : 1.0;
#pragma unroll
        for (int c = 0; c < R; ++c) {
          const size_t q = (size_t)i * R + c;
          x[q] += alpha[c] * p[q];
          const double rn = r[q] - alpha[c] * Ap[q];
          r[q] = rn;
          l2[c] += rn * rn; l2[R + c] += rn * (di * rn);
        }
      }
      double beta[R];
#pragma unroll
      for (int c = 0; c < R; ++c) {
        rr[c] = csum(l2[c]);
        const double rzn = csum(l2[R + c]);
        beta[c] = rz[c] != 0.0 ? rzn / rz[c] : 0.0;
        rz[c] = rzn;
      }
      ++it;
      if (all_done()) break;
      for (int i = t; i < n; i += nt) {
        const double di = dinv ? dinv[i] : 1.0;
#pragma unroll
        for (int c = 0; c < R; ++c) { const size_t q = (size_t)i * R + c; p[q] = di * r[q] + beta[c] * p[q]; }
      }
      __syncthreads();
    }
  }
  if (project_mean) {
    double l = 0.0;
    for (int i = t; i < n; i += nt) l += x[i];
    const double mx = csum(l) / (double)n;
    for (int i = t; i < n; i += nt) x[i] -= mx;
  }
  if (t == 0) {
#pragma unroll
    for (int c = 0; c < R; ++c) { out[c] = rr[c]; out[R + c] = bb[c]; }
    flags[0] = (zero_rhs || all_done()) ? 1 : 0;
    flags[1] = it;
  }
}

template <int R>
static int cg_small(fs_csr* a, const double* d_b, double* d_x, double rtol, int maxit, int precond, int project_mean,
                    double* relres) {
  const int64_t n = a->n;
  ensure_ws(a, 4 * (size_t)n * R);
  const double* dinv = nullptr;
  if (precond == FS_PRECOND_JACOBI) { jacobi_prepare(a); dinv = a->dinv.p; }
  cudaStream_t st = stream();
  double* out = a->scal.p;
  int* flags = reinterpret_cast<int*>(a->scal.p + 32);
  k_cg_small<R><<<1, 1024, 0, st>>>(a->view(), dinv, d_b, d_x, a->ws.p, rtol * rtol, maxit, project_mean, out, flags);
  FS_LAUNCH_CHECK();
  double ho[2 * R];
  int hf[2];
  FS_CUDA(cudaMemcpyAsync(ho, out, sizeof(ho), cudaMemcpyDeviceToHost, st));
  FS_CUDA(cudaMemcpyAsync(hf, flags, sizeof(hf), cudaMemcpyDeviceToHost, st));
  FS_CUDA(cudaStreamSynchronize(st));
  double worst = 0.0;
  for (int c = 0; c < R; ++c) if (ho[R + c] > 0.0) worst = std::max(worst, std::sqrt(ho[c] / ho[R + c]));
  if (relres) *relres = worst;
  return hf[0] ? hf[1] : -hf[1] - 1;
}

// B independent systems with the same (small) matrix: one launch, one CTA per system.  x holds the initial guesses.
// ws: B*4*n*nrhs doubles, out: B*2*nrhs, flags: B*2 ints (device).  No host synchronisation.
void cg_small_batch_dev(fs_csr* a, const double* d_b, double* d_x, int nrhs, int B, double rtol, int maxit, int precond,
                        int project_mean, double* ws, double* out, int* flags) {
  FS_REQUIRE(a->n <= kSmallCgN, "cg_small_batch_dev: the matrix is too large for the single-CTA solver");
  const double* dinv = nullptr;
  if (precond == FS_PRECOND_JACOBI) { jacobi_prepare(a); dinv = a->dinv.p; }
  if (nrhs == 1) k_cg_small<1><<<B, 1024, 0, stream()>>>(a->view(), dinv, d_b, d_x, ws, rtol * rtol, maxit, project_mean, out, flags);
  else k_cg_small<2><<<B, 1024, 0, stream()>>>(a->view(), dinv, d_b, d_x, ws, rtol * rtol, maxit, project_mean, out, flags);
  FS_LAUNCH_CHECK();
}
int cg_small_limit() { return kSmallCgN; }

// FS_CG_MODE=multi forces the 3-kernels-per-iteration path (A/B testing); default persistent
static int cg_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = std::getenv("FS_CG_MODE");
    mode = (e && std::string(e) == "multi") ? 0 : 1;
  }
  return mode;
}

template <int R>
static int cg_impl(fs_csr* a, const double* d_b, double* d_x, double rtol, int maxit, int precond, int project_mean,
                   double* relres) {
  const int64_t n = a->n;
  if (n <= kSmallCgN && (precond == FS_PRECOND_JACOBI || precond == FS_PRECOND_NONE) && (!project_mean || R == 1) &&
      cg_mode() == 1)
    return cg_small<R>(a, d_b, d_x, rtol, maxit, precond, project_mean, relres);
  const size_t len = (size_t)n * R;
  ensure_ws(a, 4 * len);
  double* r = a->ws.p;
  double* p = r + len;
  double* Ap = p + len;
  double* bproj = Ap + len;
  const double* dinv = nullptr;
  if (precond == FS_PRECOND_JACOBI) { jacobi_prepare(a); dinv = a->dinv.p; }
 cudaStream_t st = stream();
  ensure_tiles(a);
  const CsrView A = a->view();
  double* partA = a->partials.p;                    // pass A partials (R per block)
  double* partB = a->partials.p + kMaxBlocks * 4;   // pass B partials (2R per block)
  double* part0 = a->partials.p + kMaxBlocks * 8;   // init partials (3R per block) / mean sums
  CgScal sc{a->scal.p, reinterpret_cast<int*>(a->scal.p + 32)};
  const int gv = vec_grid(n);
  const double tol2 = rtol * rtol;

  const double* b_use = d_b;
  if (project_mean) {
    FS_REQUIRE(R == 1, "project_mean needs nrhs == 1");
    k_sum<<<gv, kBlock, 0, st>>>(n, d_b, part0);
    FS_LAUNCH_CHECK();
    k_sub_mean<<<gv, kBlock, 0, st>>>(n, d_b, bproj, part0, gv);
    FS_LAUNCH_CHECK();
    b_use = bproj;
  }
  // two right-hand sides on a large matrix (the viscous solve): SELL-32 copy, A streamed once for both
  static const bool sell2 = [] { const char* e = std::getenv("FS_CG_SELL2"); return !e || std::atoi(e) != 0; }();
  const bool use_sell2 = R == 2 && sell2 && n > 100000;
  if (use_sell2 && !a->sell64) {
    a->sell64 = new fs_sell();
    sell_build(*a, false, *a->sell64);
  }
  if (use_sell2) spmv_sell2(*a->sell64, d_x, Ap, nullptr, nullptr);
  else launch_spmv<R, false>(A, d_x, Ap, nullptr, nullptr);
  k_cg_init<R><<<gv, kBlock, 0, st>>>(n, b_use, Ap, dinv, r, p, part0);
  FS_LAUNCH_CHECK();
  k_cg_init_fin<R><<<1, 32, 0, st>>>(part0, gv, sc, tol2);
  FS_LAUNCH_CHECK();

  struct HostScal { double d[8]; int flags[2]; } hs;
  auto poll = [&]() {
    FS_CUDA(cudaMemcpyAsync(hs.d, sc.d, sizeof(double) * 4 * R, cudaMemcpyDeviceToHost, st));
    FS_CUDA(cudaMemcpyAsync(hs.flags, sc.flags, sizeof(int) * 2, cudaMemcpyDeviceToHost, st));
    FS_CUDA(cudaStreamSynchronize(st));
  };
  poll();
  bool zero_rhs = true;
  for (int c = 0; c < R; ++c) if (hs.d[2 * R + c] != 0.0) zero_rhs = false;
  if (zero_rhs) {
    FS_CUDA(cudaMemsetAsync(d_x, 0, len * sizeof(double), st));
    if (relres) *relres = 0.0;
    return 0;
  }
  // single-RHS solves run as one persistent cooperative kernel (cg_persistent.cu)
  if (R == 1 && !hs.flags[0] && cg_mode() == 1 && cg_persistent_supported(A, nullptr)) {
    cg_persistent_launch(A, d_x, r, p, Ap, dinv, partA, partB, sc.d, sc.flags, maxit, tol2);
    double tim[4];
    FS_CUDA(cudaMemcpyAsync(tim, sc.d + 8, sizeof(tim), cudaMemcpyDeviceToHost, st));
    poll();
    if (g_prof.every > 0) {
      for (int j = 0; j < 3; ++j) g_prof.ms[j] += tim[j] * 1e-6;
      g_prof.samples += (long long)tim[3];
      g_prof.iters += (long long)tim[3];
    }
    maxit = 0;   // skip the multi-kernel loop below
  }
  const int ga = use_sell2 ? spmv_sell_grid(*a->sell64) : spmv_launch_grid<R>(A);
  int launched = 0, chunk = 8, slot = 0;
  while (!hs.flags[0] && launched < maxit) {
    int todo = std::min(chunk, maxit - launched);
    for (int k = 0; k < todo; ++k) {
      const bool samp = R == 1 && g_prof.every > 0 && ((launched + k) % g_prof.every == 0) &&
                        g_prof.used + 4 <= (int)g_prof.ev.size();
      if (samp) cudaEventRecord(g_prof.ev[g_prof.used++], st);
      if (use_sell2) spmv_sell2(*a->sell64, p, Ap, partA, sc.flags);
      else launch_spmv<R, true>(A, p, Ap, partA, sc.flags, ga);
      if (samp) cudaEventRecord(g_prof.ev[g_prof.used++], st);
      k_cg_update_xr<R><<<gv, kBlock, 0, st>>>(n, p, Ap, dinv, d_x, r, partA, ga, sc, slot, partB);
      FS_LAUNCH_CHECK();
      if (samp) cudaEventRecord(g_prof.ev[g_prof.used++], st);
      k_cg_update_p<R><<<gv, kBlock, 0, st>>>(n, r, dinv, p, partB, gv, sc, slot, tol2);
      FS_LAUNCH_CHECK();
      if (samp) cudaEventRecord(g_prof.ev[g_prof.used++], st);
      slot ^= 1;
    }
    launched += todo;
    poll();
    if (g_prof.every > 0 && R == 1) {
      prof_harvest();
      g_prof.iters = g_prof.iters + todo;
    }
    chunk = std::min(chunk * 2, 128);
  }
  if (project_mean) {
    k_sum<<<gv, kBlock, 0, st>>>(n, d_x, part0);
    FS_LAUNCH_CHECK();
    k_sub_mean<<<gv, kBlock, 0, st>>>(n, d_x, d_x, part0, gv);
    FS_LAUNCH_CHECK();
  }
  double worst = 0.0;
  for (int c = 0; c < R; ++c) {
    double bb = hs.d[2 * R + c], rr = hs.d[3 * R + c];
    if (bb > 0.0) worst = std::max(worst, std::sqrt(rr / bb));
  }
  if (relres) *relres = worst;
  return hs.flags[0] ? hs.flags[1] : -hs.flags[1] - 1;   // negative: not converged
}

__global__ void k_lin3(int64_t n, double a, const double* x, double b, const double* y, double c, const double* z, double* out);

// ---- CG preconditioned by one AMG V-cycle (amg.cu).  ~50x fewer iterations than Jacobi on
// the 4M-triangle pressure operator.  All scalars live on the device:
//   sc[0], sc[1]  r.z of the previous / current iteration (slot = iteration parity)
//   sc[2] b.b   sc[3] latest r.r   sc[4] tol^2      flags[0] converged   flags[1] iterations done
// alpha is formed in k_pcg_xr from the SpMV's partials, beta and the convergence test in k_pcg_p
// from the partials of r.r (k_pcg_xr) and r.z (the V-cycle's last kernel); both kernels return at
// once when flags[0] is set, so x stays the converged iterate however many more iterations are
// queued.  The host therefore queues iterations without reading anything back for as long as the
// previous solve on this matrix needed (minus 3), and only then polls the flag once per iteration.
static void dot2(int64_t n, const double* a, const double* b, const double* c, const double* d, double* part, double* out2);

__global__ void k_pcg_init(const double* __restrict__ partRZ, int nrz, double bb, double rr, double tol2,
                           double* __restrict__ sc, int* __restrict__ flags) {
  __shared__ double sm[1];
  double rz[1];
  reduce_partials<1>(partRZ, nrz, rz, sm);
  if (threadIdx.x == 0) {
    sc[0] = rz[0]; sc[1] = 0.0; sc[2] = bb; sc[3] = rr; sc[4] = tol2;
    flags[0] = 0; flags[1] = 0;
  }
}

// x += alpha p ; r -= alpha Ap with alpha = rz / (p.Ap); partial r.r out.
__global__ void __launch_bounds__(kBlock)
k_pcg_xr(int64_t n, const double* __restrict__ p, const double* __restrict__ Ap, double* __restrict__ x, double* __restrict__ r,
         const double* __restrict__ partA, int nblkA, const double* __restrict__ sc, int slot, const int* __restrict__ flags,
         double* __restrict__ partB, double* __restrict__ x0out, const double* __restrict__ dinv, double w, float* __restrict__ r32) {
  __shared__ double red[32];
  __shared__ double sm[1];
  pdl_launch();
  pdl_wait();
  if (flags[0]) return;
  double pAp[1];
  reduce_partials<1>(partA, nblkA, pAp, sm);
  const double rz = sc[slot];
  const double alpha = (pAp[0] != 0.0) ? rz / pAp[0] : 0.0;
  double acc[1] = {0.0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double rn = r[i] - alpha * Ap[i];
    x[i] += alpha * p[i];
    r[i] = rn;
    if (r32) r32[i] = (float)rn;                // fp32 mirror for the V-cycle's finest-level gathers
    if (x0out) x0out[i] = w * dinv[i] * rn;     // the unfolded V-cycle's pre-smoothed iterate, for free
    acc[0] += rn * rn;
  }
  block_reduce<1>(acc, red);
  if (threadIdx.x == 0) partB[blockIdx.x] = acc[0];
}

__global__ void __launch_bounds__(kBlock)
k_pcg_rz(int64_t n, const double* __restrict__ r, const double* __restrict__ z, double* __restrict__ partC) {
  __shared__ double red[32];
  double acc[1] = {0.0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) acc[0] += r[i] * z[i];
  block_reduce<1>(acc, red);
  if (threadIdx.x == 0) partC[blockIdx.x] = acc[0];
}

// convergence test, beta = r.z / (r.z)_old, p = z + beta p; block 0 publishes the scalars
__global__ void __launch_bounds__(kBlock)
k_pcg_p(int64_t n, const double* __restrict__ z, double* __restrict__ p, const double* __restrict__ partB, int nB,
        const double* __restrict__ partRZ, int nRZ, double* __restrict__ sc, int slot, int* __restrict__ flags) {
  __shared__ double sm[1];
  pdl_launch();
  pdl_wait();
  if (block_done(flags)) return;      // block 0 of THIS launch may set the flag: one read per CTA
  double rr[1], rzn[1];
  reduce_partials<1>(partB, nB, rr, sm);
  reduce_partials<1>(partRZ, nRZ, rzn, sm);
  const double rz_old = sc[slot];
  const bool conv = rr[0] <= sc[4] * sc[2];
  if (!conv) {
    const double beta = rz_old != 0.0 ? rzn[0] / rz_old : 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
      p[i] = z[i] + beta * p[i];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    sc[slot ^ 1] = rzn[0];
    sc[3] = rr[0];
    flags[1] += 1;
    if (conv) flags[0] = 1;
  }
}

// L2 residency of the CG vectors (FS_L2_PERSIST=1; OFF by default -- measured slower): an access-policy window on the
// library stream marks their lines persisting for the duration of a solve (the B200 lets 79 MB of its 126 MB L2 be set
// aside, the window is r, p, Ap, z, r32 = 76 MB at 4M triangles); everything else keeps its normal policy, the matrix
// streams carry their own evict-first hint.  Result at 4M triangles (profiles/r02_ab_round2b.txt): the two vector
// kernels 39.3 -> 34.2 us per iteration, but A*p 39.4 -> 42.5 us and the V-cycle 125.6 -> 136.9 us -- the ~430 MB of
// matrix entries per iteration then stream through the remaining 47 MB and evict each other's gather lines.
struct L2Persist {
  bool on = false;
  L2Persist(void* base, size_t bytes) {
    static const int mode = [] { const char* e = std::getenv("FS_L2_PERSIST"); return e ? std::atoi(e) : 0; }();
    static size_t max_persist = 0, max_window = 0, set_aside = 0;
    static bool queried = false;
    if (!mode || bytes < (8u << 20)) return;
    if (!queried) {
      queried = true;
      int dev = 0, v = 0;
      cudaGetDevice(&dev);
      if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxPersistingL2CacheSize, dev) == cudaSuccess) max_persist = (size_t)std::max(v, 0);
      if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxAccessPolicyWindowSize, dev) == cudaSuccess) max_window = (size_t)std::max(v, 0);
      if (std::getenv("FS_AMG_VERBOSE")) std::fprintf(stderr, "[l2] max persisting %zu MB, max window %zu MB\n", max_persist >> 20, max_window >> 20);
      cudaGetLastError();
    }
    if (!max_persist || !max_window) return;
    const size_t want = std::min(bytes, max_persist);
    if (want > set_aside) {
      if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) != cudaSuccess) { cudaGetLastError(); return; }
      set_aside = want;
    }
    cudaStreamAttrValue v = {};
    v.accessPolicyWindow.base_ptr = base;
    v.accessPolicyWindow.num_bytes = std::min(bytes, max_window);
    v.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)set_aside / (double)v.accessPolicyWindow.num_bytes);
    v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    v.accessPolicyWindow.missProp = mode == 2 ? cudaAccessPropertyStreaming : cudaAccessPropertyNormal;
    if (cudaStreamSetAttribute(stream(), cudaStreamAttributeAccessPolicyWindow, &v) != cudaSuccess) { cudaGetLastError(); return; }
    on = true;
  }
  ~L2Persist() {
    if (!on) return;
    cudaStreamAttrValue v = {};
    v.accessPolicyWindow.num_bytes = 0;
    cudaStreamSetAttribute(stream(), cudaStreamAttributeAccessPolicyWindow, &v);
    static const bool reset = [] { const char* e = std::getenv("FS_L2_PERSIST_RESET"); return e && std::atoi(e) != 0; }();
    if (reset) cudaCtxResetPersistingL2Cache();
    cudaGetLastError();
  }
};

// fp32 mirror of r for the gathers of the packed V-cycle kernels (FS_PCG_R32=0: fp64 gathers, the arithmetic of the
// partitioned cycle; read at every solve so that a caller can compare the two)
static bool pcg_use_r32() {
  const char* e = std::getenv("FS_PCG_R32");
  return !e || std::atoi(e) != 0;
}

__global__ void k_mirror_f32(int64_t n, const double* __restrict__ in, float* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = (float)in[i];
}

static int pcg_amg_impl(fs_csr* a, const double* d_b, double* x, double rtol, int maxit, int project_mean, double* relres) {
  const int64_t n = a->n;
  const size_t n32 = ((size_t)n + 1) / 2;      // doubles that hold the fp32 mirror of r
  ensure_ws(a, 5 * (size_t)n + n32);
  double *r = a->ws.p, *p = r + n, *Ap = p + n, *z = Ap + n, *bproj = z + n + n32;
  // fp32 gathers in the cycle's finest kernels perturb the preconditioner by ~1e-7 of the gathered terms per application;
  // on smooth residuals those terms cancel, and below a relative residual of ~1e-12 the perturbation shows: 133 instead
  // of 66 iterations to 1e-13 at 4M triangles, no difference at 1e-12 and above (scripts/exp_tight_rtol.py).  Tighter
  // solves gather in fp64.
  float* r32 = (pcg_use_r32() && rtol >= 1e-12) ? reinterpret_cast<float*>(z + n) : nullptr;
  // the vectors every kernel of the iteration re-reads (r, p, Ap, z, r32: 76 MB at 4M triangles) stay in L2 while the
  // matrix streams (~430 MB per iteration) pass through it
  L2Persist keep(r, (4 * (size_t)n + n32) * sizeof(double));
  if (!a->amg) a->amg = amg_setup(a);
  ensure_tiles(a);
  static const bool sell_ap = [] { const char* e = std::getenv("FS_PCG_SELL"); return !e || std::atoi(e) != 0; }();
  static const bool lagged = [] { const char* e = std::getenv("FS_PCG_LAGGED"); return !e || std::atoi(e) != 0; }();
  if (sell_ap && !a->sell64 && n > 100000) {   // the CG's own A*p in the same SELL-32 layout (fp64 values)
    a->sell64 = new fs_sell();
    sell_build(*a, false, *a->sell64);
  }
  const CsrView A = a->view();
  cudaStream_t st = stream();
  const int g = vec_grid(n);
  double* part = a->partials.p;
  double* part0 = a->partials.p + kMaxBlocks * 8;
  double* partA = a->partials.p + kMaxBlocks * 10;     // p.Ap partials of the SpMV
  double* partB = a->partials.p + kMaxBlocks * 12;     // r.r partials of k_pcg_xr
  double* partRZ = partB + kMaxBlocks;                 // r.z partials: the cycle's last kernel (folded cycle) or k_pcg_rz
  double* sc = a->scal.p;
  int* flags = reinterpret_cast<int*>(a->scal.p + 32);
  const double* b = d_b;
  if (project_mean) {
    k_sum<<<g, kBlock, 0, st>>>(n, d_b, part0); FS_LAUNCH_CHECK();
    k_sub_mean<<<g, kBlock, 0, st>>>(n, d_b, bproj, part0, g); FS_LAUNCH_CHECK();
    b = bproj;
  }
  // A constant component of z is harmless for a mean-free r (r.z, A p do not see it) and is
  // removed from x at the end, so z is not projected every iteration.
  // Profiling (bench.py): 7 events per iteration from the pool -- A*p [0,1], V-cycle [2,3], end [4], and on every
  // 8th iteration the cycle runs eagerly with [5,6] around its largest kernel; read back after the solve.
  const bool prof = g_prof.every > 0;
  constexpr int kEvPer = 7;
  if (prof && g_prof.ev.size() < 640) {
    const size_t old = g_prof.ev.size();
    g_prof.ev.resize(640);
    for (size_t k = old; k < g_prof.ev.size(); ++k) cudaEventCreate(&g_prof.ev[k]);
  }
  const int ev_iters = prof ? (int)(g_prof.ev.size() / kEvPer) : 0;
  std::vector<char> top_sampled;
  double* x0 = nullptr;
  const double* dinv0 = nullptr;
  double w0 = 0.0;
  amg_presmooth_target(a->amg, &x0, &dinv0, &w0);
  int napply = 0;
  auto precond = [&](const double* rin, double* zout, bool x0_ready, cudaEvent_t* ev /* 7 events or null */) -> int {
    if (ev) cudaEventRecord(ev[2], st);
    const bool samp_top = ev && (napply % 8 == 4);
    ++napply;
    int nrz = amg_apply(a->amg, rin, zout, x0_ready, partRZ, samp_top ? ev + 5 : nullptr, r32);
    if (ev) { cudaEventRecord(ev[3], st); top_sampled.push_back(samp_top ? 1 : 0); }
    if (!nrz) {
      k_pcg_rz<<<g, kBlock, 0, st>>>(n, rin, zout, partRZ); FS_LAUNCH_CHECK();
      nrz = g;
    }
    return nrz;
  };
  double d2[2];
  if (!spmv_warp(A, EPI_RESID, x, r, b, nullptr, 0.0, nullptr, nullptr)) {
    spmv_dev(A, x, Ap);
    k_lin3<<<g, kBlock, 0, st>>>(n, 1.0, b, -1.0, Ap, 0.0, nullptr, r); FS_LAUNCH_CHECK();
  }
  dot2(n, b, b, r, r, part, d2);
  const double bb = d2[0];
  double rr = d2[1];
  if (bb == 0.0) { FS_CUDA(cudaMemsetAsync(x, 0, n * sizeof(double), st)); if (relres) *relres = 0.0; return 0; }
  const double tol2 = rtol * rtol;
  int it = 0;
  bool done = rr <= tol2 * bb;
  if (!done) {
    if (r32) { k_mirror_f32<<<g, kBlock, 0, st>>>(n, r, r32); FS_LAUNCH_CHECK(); }
    int nrz = precond(r, z, false, nullptr);
    FS_CUDA(cudaMemcpyAsync(p, z, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    k_pcg_init<<<1, kBlock, 0, st>>>(partRZ, nrz, bb, rr, tol2, sc, flags); FS_LAUNCH_CHECK();
    const int unchecked = lagged ? std::max(0, std::min(std::min(a->pcg_hint, a->pcg_hint_prev) - 3, maxit)) : 0;
    int queued = 0, hflags[2] = {0, 0};
    while (queued < maxit) {
      cudaEvent_t* ev = (prof && queued < ev_iters) ? &g_prof.ev[(size_t)queued * kEvPer] : nullptr;
      const int slot = queued & 1;
      if (ev) cudaEventRecord(ev[0], st);
      int ga = a->sell64 ? spmv_sell(*a->sell64, p, Ap, nullptr, partA) : 0;
      if (!ga) ga = spmv_warp(A, EPI_AX, p, Ap, nullptr, nullptr, 0.0, nullptr, partA);
      if (!ga) { ga = spmv_launch_grid<1>(A); launch_spmv<1, true>(A, p, Ap, partA, nullptr, ga); }
      if (ev) cudaEventRecord(ev[1], st);
      launch_pdl(k_pcg_xr, g, kBlock, 0, n, p, Ap, x, r, partA, ga, sc, slot, flags, partB, x0, dinv0, w0, r32); FS_LAUNCH_CHECK();
      nrz = precond(r, z, x0 != nullptr, ev);
      launch_pdl(k_pcg_p, g, kBlock, 0, n, z, p, partB, g, partRZ, nrz, sc, slot, flags); FS_LAUNCH_CHECK();
      if (ev) cudaEventRecord(ev[4], st);
      ++queued;
      if (queued > unchecked) {
        FS_CUDA(cudaMemcpyAsync(hflags, flags, sizeof(hflags), cudaMemcpyDeviceToHost, st));
        FS_CUDA(cudaStreamSynchronize(st));
        if (hflags[0]) break;
      }
    }
    double hsc[5];
    FS_CUDA(cudaMemcpyAsync(hflags, flags, sizeof(hflags), cudaMemcpyDeviceToHost, st));
    FS_CUDA(cudaMemcpyAsync(hsc, sc, sizeof(hsc), cudaMemcpyDeviceToHost, st));
    FS_CUDA(cudaStreamSynchronize(st));
    done = hflags[0] != 0;
    it = hflags[1];
    rr = hsc[3];
    if (done) { a->pcg_hint_prev = a->pcg_hint; a->pcg_hint = it; }
    if (prof) {
      const int nk = std::min(std::min(it, queued), ev_iters);
      for (int k = 0; k < nk; ++k) {
        cudaEvent_t* ev = &g_prof.ev[(size_t)k * kEvPer];
        float t_spmv = 0.f, t_v = 0.f, t_it = 0.f, t_top = 0.f;
        if (cudaEventElapsedTime(&t_spmv, ev[0], ev[1]) != cudaSuccess || cudaEventElapsedTime(&t_v, ev[2], ev[3]) != cudaSuccess ||
            cudaEventElapsedTime(&t_it, ev[0], ev[4]) != cudaSuccess) { cudaGetLastError(); continue; }
        g_prof.ms[0] += t_spmv; g_prof.ms[1] += t_v; g_prof.ms[2] += t_it - t_spmv - t_v;
        g_prof.samples += 1;
        if (k < (int)top_sampled.size() && top_sampled[k] && cudaEventElapsedTime(&t_top, ev[5], ev[6]) == cudaSuccess && t_top > 0.f) {
          g_prof.top_ms += t_top; g_prof.top_samples += 1; g_prof.top_bytes = amg_top_bytes(a->amg);
        }
      }
      cudaGetLastError();
      g_prof.iters += it;
    }
  }
  if (project_mean) {
    k_sum<<<g, kBlock, 0, st>>>(n, x, part0); FS_LAUNCH_CHECK();
    k_sub_mean<<<g, kBlock, 0, st>>>(n, x, x, part0, g); FS_LAUNCH_CHECK();
  }
  if (relres) *relres = std::sqrt(rr / bb);
  return done ? it : -it - 1;
}

int cg_dev(fs_csr* a, const double* d_b, double* d_x, int nrhs, double rtol, int maxit, int precond, int project_mean,
           double* relres) {
  FS_REQUIRE(nrhs == 1 || nrhs == 2, "nrhs must be 1 or 2");
  if (precond == FS_PRECOND_AUTO)   // multigrid pays off once the Jacobi iteration count (~1/h) is large
    precond = (nrhs == 1 && a->n > 20000) ? FS_PRECOND_AMG : FS_PRECOND_JACOBI;
  if (precond == FS_PRECOND_AMG) {
    FS_REQUIRE(nrhs == 1, "the AMG preconditioner handles one right-hand side");
    return pcg_amg_impl(a, d_b, d_x, rtol, maxit, project_mean, relres);
  }
  return nrhs == 1 ? cg_impl<1>(a, d_b, d_x, rtol, maxit, precond, project_mean, relres)
                   : cg_impl<2>(a, d_b, d_x, rtol, maxit, precond, project_mean, relres);
}

// ---- small vector kernels for BiCGStab (host-driven scalars; small systems) -----
__global__ void __launch_bounds__(kBlock)
k_dot2(int64_t n, const double* __restrict__ a, const double* __restrict__ b, const double* __restrict__ c,
       const double* __restrict__ d, double* __restrict__ part) {
  __shared__ double red[64];
  double acc[2] = {0.0, 0.0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    acc[0] += a[i] * b[i];
    if (c) acc[1] += c[i] * d[i];
  }
  block_reduce<2>(acc, red);
  if (threadIdx.x == 0) { part[2 * blockIdx.x] = acc[0]; part[2 * blockIdx.x + 1] = acc[1]; }
}

// out = a*x + b*y + c*z (any of y,z may be null)
// out = a x + b y + c z, element by element; out may be any of the inputs (no __restrict__)
__global__ void k_lin3(int64_t n, double a, const double* x, double b, const double* y, double c, const double* z, double* out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double v = a * x[i];
    if (y) v += b * y[i];
    if (z) v += c * z[i];
    out[i] = v;
  }
}

__global__ void k_mul(int64_t n, const double* __restrict__ d, const double* __restrict__ x, double* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = d ? d[i] * x[i] : x[i];
}

static void dot2(int64_t n, const double* a, const double* b, const double* c, const double* d, double* part, double* out2) {
  int g = std::min(vec_grid(n), 256);
  k_dot2<<<g, kBlock, 0, stream()>>>(n, a, b, c, d, part);
  FS_LAUNCH_CHECK();
  std::vector<double> h(2 * g);
  FS_CUDA(cudaMemcpyAsync(h.data(), part, sizeof(double) * 2 * g, cudaMemcpyDeviceToHost, stream()));
  FS_CUDA(cudaStreamSynchronize(stream()));
  out2[0] = out2[1] = 0.0;
  for (int i = 0; i < g; ++i) { out2[0] += h[2 * i]; out2[1] += h[2 * i + 1]; }
}

// y = A x through the fastest layout the handle has (SELL-32 copy, else CSR-stream, else CSR-vector)
void spmv_best_dev(fs_csr* a, const double* x, double* y) {
  if (a->sell64 && spmv_sell(*a->sell64, x, y, nullptr, nullptr)) return;
  ensure_tiles(a);
  const CsrView A = a->view();
  if (!spmv_warp(A, EPI_AX, x, y, nullptr, nullptr, 0.0, nullptr, nullptr)) spmv_dev(A, x, y);
}

// squared norms of b - y0, b - (2 y0 - y1), b - (3 y0 - 3 y1 + y2): the residuals of the three warm-start
// candidates q, 2q - q1, 3q - 3q1 + q2 given y_k = A q_k (y2 may be null: only the first two)
__global__ void __launch_bounds__(kBlock)
k_cand3(int64_t n, const double* __restrict__ b, const double* __restrict__ y0, const double* __restrict__ y1,
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

void cand_norms3_dev(fs_csr* a, const double* b, const double* y0, const double* y1, const double* y2, double* out3) {
  ensure_ws(a, 4 * (size_t)a->n);
  const int64_t n = a->n;
  const int g = std::min(vec_grid(n), 256);
  k_cand3<<<g, kBlock, 0, stream()>>>(n, b, y0, y1, y2, a->partials.p);
  FS_LAUNCH_CHECK();
  std::vector<double> h(3 * (size_t)g);
  FS_CUDA(cudaMemcpyAsync(h.data(), a->partials.p, sizeof(double) * 3 * g, cudaMemcpyDeviceToHost, stream()));
  FS_CUDA(cudaStreamSynchronize(stream()));
  out3[0] = out3[1] = out3[2] = 0.0;
  for (int i = 0; i < g; ++i) for (int k = 0; k < 3; ++k) out3[k] += h[3 * i + k];
}

void lin3_dev(int64_t n, double a, const double* x, double b, const double* y, double* out) {
  k_lin3<<<vec_grid(n), kBlock, 0, stream()>>>(n, a, x, b, y, 0.0, nullptr, out);
  FS_LAUNCH_CHECK();
}

// |b - A x|^2 (host value; `work` holds n doubles).  Used to choose between warm-start candidates.
double resid_norm2_dev(fs_csr* a, const double* b, const double* x, double* work) {
  ensure_ws(a, 4 * (size_t)a->n);
  ensure_tiles(a);
  const int64_t n = a->n;
  const CsrView A = a->view();
  if (!spmv_warp(A, EPI_RESID, x, work, b, nullptr, 0.0, nullptr, nullptr)) {
    spmv_dev(A, x, work);
    k_lin3<<<vec_grid(n), kBlock, 0, stream()>>>(n, 1.0, b, -1.0, work, 0.0, nullptr, work); FS_LAUNCH_CHECK();
  }
  double d2[2];
  dot2(n, work, work, nullptr, nullptr, a->partials.p, d2);
  return d2[0];
}

// ---- small systems: the whole (Jacobi right-preconditioned) BiCGStab in ONE CTA --------------
// The Poisson / heat systems of the shipped meshes have a few hundred rows: with one launch per
// vector operation every iteration is pure launch latency.  Here one CTA of 1024 threads runs
// the complete solve (rows strided over the threads, block-level reductions, no host reads).
constexpr int kSmallN = 16384;

__device__ __forceinline__ double cta_sum(double v, double* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();                      // red may still be read from the previous reduction
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
  return s;
}

__global__ void __launch_bounds__(1024)
k_bicgstab_small(CsrView A, const double* __restrict__ dinv, const double* __restrict__ b, double* __restrict__ x,
                 double* __restrict__ ws, double tol2, int maxit, double* __restrict__ out /* {rr, bb} */,
                 int* __restrict__ flags /* {converged, iterations} */) {
  __shared__ double red[32];
  const int n = A.n, t = threadIdx.x, nt = blockDim.x;
  double *r = ws, *r0 = r + n, *p = r0 + n, *v = p + n, *s = v + n, *tt = s + n, *ph = tt + n, *sh = ph + n;
  auto spmv = [&](const double* in, double* outv) {
    for (int i = t; i < n; i += nt) {
      double acc = 0.0;
      for (int k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) acc += A.vals[k] * in[A.colidx[k]];
      outv[i] = acc;
    }
    __syncthreads();
  };
  spmv(x, v);
  double lb = 0.0, lr = 0.0;
  for (int i = t; i < n; i += nt) {
    const double ri = b[i] - v[i];
    r[i] = ri; r0[i] = ri; p[i] = 0.0; v[i] = 0.0;
    lb += b[i] * b[i]; lr += ri * ri;
  }
  const double bb = cta_sum(lb, red);
  double rr = cta_sum(lr, red);
  double rho = 1.0, alpha = 1.0, omega = 1.0;
  int it = 0;
  bool broke = false;
  if (bb > 0.0) {
    while (rr > tol2 * bb && it < maxit) {
      double l = 0.0;
      for (int i = t; i < n; i += nt) l += r0[i] * r[i];
      const double rho_new = cta_sum(l, red);
      if (rho_new == 0.0) { broke = true; break; }
      const double beta = (rho_new / rho) * (alpha / omega);
      for (int i = t; i < n; i += nt) {
        const double pi = r[i] + beta * (p[i] - omega * v[i]);
        p[i] = pi;
        ph[i] = dinv ? dinv[i] * pi : pi;
      }
      __syncthreads();
      spmv(ph, v);
      l = 0.0;
      for (int i = t; i < n; i += nt) l += r0[i] * v[i];
      const double r0v = cta_sum(l, red);
      if (r0v == 0.0) { broke = true; break; }
      alpha = rho_new / r0v;
      for (int i = t; i < n; i += nt) {
        const double si = r[i] - alpha * v[i];
        s[i] = si;
        sh[i] = dinv ? dinv[i] * si : si;
      }
      __syncthreads();
      spmv(sh, tt);
      double l1 = 0.0, l2 = 0.0;
      for (int i = t; i < n; i += nt) { l1 += tt[i] * s[i]; l2 += tt[i] * tt[i]; }
      const double ts = cta_sum(l1, red), t2 = cta_sum(l2, red);
      omega = (t2 != 0.0) ? ts / t2 : 0.0;
      l = 0.0;
      for (int i = t; i < n; i += nt) {
        x[i] += alpha * ph[i] + omega * sh[i];
        const double ri = s[i] - omega * tt[i];
        r[i] = ri;
        l += ri * ri;
      }
      rr = cta_sum(l, red);
      rho = rho_new;
      ++it;
      if (omega == 0.0) { broke = true; break; }
    }
  } else {
    for (int i = t; i < n; i += nt) x[i] = 0.0;
    rr = 0.0;
  }
  if (t == 0) {
    out[0] = rr; out[1] = bb;
    flags[0] = (bb == 0.0 || rr <= tol2 * bb) ? 1 : 0;
    flags[1] = it;
    (void)broke;
  }
}

static int bicgstab_small(fs_csr* a, const double* d_b, double* x, double rtol, int maxit, int precond, double* relres) {
  const int64_t n = a->n;
  ensure_ws(a, 8 * (size_t)n);
  const double* dinv = nullptr;
  if (precond == FS_PRECOND_JACOBI) { jacobi_prepare(a); dinv = a->dinv.p; }
  cudaStream_t st = stream();
  double* out = a->scal.p;
  int* flags = reinterpret_cast<int*>(a->scal.p + 32);
  k_bicgstab_small<<<1, 1024, 0, st>>>(a->view(), dinv, d_b, x, a->ws.p, rtol * rtol, maxit, out, flags);
  FS_LAUNCH_CHECK();
  double ho[2];
  int hf[2];
  FS_CUDA(cudaMemcpyAsync(ho, out, sizeof(ho), cudaMemcpyDeviceToHost, st));
  FS_CUDA(cudaMemcpyAsync(hf, flags, sizeof(hf), cudaMemcpyDeviceToHost, st));
  FS_CUDA(cudaStreamSynchronize(st));
  if (relres) *relres = ho[1] > 0.0 ? std::sqrt(ho[0] / ho[1]) : 0.0;
  return hf[0] ? hf[1] : -hf[1] - 1;
}

static int bicgstab_impl(fs_csr* a, const double* d_b, double* x, double rtol, int maxit, int precond, double* relres) {
  const int64_t n = a->n;
  if (n <= kSmallN) return bicgstab_small(a, d_b, x, rtol, maxit, precond, relres);
  ensure_ws(a, 8 * (size_t)n);
  double *r = a->ws.p, *r0 = r + n, *p = r0 + n, *v = p + n, *s = v + n, *t = s + n, *ph = t + n, *sh = ph + n;
  const double* dinv = nullptr;
  if (precond == FS_PRECOND_JACOBI) { jacobi_prepare(a); dinv = a->dinv.p; }
  ensure_tiles(a);
  const CsrView A = a->view();
  cudaStream_t st = stream();
  const int g = vec_grid(n);
  double* part = a->partials.p;
  double d2[2];
  spmv_dev(A, x, v);
  k_lin3<<<g, kBlock, 0, st>>>(n, 1.0, d_b, -1.0, v, 0.0, nullptr, r); FS_LAUNCH_CHECK();
  FS_CUDA(cudaMemcpyAsync(r0, r, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
  dot2(n, d_b, d_b, r, r, part, d2);
  const double bb = d2[0];
  double rr = d2[1];
  if (bb == 0.0) { FS_CUDA(cudaMemsetAsync(x, 0, n * sizeof(double), st)); if (relres) *relres = 0.0; return 0; }
  const double tol2 = rtol * rtol;
  double rho = 1.0, alpha = 1.0, omega = 1.0;
  FS_CUDA(cudaMemsetAsync(p, 0, n * sizeof(double), st));
  FS_CUDA(cudaMemsetAsync(v, 0, n * sizeof(double), st));
  int it = 0;
  while (rr > tol2 * bb && it < maxit) {
    dot2(n, r0, r, nullptr, nullptr, part, d2);
    double rho_new = d2[0];
    if (rho_new == 0.0) break;   // breakdown
    double beta = (rho_new / rho) * (alpha / omega);
    // p = r + beta*(p - omega*v)
    k_lin3<<<g, kBlock, 0, st>>>(n, 1.0, r, beta, p, -beta * omega, v, p); FS_LAUNCH_CHECK();
    k_mul<<<g, kBlock, 0, st>>>(n, dinv, p, ph); FS_LAUNCH_CHECK();
    spmv_dev(A, ph, v);
    dot2(n, r0, v, nullptr, nullptr, part, d2);
    if (d2[0] == 0.0) break;
    alpha = rho_new / d2[0];
    k_lin3<<<g, kBlock, 0, st>>>(n, 1.0, r, -alpha, v, 0.0, nullptr, s); FS_LAUNCH_CHECK();
    k_mul<<<g, kBlock, 0, st>>>(n, dinv, s, sh); FS_LAUNCH_CHECK();
    spmv_dev(A, sh, t);
    dot2(n, t, s, t, t, part, d2);
    omega = (d2[1] != 0.0) ? d2[0] / d2[1] : 0.0;
    // x += alpha*ph + omega*sh ; r = s - omega*t
    k_lin3<<<g, kBlock, 0, st>>>(n, 1.0, x, alpha, ph, omega, sh, x); FS_LAUNCH_CHECK();
    k_lin3<<<g, kBlock, 0, st>>>(n, 1.0, s, -omega, t, 0.0, nullptr, r); FS_LAUNCH_CHECK();
    dot2(n, r, r, nullptr, nullptr, part, d2);
    rr = d2[0];
    rho = rho_new;
    ++it;
    if (omega == 0.0) break;
  }
  if (relres) *relres = std::sqrt(rr / bb);
  return (rr <= tol2 * bb) ? it : -it - 1;
}

}  // namespace fs

using namespace fs;

extern "C" {

int fs_profile(int every) {
  FS_API_BEGIN
  FS_REQUIRE(every >= 0, "every must be >= 0");
  g_prof.every = every;
  g_prof.used = 0;
  g_prof.ms[0] = g_prof.ms[1] = g_prof.ms[2] = 0.0;
  g_prof.samples = 0;
  g_prof.iters = 0;
  g_prof.top_ms = 0.0;
  g_prof.top_samples = 0;
  if (every > 0 && g_prof.ev.empty()) {
    g_prof.ev.resize(640);
    for (auto& e : g_prof.ev) FS_CUDA(cudaEventCreate(&e));
  }
  FS_API_END
}

int fs_profile_read(double* ms3, int64_t* samples, int64_t* iters) {
  FS_API_BEGIN
  if (ms3) for (int j = 0; j < 3; ++j) ms3[j] = g_prof.ms[j];
  if (samples) *samples = g_prof.samples;
  if (iters) *iters = g_prof.iters;
  FS_API_END
}

int fs_profile_read_top(double* ms, int64_t* samples, double* bytes_per_launch) {
  FS_API_BEGIN
  if (ms) *ms = g_prof.top_ms;
  if (samples) *samples = g_prof.top_samples;
  if (bytes_per_launch) *bytes_per_launch = g_prof.top_bytes;
  FS_API_END
}

int fs_csr_create(int64_t n, int64_t nnz, const int32_t* rowptr, const int32_t* colidx, const double* vals, fs_csr** out) {
  FS_API_BEGIN
  FS_REQUIRE(out, "out is NULL");
  *out = nullptr;
  FS_REQUIRE(n > 0 && nnz >= 0 && rowptr && (nnz == 0 || (colidx && vals)), "bad CSR arguments");
  std::unique_ptr<fs_csr> a(new fs_csr());
  a->n = n; a->nnz = nnz;
  a->rowptr_own.alloc(n + 1); a->rowptr_own.upload(rowptr, n + 1);
  a->colidx_own.alloc(nnz); a->colidx_own.upload(colidx, nnz);
  a->vals.alloc(nnz); a->vals.upload(vals, nnz);
  a->rowptr = a->rowptr_own.p; a->colidx = a->colidx_own.p;
  fs::sync();
  *out = a.release();
  FS_API_END
}

int fs_csr_from_mesh(fs_mesh* m, const double* vals, fs_csr** out) {
  FS_API_BEGIN
  FS_REQUIRE(m && vals && out, "NULL argument");
  std::unique_ptr<fs_csr> a(new fs_csr());
  a->n = m->N; a->nnz = m->pat.nnz;
  a->rowptr = m->pat.rowptr.p; a->colidx = m->pat.colidx.p;   // borrowed: mesh must outlive the matrix
  a->vals.alloc(a->nnz); a->vals.upload(vals, a->nnz);
  fs::sync();
  *out = a.release();
  FS_API_END
}

int fs_csr_destroy(fs_csr* a) {
  FS_API_BEGIN
  if (a) { cudaStreamSynchronize(stream()); delete a; }
  FS_API_END
}

int fs_csr_sizes(const fs_csr* a, int64_t* n, int64_t* nnz) {
  FS_API_BEGIN
  FS_REQUIRE(a, "matrix is NULL");
  if (n) *n = a->n;
  if (nnz) *nnz = a->nnz;
  FS_API_END
}

int fs_csr_get(const fs_csr* a, int32_t* rowptr, int32_t* colidx, double* vals) {
  FS_API_BEGIN
  FS_REQUIRE(a, "matrix is NULL");
  cudaStream_t st = stream();
  if (rowptr) FS_CUDA(cudaMemcpyAsync(rowptr, a->rowptr, (a->n + 1) * sizeof(int), cudaMemcpyDefault, st));
  if (colidx && a->nnz) FS_CUDA(cudaMemcpyAsync(colidx, a->colidx, a->nnz * sizeof(int), cudaMemcpyDefault, st));
  if (vals && a->nnz) FS_CUDA(cudaMemcpyAsync(vals, a->vals.p, a->nnz * sizeof(double), cudaMemcpyDefault, st));
  fs::sync();
  FS_API_END
}

int fs_spmv(fs_csr* a, const double* x, double* y) {
  FS_API_BEGIN
  FS_REQUIRE(a && x && y, "NULL argument");
  In<double> ix(x, a->n);
  Out<double> oy(y, a->n);
  ensure_tiles(a);
  spmv_dev(a->view(), ix.d, oy.d);
  oy.commit();
  fs::sync();
  FS_API_END
}

int fs_precond_bytes(fs_csr* a, double* bytes_per_apply) {
  FS_API_BEGIN
  FS_REQUIRE(a && bytes_per_apply, "NULL argument");
  if (!a->amg) a->amg = amg_setup(a);
  *bytes_per_apply = amg_cycle_bytes(a->amg, pcg_use_r32());
  FS_API_END
}

int fs_precond_apply(fs_csr* a, const double* r, double* z) {
  FS_API_BEGIN
  FS_REQUIRE(a && r && z, "NULL argument");
  FS_REQUIRE(a->n > 0, "empty matrix");
  In<double> ir(r, a->n);
  Out<double> oz(z, a->n);
  if (!a->amg) a->amg = amg_setup(a);
  ensure_tiles(a);
  DBuf<float> r32;
  if (pcg_use_r32()) {      // the same arithmetic as inside fs_cg
    r32.alloc((size_t)a->n);
    k_mirror_f32<<<vec_grid(a->n), kBlock, 0, stream()>>>(a->n, ir.d, r32.p);
    FS_LAUNCH_CHECK();
  }
  amg_apply(a->amg, ir.d, oz.d, false, nullptr, nullptr, r32.p);
  oz.commit();
  fs::sync();
  FS_API_END
}

int fs_cg(fs_csr* a, const double* b, double* x, int nrhs, double rtol, int maxit, int precond, int project_mean,
          int* iters, double* relres) {
  FS_API_BEGIN
  FS_REQUIRE(a && b && x, "NULL argument");
  FS_REQUIRE(nrhs == 1 || nrhs == 2, "nrhs must be 1 or 2");
  In<double> ib(b, a->n * nrhs);
  Out<double> ox(x, a->n * nrhs, true);
  double rr = 0.0;
  int it = cg_dev(a, ib.d, ox.d, nrhs, rtol, maxit, precond, project_mean, &rr);
  ox.commit();
  fs::sync();
  if (iters) *iters = it >= 0 ? it : -it - 1;
  if (relres) *relres = rr;
  if (it < 0) throw Error(FS_ERR_NOCONV, "fs_cg: no convergence within maxit");
  FS_API_END
}

int fs_bicgstab(fs_csr* a, const double* b, double* x, double rtol, int maxit, int precond, int* iters, double* relres) {
  FS_API_BEGIN
  FS_REQUIRE(a && b && x, "NULL argument");
  In<double> ib(b, a->n);
  Out<double> ox(x, a->n, true);
  double rr = 0.0;
  int it = bicgstab_impl(a, ib.d, ox.d, rtol, maxit, precond, &rr);
  ox.commit();
  fs::sync();
  if (iters) *iters = it >= 0 ? it : -it - 1;
  if (relres) *relres = rr;
  if (it < 0) throw Error(FS_ERR_NOCONV, "fs_bicgstab: no convergence (maxit or breakdown)");
  FS_API_END
}

}  // extern "C"
