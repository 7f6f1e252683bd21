// cg_persistent.cu -- the whole (Jacobi-)preconditioned CG solve as ONE cooperative,
// persistent kernel: 2 CTAs of 512 threads per SM, each owning a contiguous block of
// rows for all three passes of an iteration; grid-wide barriers between the passes;
// convergence is decided on the device, the host is not involved until the end.
//
//   pass A  Ap = A p (+ p.Ap)   warp-granular CSR-stream: a warp owns 32-row mini-tiles;
//           the mini-tile's values/columns are streamed with coalesced evict-first loads
//           issued one mini-tile ahead (register prefetch, in flight while the previous
//           one is reduced; row pointers two ahead), products parked in the warp's
//           shared-memory slice, each lane sums one row in column order.  No block barrier.
//   pass B  x += a p ; r -= a Ap (+ r.r, r.z with z = Dinv r)
//   pass C  p = z + b p
// Dot products: per-CTA partials, re-reduced by every CTA after the barrier in a fixed
// order => deterministic, no atomics.  The CG vectors carry an L2 evict-last policy,
// the matrix stream evict-first, so the 5 vectors (8N bytes each) can stay L2-resident.
#include <cooperative_groups.h>

#include "internal.cuh"

namespace cg = cooperative_groups;

namespace fs {

constexpr int kPT = 512;       // threads per CTA == rows per CTA tile (row-block granularity)
constexpr int kPW = kPT / 32;  // warps per CTA
constexpr int kPU = 8;         // prefetched nonzeros per lane per 32-row mini-tile (window = 256 nonzeros)
constexpr int kPVU = 2;        // unroll of the vector passes (64-register budget: 2 CTAs x 512 threads per SM)

__device__ __forceinline__ uint64_t p_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t p_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// read-only stream (matrix): non-coherent path, no L1 allocation, L2 evict-first
__device__ __forceinline__ double ld_stream_f64(const double* a, uint64_t pol) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(a), "l"(pol));
  return v;
}
__device__ __forceinline__ int ld_stream_s32(const int* a, uint64_t pol) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(pol));
  return v;
}
// CG vectors: coherent loads (they are rewritten between barriers), L2 evict-last
__device__ __forceinline__ double ld_vec(const double* a, uint64_t pol) {
  double v;
  asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(a), "l"(pol) : "memory");
  return v;
}
__device__ __forceinline__ void st_vec(double* a, double v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(a), "d"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

struct CgPersistArgs {
  CsrView A;
  double* x;
  double* r;
  double* p;
  double* Ap;
  const double* dinv;   // nullable
  double* partA;        // gridDim
  double* partB;        // 2*gridDim
  double* scal;         // [0]=rz (in/out) [2]=bb (in) [3]=rr (out) [8..10]=ns per pass [11]=timed iterations
  int* flags;           // [0]=converged [1]=iterations
  int maxit;
  double tol2;
  int ntiles;
  int wcap;             // shared-memory doubles per warp (>= max nonzeros of a 32-row mini-tile)
};

template <int K>
__device__ __forceinline__ void blk_reduce(double (&v)[K], double* red /* K*16 */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < K; ++k) red[k * 16 + warp] = v[k];
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double t = (lane < kPT / 32) ? red[k * 16 + lane] : 0.0;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      v[k] = t;
    }
  }
  __syncthreads();
}

// fixed-order sum of the per-CTA partials (nblk x K), identical in every CTA
template <int K>
__device__ __forceinline__ void sum_partials(const double* part, int nblk, double (&out)[K], double* sm /* K */) {
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double t = 0.0;
      for (int b = lane; b < nblk; b += 32) t += __ldcg(part + (size_t)b * K + k);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (lane == 0) sm[k] = t;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) out[k] = sm[k];
  __syncthreads();
}

__global__ void __launch_bounds__(kPT, 2) k_cg_persistent(CgPersistArgs a) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ double prod[];
  __shared__ double red[2 * 16];
  __shared__ double sm[2];
  const int t = threadIdx.x, nb = gridDim.x, b = blockIdx.x;
  const int lane = t & 31, warp = t >> 5;
  const uint64_t pf = p_evict_first(), pl = p_evict_last();
  const int tile0 = (int)(((long long)a.ntiles * b) / nb);
  const int tile1 = (int)(((long long)a.ntiles * (b + 1)) / nb);
  const int R0 = tile0 * kPT, R1 = min(a.A.n, tile1 * kPT);
  const int* __restrict__ rowptr = a.A.rowptr;
  const double* __restrict__ vals = a.A.vals;
  const int* __restrict__ colidx = a.A.colidx;
  const double* __restrict__ dinv = a.dinv;
  double rz_old = a.scal[0];
  const double bb = a.scal[2];
  double rr = a.scal[3];
  int it = 0;
  bool converged = false;
  unsigned long long t0 = 0, nsA = 0, nsB = 0, nsC = 0;
  const bool timer = (b == 0 && t == 0);
  if (timer) t0 = gtime();

  while (it < a.maxit) {
    // ------------------------------------------------------------------ pass A
    // Warp-granular CSR-stream: every warp owns 32-row mini-tiles (stride 16 inside the
    // CTA's row block) and needs no block barrier.  Three-deep software pipeline per warp:
    // row pointers two mini-tiles ahead, the value/column stream one ahead (registers),
    // gathers + products + row sums on the current one.
    double accA[1] = {0.0};
    {
      const unsigned full = 0xffffffffu;
      double* pw = prod + (size_t)warp * a.wcap;
      const int nmt = (R1 - R0 + 31) >> 5;
      double va[kPU];
      int ca[kPU];
      auto load_rp = [&](int mt, int& rp, int& rend) {
        if (mt < nmt) {
          const int r0 = R0 + (mt << 5);
          const int nr = min(32, R1 - r0);
          rp = __ldg(rowptr + r0 + min(lane, nr));
          rend = __ldg(rowptr + r0 + nr);
        } else { rp = 0; rend = 0; }
      };
      auto stream = [&](int base, int cnt) {
#pragma unroll
        for (int j = 0; j < kPU; ++j) {
          const int k = (j << 5) + lane;
          const bool ok = k < cnt;
          va[j] = ok ? ld_stream_f64(vals + base + k, pf) : 0.0;
          ca[j] = ok ? ld_stream_s32(colidx + base + k, pf) : -1;
        }
      };
      int rpA, endA, rpB, endB;
      load_rp(warp, rpA, endA);
      load_rp(warp + kPW, rpB, endB);
      int baseA = __shfl_sync(full, rpA, 0);
      int cntA = endA - baseA;
      stream(baseA, cntA);
      for (int mt = warp; mt < nmt; mt += kPW) {
#pragma unroll
        for (int j = 0; j < kPU; ++j)
          if (ca[j] >= 0) pw[(j << 5) + lane] = va[j] * ld_vec(a.p + ca[j], pl);
        for (int k = (kPU << 5) + lane; k < cntA; k += 32)   // rows beyond the prefetch window (rare)
          pw[k] = ld_stream_f64(vals + baseA + k, pf) * ld_vec(a.p + ld_stream_s32(colidx + baseA + k, pf), pl);
        __syncwarp();
        int rpC, endC;
        load_rp(mt + 2 * kPW, rpC, endC);
        const int baseB = __shfl_sync(full, rpB, 0);
        const int cntB = endB - baseB;
        stream(baseB, cntB);                                  // in flight while this mini-tile is reduced
        int nxt = __shfl_down_sync(full, rpA, 1);
        if (lane == 31) nxt = endA;
        const int row = R0 + (mt << 5) + lane;
        if (row < R1) {
          double s = 0.0;
          for (int k = rpA - baseA; k < nxt - baseA; ++k) s += pw[k];
          st_vec(a.Ap + row, s, pl);
          accA[0] += ld_vec(a.p + row, pl) * s;
        }
        __syncwarp();
        rpA = rpB; endA = endB; baseA = baseB; cntA = cntB;
        rpB = rpC; endB = endC;
      }
    }
    blk_reduce<1>(accA, red);
    if (t == 0) __stcg(a.partA + b, accA[0]);
    grid.sync();
    if (timer) { unsigned long long t1 = gtime(); nsA += t1 - t0; t0 = t1; }

    // ------------------------------------------------------------------ pass B
    double pAp[1];
    sum_partials<1>(a.partA, nb, pAp, sm);
    const double alpha = (pAp[0] != 0.0) ? rz_old / pAp[0] : 0.0;
    double accB[2] = {0.0, 0.0};
    for (int i0 = R0; i0 < R1; i0 += kPT * kPVU) {
      double pv[kPVU], av[kPVU], xv[kPVU], rv[kPVU], dv[kPVU];
#pragma unroll
      for (int j = 0; j < kPVU; ++j) {
        const int i = i0 + j * kPT + t;
        const bool ok = i < R1;
        pv[j] = ok ? ld_vec(a.p + i, pl) : 0.0;
        av[j] = ok ? ld_vec(a.Ap + i, pl) : 0.0;
        xv[j] = ok ? ld_vec(a.x + i, pl) : 0.0;
        rv[j] = ok ? ld_vec(a.r + i, pl) : 0.0;
        dv[j] = (ok && dinv) ? ld_vec(dinv + i, pl) : 1.0;
      }
#pragma unroll
      for (int j = 0; j < kPVU; ++j) {
        const int i = i0 + j * kPT + t;
        if (i < R1) {
          const double xn = xv[j] + alpha * pv[j];
          const double rn = rv[j] - alpha * av[j];
          st_vec(a.x + i, xn, pl);
          st_vec(a.r + i, rn, pl);
          accB[0] += rn * rn;
          accB[1] += rn * (dv[j] * rn);
        }
      }
    }
    blk_reduce<2>(accB, red);
    if (t == 0) { __stcg(a.partB + 2 * b, accB[0]); __stcg(a.partB + 2 * b + 1, accB[1]); }
    grid.sync();
    if (timer) { unsigned long long t1 = gtime(); nsB += t1 - t0; t0 = t1; }

    double sB[2];
    sum_partials<2>(a.partB, nb, sB, sm);
    rr = sB[0];
    ++it;
    const double rz_new = sB[1];
    if (rr <= a.tol2 * bb) { converged = true; rz_old = rz_new; break; }
    const double beta = (rz_old != 0.0) ? rz_new / rz_old : 0.0;
    rz_old = rz_new;

    // ------------------------------------------------------------------ pass C
    for (int i0 = R0; i0 < R1; i0 += kPT * kPVU) {
      double rv[kPVU], pv[kPVU], dv[kPVU];
#pragma unroll
      for (int j = 0; j < kPVU; ++j) {
        const int i = i0 + j * kPT + t;
        const bool ok = i < R1;
        rv[j] = ok ? ld_vec(a.r + i, pl) : 0.0;
        pv[j] = ok ? ld_vec(a.p + i, pl) : 0.0;
        dv[j] = (ok && dinv) ? ld_vec(dinv + i, pl) : 1.0;
      }
#pragma unroll
      for (int j = 0; j < kPVU; ++j) {
        const int i = i0 + j * kPT + t;
        if (i < R1) st_vec(a.p + i, dv[j] * rv[j] + beta * pv[j], pl);
      }
    }
    grid.sync();
    if (timer) { unsigned long long t1 = gtime(); nsC += t1 - t0; t0 = t1; }
  }
  if (timer) {
    a.scal[0] = rz_old;
    a.scal[3] = rr;
    a.scal[8] = (double)nsA;
    a.scal[9] = (double)nsB;
    a.scal[10] = (double)nsC;
    a.scal[11] = (double)it;
    a.flags[0] = converged ? 1 : 0;
    a.flags[1] = it;
  }
}

static int g_persist_blocks_per_sm = -1;

// Returns false when this matrix cannot use the persistent kernel (tile too large for
// shared memory, cooperative launch unsupported); the caller then runs the 3-kernel path.
bool cg_persistent_supported(const CsrView& A, size_t* smem_out) {
  if (A.wtile_nnz_max <= 0) return false;
  const size_t wcap = ((size_t)A.wtile_nnz_max + 31) / 32 * 32;
  const size_t smem = (size_t)kPW * wcap * sizeof(double);
  if (smem > 100 * 1024) return false;
  if (g_persist_blocks_per_sm < 0) {
    int dev = 0, coop = 0;
    FS_CUDA(cudaGetDevice(&dev));
    FS_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    if (!coop) { g_persist_blocks_per_sm = 0; return false; }
    FS_CUDA(cudaFuncSetAttribute(k_cg_persistent, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    g_persist_blocks_per_sm = 1;
  }
  if (g_persist_blocks_per_sm == 0) return false;
  if (smem_out) *smem_out = smem;
  return true;
}

// Launch the persistent solve.  scal/flags as documented in CgPersistArgs.
void cg_persistent_launch(const CsrView& A, double* x, double* r, double* p, double* Ap, const double* dinv,
                          double* partA, double* partB, double* scal, int* flags, int maxit, double tol2) {
  size_t smem = 0;
  FS_REQUIRE(cg_persistent_supported(A, &smem), "persistent CG not supported for this matrix");
  int per_sm = 0;
  FS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cg_persistent, kPT, smem));
  FS_REQUIRE(per_sm >= 1, "persistent CG kernel does not fit on an SM");
  per_sm = std::min(per_sm, 2);
  const int ntiles = div_up(A.n, kPT);
  const int grid = std::max(1, std::min(sm_count() * per_sm, ntiles));
  const int wcap = (A.wtile_nnz_max + 31) / 32 * 32;
  CgPersistArgs args{A, x, r, p, Ap, dinv, partA, partB, scal, flags, maxit, tol2, ntiles, wcap};
  void* kargs[] = {&args};
  FS_CUDA(cudaLaunchCooperativeKernel((void*)k_cg_persistent, dim3(grid), dim3(kPT), kargs, smem, stream()));
  count_launch();
}

}  // namespace fs
