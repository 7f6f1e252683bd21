// spmv_warp.cu -- the warp-granular CSR-stream SpMV of the persistent CG kernel (pass A of
// cg_persistent.cu) as a stand-alone kernel with fused epilogues: the unfolded AMG V-cycle
// (FS_AMG_FOLD=0), the folded cycle without SELL copies (FS_AMG_SELL=0), initial residuals of the
// Krylov solvers.  The default V-cycle and the AMG-PCG's A*p run on spmv_sell.cu instead.
//   EPI_AX      y  = A x
//   EPI_RESID   y  = b - A x
//   EPI_JACOBI  y  = x + w D^-1 (b - A x)                      (y must not alias x)
//   EPI_PRESM   x0 = w D^-1 b (stored to xout),  y = b - A x0   (pre-smoothing from a zero guess
//               and the residual in one pass: the gather reads dinv[col]*b[col])
//   EPI_ADD     y += A x
//   EPI_AX2     y  = A x,  xout = w D^-1 y                       (restriction that also pre-smooths)
//   EPI_AXS     y  = A [x; x2]: columns >= nsplit gather from x2 (the folded V-cycle's
//               up-sweep operator [G | P~] applied to [b; x_coarse] without concatenating them)
// plus an optional fused dot product x.(A x) (per-CTA partials) for the CG.
// k_spmv_sub is the fallback for matrices whose 32-row mini-tiles do not fit shared memory
// (very long rows): LPR lanes per row, same gather options.
// A warp owns 32-row mini-tiles; values/columns of the next mini-tile are in flight in registers
// while the current one is multiplied and reduced through the warp's shared-memory slice.
#include "dist.cuh"

namespace fs {

constexpr int kWT = 512;        // threads per CTA
constexpr int kWW = kWT / 32;   // warps per CTA
constexpr int kWU = 8;          // prefetched nonzeros per lane (window of 256 per mini-tile)

__device__ __forceinline__ uint64_t w_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ double w_ld_stream_f64(const double* a, uint64_t pol) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(a), "l"(pol));
  return v;
}
__device__ __forceinline__ int w_ld_stream_s32(const int* a, uint64_t pol) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(pol));
  return v;
}

struct SpmvWarpArgs {
  CsrView A;
  const double* x;
  double* y;
  const double* b;
  const double* dinv;
  double w;
  double* xout;
  double* part;   // per-CTA partial of x.(A x) when DOT
  int wcap;
  int ntiles;     // 512-row tiles
  const double* x2 = nullptr;   // EPI_AXS: second gather source
  int nsplit = 0;               // EPI_AXS: first column served by x2
  float* yf = nullptr;          // k_spmv_sub: optional fp32 mirror of y (gather source of the packed SELL kernels)
};

__device__ __forceinline__ float w_ld_stream_f32(const float* a, uint64_t pol) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(a), "l"(pol));
  return v;
}

// F32: the matrix values are streamed from the fp32 copy (8 instead of 12 bytes per nonzero);
// vectors and accumulation stay fp64.  Used for the AMG V-cycle, which is only a preconditioner.
template <int EPI, bool DOT, bool F32>
__global__ void __launch_bounds__(kWT, 2) k_spmv_warp(SpmvWarpArgs a) {
  extern __shared__ double prod[];
  __shared__ double red[kWW];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nb = gridDim.x, bi = blockIdx.x;
  const uint64_t pf = w_evict_first();
  const int tile0 = (int)(((long long)a.ntiles * bi) / nb);
  const int tile1 = (int)(((long long)a.ntiles * (bi + 1)) / nb);
  const int R0 = tile0 * kWT, R1 = min(a.A.n, tile1 * kWT);
  const int* __restrict__ rowptr = a.A.rowptr;
  const double* __restrict__ vals = a.A.vals;
  const float* __restrict__ vals32 = a.A.vals32;
  const int* __restrict__ colidx = a.A.colidx;
  const unsigned full = 0xffffffffu;
  double* pw = prod + (size_t)warp * a.wcap;
  const int nmt = (R1 - R0 + 31) >> 5;
  double acc = 0.0;
  double va[kWU];
  int ca[kWU];
  auto xval = [&](int j) -> double {
    if (EPI == EPI_PRESM) return a.w * __ldg(a.dinv + j) * __ldg(a.b + j);
    if (EPI == EPI_AXS) return j < a.nsplit ? __ldg(a.x + j) : __ldg(a.x2 + (j - a.nsplit));
    return __ldg(a.x + j);
  };
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
    for (int j = 0; j < kWU; ++j) {
      const int k = (j << 5) + lane;
      const bool ok = k < cnt;
      va[j] = ok ? (F32 ? (double)w_ld_stream_f32(vals32 + base + k, pf) : w_ld_stream_f64(vals + base + k, pf)) : 0.0;
      ca[j] = ok ? w_ld_stream_s32(colidx + base + k, pf) : -1;
    }
  };
  int rpA, endA, rpB, endB;
  load_rp(warp, rpA, endA);
  load_rp(warp + kWW, rpB, endB);
  int baseA = __shfl_sync(full, rpA, 0);
  int cntA = endA - baseA;
  stream(baseA, cntA);
  for (int mt = warp; mt < nmt; mt += kWW) {
    // the mini-tile's nonzeros in chunks of 256 (kWU per lane); the registers always hold the chunk
    // that is multiplied next, the following one (of this mini-tile or the first of the next) is
    // requested before the current one is consumed any further
    int rpC, endC;
    load_rp(mt + 2 * kWW, rpC, endC);
    const int baseB = __shfl_sync(full, rpB, 0);
    const int cntB = endB - baseB;
    constexpr int kChunk = kWU << 5;
    for (int c0 = 0;; c0 += kChunk) {
#pragma unroll
      for (int j = 0; j < kWU; ++j)
        if (ca[j] >= 0) pw[c0 + (j << 5) + lane] = va[j] * xval(ca[j]);
      const bool last = c0 + kChunk >= cntA;
      if (last) stream(baseB, cntB);                         // first chunk of the next mini-tile
      else stream(baseA + c0 + kChunk, cntA - c0 - kChunk);  // next chunk of this one
      if (last) break;
    }
    __syncwarp();
    int nxt = __shfl_down_sync(full, rpA, 1);
    if (lane == 31) nxt = endA;
    const int row = R0 + (mt << 5) + lane;
    if (row < R1) {
      double s = 0.0;
      for (int k = rpA - baseA; k < nxt - baseA; ++k) s += pw[k];
      if (DOT) acc += __ldg(a.x + row) * s;
      if (EPI == EPI_AX || EPI == EPI_AXS) a.y[row] = s;
      else if (EPI == EPI_RESID) a.y[row] = a.b[row] - s;
      else if (EPI == EPI_JACOBI) a.y[row] = a.x[row] + a.w * a.dinv[row] * (a.b[row] - s);
      else if (EPI == EPI_PRESM) { const double bv = a.b[row]; a.xout[row] = a.w * a.dinv[row] * bv; a.y[row] = bv - s; }
      else if (EPI == EPI_ADD) a.y[row] += s;
      else if (EPI == EPI_AX2) { a.y[row] = s; a.xout[row] = a.w * a.dinv[row] * s; }
    }
    __syncwarp();
    rpA = rpB; endA = endB; baseA = baseB; cntA = cntB;
    rpB = rpC; endB = endC;
  }
  if (DOT) {
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(full, acc, o);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (t == 0) {
      double s = 0.0;
      for (int k = 0; k < kWW; ++k) s += red[k];
      a.part[bi] = s;
    }
  }
}

// LPR lanes per row (4..32), lane-strided partial sums combined by an xor tree: deterministic, any
// row length; y = A x or A [x; x2]
// DIST (partitioned cycle, small levels): wait for the input channels' halo flags first, store the rows other ranks
// read into their halo slots as they are produced, and let the last CTA release the output channel's flags --
// one launch instead of wait kernel + SpMV + push kernel.
struct DistSub {
  Comm c;
  HaloWait w;
  PushSpec ps;
};

template <int LPR, bool F32, bool DIST>
__global__ void __launch_bounds__(256) k_spmv_sub(SpmvWarpArgs a, DistSub d) {
  if (!DIST) pdl_launch();
  if (DIST) {
    if (d.c.done && *d.c.done) return;
    halo_wait(d.c, d.w);
  }
  const int row = (int)((blockIdx.x * 256ll + threadIdx.x) / LPR), sub = threadIdx.x % LPR;
  double s = 0.0;
  if (row < a.A.n) {
    const int rs = __ldg(a.A.rowptr + row), re = __ldg(a.A.rowptr + row + 1);
    // four strided entries per lane in flight: loads first, then the gathers, then the sums
    for (int k0 = rs + sub; k0 < re; k0 += 4 * LPR) {
      double v[4];
      int c[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + j * LPR;
        const bool ok = k < re;
        v[j] = ok ? (F32 ? (double)__ldg(a.A.vals32 + k) : __ldg(a.A.vals + k)) : 0.0;
        c[j] = ok ? __ldg(a.A.colidx + k) : 0;
      }
      // programmatic dependent launch: the row's entries are on their way before the producer of x has finished
      if (!DIST) pdl_wait();
      double xx[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) xx[j] = c[j] < a.nsplit ? __ldg(a.x + c[j]) : __ldg(a.x2 + (c[j] - a.nsplit));
#pragma unroll
      for (int j = 0; j < 4; ++j) s += v[j] * xx[j];
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (!DIST) pdl_wait();      // rows without entries have not waited yet: no write before the predecessor is done
  if (row < a.A.n && sub == 0) {
    a.y[row] = s;
    if (a.yf) a.yf[row] = (float)s;
    if (DIST && d.ps.enabled) push_row(d.ps, row, s);
  }
  if (DIST && d.ps.enabled) push_finish(d.c, d.ps, true, dist_seq(d.c), (int)gridDim.x);
}

template <int LPR>
static void launch_sub(const SpmvWarpArgs& args, const DistSub* d = nullptr) {
  const int grid = (int)div_up((int64_t)args.A.n * LPR, 256);
  static const DistSub none{};
  if (d) {
    if (args.A.vals32) k_spmv_sub<LPR, true, true><<<grid, 256, 0, stream()>>>(args, *d);
    else k_spmv_sub<LPR, false, true><<<grid, 256, 0, stream()>>>(args, *d);
  } else {
    if (args.A.vals32) launch_pdl(k_spmv_sub<LPR, true, false>, grid, 256, 0, args, none);
    else launch_pdl(k_spmv_sub<LPR, false, false>, grid, 256, 0, args, none);
  }
  FS_LAUNCH_CHECK();
}

// y = A [x; x2] (x2 may be null: plain y = A x) for any CSR matrix.  Lanes per row: about a quarter of
// the mean row length (each lane keeps four entries in flight), so that small matrices still
// spread over the whole machine in one wave.
static void spmv_sub_impl(const CsrView& A, const double* x, double* y, const double* x2, int nsplit, const DistSub* d,
                          float* yf = nullptr) {
  SpmvWarpArgs args{A, x, y, nullptr, nullptr, 0.0, nullptr, nullptr, 0, 0};
  args.yf = yf;
  args.x2 = x2;
  args.nsplit = x2 ? nsplit : 0x7fffffff;
  static const double per_lane = [] { const char* e = std::getenv("FS_SUB_PER_LANE"); return e ? std::atof(e) : 4.0; }();
  const double avg = A.n > 0 ? (double)A.nnz / A.n : 0.0;
  const double lanes = avg / per_lane;
  if (lanes <= 4.0) launch_sub<4>(args, d);
  else if (lanes <= 8.0) launch_sub<8>(args, d);
  else if (lanes <= 16.0) launch_sub<16>(args, d);
  else launch_sub<32>(args, d);
}
void spmv_sub(const CsrView& A, const double* x, double* y, const double* x2, int nsplit, float* yf) {
  spmv_sub_impl(A, x, y, x2, nsplit, nullptr, yf);
}
void spmv_sub_dist(const CsrView& A, const double* x, double* y, const double* x2, int nsplit, const Comm& c, const HaloWait& w,
                   const PushSpec& ps) {
  DistSub d;
  d.c = c; d.w = w; d.ps = ps;
  spmv_sub_impl(A, x, y, x2, nsplit, &d);
}

static bool g_warp_attr = false;

template <int EPI, bool DOT>
static void launch_one(const SpmvWarpArgs& args, int grid, size_t smem) {
  if (args.A.vals32) k_spmv_warp<EPI, DOT, true><<<grid, kWT, smem, stream()>>>(args);
  else k_spmv_warp<EPI, DOT, false><<<grid, kWT, smem, stream()>>>(args);
}

template <int EPI, bool DOT>
static void set_smem_attr(size_t bytes) {
  FS_CUDA(cudaFuncSetAttribute(k_spmv_warp<EPI, DOT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  FS_CUDA(cudaFuncSetAttribute(k_spmv_warp<EPI, DOT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

// Returns the grid size used (>0), or 0 when the matrix does not fit this kernel (the caller
// falls back to the tile / vector kernels).
int spmv_warp(const CsrView& A, int epi, const double* x, double* y, const double* b, const double* dinv, double w,
              double* xout, double* dot_partials, const double* x2, int nsplit) {
  if (A.wtile_nnz_max <= 0) return 0;
  const int wcap = (A.wtile_nnz_max + 31) / 32 * 32;
  const size_t smem = (size_t)kWW * wcap * sizeof(double);
  const size_t big = 200 * 1024;   // above 100 KB one CTA per SM
  if (smem > big) return 0;
  if (!g_warp_attr) {
    set_smem_attr<EPI_AX, false>(big);
    set_smem_attr<EPI_AX, true>(big);
    set_smem_attr<EPI_RESID, false>(big);
    set_smem_attr<EPI_JACOBI, false>(big);
    set_smem_attr<EPI_PRESM, false>(big);
    set_smem_attr<EPI_ADD, false>(big);
    set_smem_attr<EPI_AX2, false>(big);
    set_smem_attr<EPI_AXS, false>(big);
    set_smem_attr<EPI_AXS, true>(big);
    g_warp_attr = true;
  }
  const int ntiles = div_up(A.n, kWT);
  const int per_sm = smem > 100 * 1024 ? 1 : 2;
  const int grid = std::max(1, std::min(sm_count() * per_sm, ntiles));
  SpmvWarpArgs args{A, x, y, b, dinv, w, xout, dot_partials, wcap, ntiles};
  args.x2 = x2;
  args.nsplit = nsplit;
  const bool dot = dot_partials != nullptr;
  switch (epi) {
    case EPI_AX: if (dot) launch_one<EPI_AX, true>(args, grid, smem); else launch_one<EPI_AX, false>(args, grid, smem); break;
    case EPI_RESID: launch_one<EPI_RESID, false>(args, grid, smem); break;
    case EPI_JACOBI: launch_one<EPI_JACOBI, false>(args, grid, smem); break;
    case EPI_PRESM: launch_one<EPI_PRESM, false>(args, grid, smem); break;
    case EPI_ADD: launch_one<EPI_ADD, false>(args, grid, smem); break;
    case EPI_AX2: launch_one<EPI_AX2, false>(args, grid, smem); break;
    case EPI_AXS: if (dot) launch_one<EPI_AXS, true>(args, grid, smem); else launch_one<EPI_AXS, false>(args, grid, smem); break;
    default: throw Error(FS_ERR_INTERNAL, "spmv_warp: bad epilogue");
  }
  FS_LAUNCH_CHECK();
  return grid;
}

}  // namespace fs
