// amg_tail.cu -- the coarse end of the AMG V-cycle as ONE cooperative kernel.
// Below ~100k rows a level's SpMV is a few microseconds of work; as separate launches the
// ~10 kernels of the coarse levels cost more than the fine level they correct.  Here all
// levels from `tail_start` down to the dense coarsest solve and back up run inside one grid,
// separated by grid barriers:
//   down  l:  r_l = b_l - A_l t_l            (t_l = w D^-1 b_l came with b_l)
//             b_{l+1} = P_l^T r_l,  t_{l+1} = w D_{l+1}^-1 b_{l+1}
//   bottom :  x_c = Ainv_c b_c               (dense GEMV, one warp per row)
//   up    l:  t_l += P_l x_{l+1};  x_l = t_l + w D_l^-1 (b_l - A_l t_l)
// The SpMV is the mini-tile scheme of spmv_warp.cu (a warp owns 32 rows, streams their nonzeros
// coalesced, stages the products in shared memory, one lane sums one row in ascending order) with
// a fixed 256-product window, so any row length is handled; sums are bit-identical to
// k_spmv_warp.  Vectors written inside the kernel are read with ld.global.cg (L2), matrices
// and diagonals with the read-only path.
#include <cooperative_groups.h>

#include "internal.cuh"

namespace cg = cooperative_groups;

namespace fs {

constexpr int kTT = 512;        // threads per CTA
constexpr int kTW = kTT / 32;   // warps per CTA
constexpr int kTU = 8;          // products per lane per window
constexpr int kTC = kTU * 32;   // window (products per warp)

__device__ __forceinline__ double tail_val(const TailMat& M, int k) {
  return M.vals32 ? (double)__ldg(M.vals32 + k) : __ldg(M.vals + k);
}

template <int EPI>
__device__ __forceinline__ void tail_spmv(const TailMat& M, const double* x, double* y, const double* b,
                                          const double* __restrict__ dinv, double w, double* xout, double* pw, int gwarp,
                                          int nwarps, int lane) {
  const unsigned full = 0xffffffffu;
  const int nmt = (M.n + 31) >> 5;
  for (int mt = gwarp; mt < nmt; mt += nwarps) {
    const int r0 = mt << 5, nr = min(32, M.n - r0);
    const int rp = __ldg(M.rowptr + r0 + min(lane, nr));
    const int rend = __ldg(M.rowptr + r0 + nr);
    const int base = __shfl_sync(full, rp, 0);
    int nxt = __shfl_down_sync(full, rp, 1);
    if (lane == 31) nxt = rend;
    double s = 0.0;
    for (int c0 = base; c0 < rend; c0 += kTC) {
      double v[kTU];
      int c[kTU];
#pragma unroll
      for (int j = 0; j < kTU; ++j) {
        const int k = c0 + (j << 5) + lane;
        const bool ok = k < rend;
        v[j] = ok ? tail_val(M, k) : 0.0;
        c[j] = ok ? __ldg(M.colidx + k) : -1;
      }
#pragma unroll
      for (int j = 0; j < kTU; ++j)
        if (c[j] >= 0) pw[(j << 5) + lane] = v[j] * __ldcg(x + c[j]);
      __syncwarp();
      const int lo = max(rp, c0) - c0, hi = min(nxt, c0 + kTC) - c0;
      for (int k = lo; k < hi; ++k) s += pw[k];
      __syncwarp();
    }
    if (lane < nr) {
      const int row = r0 + lane;
      if (EPI == EPI_AX) y[row] = s;
      else if (EPI == EPI_RESID) y[row] = __ldcg(b + row) - s;
      else if (EPI == EPI_JACOBI) y[row] = __ldcg(x + row) + w * __ldg(dinv + row) * (__ldcg(b + row) - s);
      else if (EPI == EPI_ADD) y[row] = __ldcg(y + row) + s;
      else if (EPI == EPI_AX2) { y[row] = s; xout[row] = w * __ldg(dinv + row) * s; }
    }
  }
}

__global__ void __launch_bounds__(kTT, 2) k_amg_tail(TailArgs a) {
  __shared__ double win[kTW * kTC];
  cg::grid_group grid = cg::this_grid();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gwarp = blockIdx.x * kTW + warp, nwarps = gridDim.x * kTW;
  double* pw = win + warp * kTC;
  const int m = a.nlev;
  const double w = a.w;
  int ndbg = 0;
  auto mark = [&]() {
    if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0) a.dbg[ndbg++] = clock64();
  };
  mark();
  for (int l = 0; l + 1 < m; ++l) {
    const TailLevel& L = a.lv[l];
    const TailLevel& N = a.lv[l + 1];
    tail_spmv<EPI_RESID>(L.A, L.t, L.r, L.b, nullptr, w, nullptr, pw, gwarp, nwarps, lane);
    mark();
    grid.sync();
    mark();
    if (l + 2 == m) tail_spmv<EPI_AX>(L.PT, L.r, N.bw, nullptr, nullptr, w, nullptr, pw, gwarp, nwarps, lane);
    else tail_spmv<EPI_AX2>(L.PT, L.r, N.bw, nullptr, N.dinv, w, N.t, pw, gwarp, nwarps, lane);
    mark();
    grid.sync();
    mark();
  }
  {
    const TailLevel& C = a.lv[m - 1];
    const int n = C.n;
    const double* __restrict__ Minv = a.Minv;
    for (int row = gwarp; row < n; row += nwarps) {
      double s = 0.0;
      for (int j = lane; j < n; j += 32) s += __ldg(Minv + (size_t)row * n + j) * __ldcg(C.b + j);
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) C.x[row] = s;
    }
  }
  mark();
  grid.sync();
  mark();
  for (int l = m - 2; l >= 0; --l) {
    const TailLevel& L = a.lv[l];
    const TailLevel& N = a.lv[l + 1];
    tail_spmv<EPI_ADD>(L.P, N.x, L.t, nullptr, nullptr, w, nullptr, pw, gwarp, nwarps, lane);
    mark();
    grid.sync();
    mark();
    tail_spmv<EPI_JACOBI>(L.A, L.t, L.x, L.b, L.dinv, w, nullptr, pw, gwarp, nwarps, lane);
    mark();
    if (l > 0) grid.sync();
  }
}

// Grid of the cooperative launch, 0 when the device cannot co-schedule it.
static int tail_grid() {
  static int grid = -1;
  if (grid < 0) {
    int dev = 0, coop = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_amg_tail, kTT, 0);
    const char* e = std::getenv("FS_AMG_TAIL_CTAS");
    const int want = e ? std::max(1, std::atoi(e)) : 1;
    grid = (coop && per_sm >= 1) ? sm_count() * std::min(want, per_sm) : 0;
  }
  return grid;
}

bool amg_tail_supported() { return tail_grid() > 0; }

void amg_tail_launch(const TailArgs& args) {
  const int grid = tail_grid();
  if (grid <= 0) throw Error(FS_ERR_INTERNAL, "amg_tail: cooperative launch unsupported");
  void* kargs[] = {(void*)&args};
  FS_CUDA(cudaLaunchCooperativeKernel((void*)k_amg_tail, dim3(grid), dim3(kTT), kargs, 0, stream()));
  count_launch();
}

}  // namespace fs
