// spmv_sell.cu -- sliced-ELLPACK (SELL-32) copies of the V-cycle operators and their SpMV.
// A slice is 32 consecutive rows padded to the slice's longest row and stored column-major:
// entry k of row (32 s + lane) sits at sptr[s] + 32 k + lane.  One warp owns a slice, lane = row:
//   * every value / column load is a fully coalesced 128-byte request and all loads of a slice are
//     independent (8 per array per lane in flight), nothing is staged in shared memory and there
//     is no reduction phase: the lane accumulates its own row in a register, in ascending column
//     order -- the same order (and bits) as the CSR kernels;
//   * for a fixed k the 32 lanes gather the k-th neighbours of 32 consecutive rows, which on a mesh
//     numbering are close together: far fewer L1 sectors per gather than the CSR-stream layout, where
//     a warp's lanes walk along single rows.
// The CSR-stream kernels measured ~1.0-1.3 nonzeros / cycle / SM whatever the value width (L1TEX
// bound: scattered gathers + shared-memory staging), so fp32 values bought nothing there.
// y = A x or A [x; x2] (columns >= nsplit read x2), optional per-CTA partials of x.y.
#include <cub/cub.cuh>

#include "internal.cuh"

namespace fs {

constexpr int kST = 256;        // threads per CTA
constexpr int kSW = kST / 32;   // warps (slices in flight) per CTA
constexpr int kSU = 8;          // entries per lane per batch

struct SellArgs {
  int n, nslices;
  const long long* sptr;
  const int* cols;
  const float* v32;
  const double* v64;
  const double* x;
  const double* x2;
  int nsplit;
  double* y;
  double* part;
};

__device__ __forceinline__ uint64_t s_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ float s_ld_f32(const float* a, uint64_t pol) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(a), "l"(pol));
  return v;
}
__device__ __forceinline__ double s_ld_f64(const double* a, uint64_t pol) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(a), "l"(pol));
  return v;
}
__device__ __forceinline__ int s_ld_s32(const int* a, uint64_t pol) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(pol));
  return v;
}

template <bool SPLIT, bool DOT, bool F32>
__global__ void __launch_bounds__(kST, 6) k_spmv_sell(SellArgs a) {
  __shared__ double red[kSW];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = gridDim.x * kSW;
  const uint64_t pf = s_evict_first();
  double dacc = 0.0;
  for (int s = blockIdx.x * kSW + warp; s < a.nslices; s += nwarps) {
    const long long off = __ldg(a.sptr + s);
    const int W = (int)((__ldg(a.sptr + s + 1) - off) >> 5);
    const int* __restrict__ c = a.cols + off + lane;
    const float* __restrict__ v32 = a.v32 + off + lane;
    const double* __restrict__ v64 = a.v64 + off + lane;
    double acc = 0.0;
    for (int k0 = 0; k0 < W; k0 += kSU) {
      double vv[kSU];
      int cc[kSU];
#pragma unroll
      for (int j = 0; j < kSU; ++j) {
        const bool ok = k0 + j < W;
        const int o = (k0 + j) << 5;
        vv[j] = ok ? (F32 ? (double)s_ld_f32(v32 + o, pf) : s_ld_f64(v64 + o, pf)) : 0.0;
        cc[j] = ok ? s_ld_s32(c + o, pf) : 0;
      }
      double xx[kSU];
#pragma unroll
      for (int j = 0; j < kSU; ++j) {
        if (SPLIT) xx[j] = cc[j] < a.nsplit ? __ldg(a.x + cc[j]) : __ldg(a.x2 + (cc[j] - a.nsplit));
        else xx[j] = __ldg(a.x + cc[j]);
      }
#pragma unroll
      for (int j = 0; j < kSU; ++j) acc += vv[j] * xx[j];
    }
    const int row = (s << 5) + lane;
    if (row < a.n) {
      a.y[row] = acc;
      if (DOT) dacc += __ldg(a.x + row) * acc;
    }
  }
  if (DOT) {
    for (int o = 16; o > 0; o >>= 1) dacc += __shfl_xor_sync(0xffffffffu, dacc, o);
    if (lane == 0) red[warp] = dacc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double sum = 0.0;
      for (int k = 0; k < kSW; ++k) sum += red[k];
      a.part[blockIdx.x] = sum;
    }
  }
}

// ---- build -------------------------------------------------------------------------------
__global__ void k_sell_width(CsrView A, int nslices, long long* __restrict__ w32) {
  const int s = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (s > nslices) return;
  int len = 0;
  const int row = (s << 5) + lane;
  if (s < nslices && row < A.n) len = A.rowptr[row + 1] - A.rowptr[row];
  for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
  if (lane == 0) w32[s] = (long long)len << 5;   // entry nslices = 0: the scan's total lands there
}

template <bool F32>
__global__ void k_sell_fill(CsrView A, int nslices, const long long* __restrict__ sptr, int* __restrict__ cols,
                            float* __restrict__ v32, double* __restrict__ v64) {
  const int s = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (s >= nslices) return;
  const long long off = sptr[s];
  const int W = (int)((sptr[s + 1] - off) >> 5);
  const int row = (s << 5) + lane;
  int rs = 0, len = 0;
  if (row < A.n) { rs = A.rowptr[row]; len = A.rowptr[row + 1] - rs; }
  for (int k = 0; k < W; ++k) {
    const long long idx = off + ((long long)k << 5) + lane;
    const bool ok = k < len;
    cols[idx] = ok ? A.colidx[rs + k] : 0;      // padding: value 0 times x[0]
    const double v = ok ? A.vals[rs + k] : 0.0;
    if (F32) v32[idx] = (float)v; else v64[idx] = v;
  }
}

void sell_free(fs_sell* s) { delete s; }

void sell_build(const fs_csr& A, bool f32, fs_sell& out) {
  cudaStream_t st = stream();
  const int n = (int)A.n;
  const int nslices = div_up(n, 32);
  DBuf<long long> w32((size_t)nslices + 1);
  out.sptr.alloc((size_t)nslices + 1);
  const int g = (int)div_up(((int64_t)nslices + 1) * 32, 256);
  k_sell_width<<<g, 256, 0, st>>>(A.view(), nslices, w32.p);
  FS_LAUNCH_CHECK();
  size_t bytes = 0;
  FS_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, w32.p, out.sptr.p, nslices + 1, st));
  DBuf<char> tmp(bytes);
  FS_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, w32.p, out.sptr.p, nslices + 1, st));
  count_launch(2);
  long long total = 0;
  FS_CUDA(cudaMemcpyAsync(&total, out.sptr.p + nslices, sizeof(long long), cudaMemcpyDeviceToHost, st));
  FS_CUDA(cudaStreamSynchronize(st));
  out.n = n;
  out.nslices = nslices;
  out.padded = total;
  out.nnz = A.nnz;
  out.cols.alloc((size_t)total);
  if (f32) out.v32.alloc((size_t)total); else out.v64.alloc((size_t)total);
  if (f32) k_sell_fill<true><<<g, 256, 0, st>>>(A.view(), nslices, out.sptr.p, out.cols.p, out.v32.p, nullptr);
  else k_sell_fill<false><<<g, 256, 0, st>>>(A.view(), nslices, out.sptr.p, out.cols.p, nullptr, out.v64.p);
  FS_LAUNCH_CHECK();
  FS_CUDA(cudaStreamSynchronize(st));
}

template <bool SPLIT, bool DOT>
static void launch_sell(const SellArgs& args, int grid) {
  if (args.v32) k_spmv_sell<SPLIT, DOT, true><<<grid, kST, 0, stream()>>>(args);
  else k_spmv_sell<SPLIT, DOT, false><<<grid, kST, 0, stream()>>>(args);
}

// Returns the grid (= number of dot partials when asked for), 0 if S is empty.
int spmv_sell(const fs_sell& S, const double* x, double* y, const double* x2, int nsplit, double* dot_partials) {
  if (!S.nslices) return 0;
  const int grid = std::max(1, std::min(div_up(S.nslices, kSW), sm_count() * 6));
  SellArgs args{S.n, S.nslices, S.sptr.p, S.cols.p, S.v32.p, S.v64.p, x, x2, x2 ? nsplit : 0x7fffffff, y, dot_partials};
  if (x2) {
    if (dot_partials) launch_sell<true, true>(args, grid);
    else launch_sell<true, false>(args, grid);
  } else {
    if (dot_partials) launch_sell<false, true>(args, grid);
    else launch_sell<false, false>(args, grid);
  }
  FS_LAUNCH_CHECK();
  return grid;
}

}  // namespace fs
