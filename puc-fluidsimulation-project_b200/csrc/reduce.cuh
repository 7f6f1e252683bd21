// reduce.cuh -- block / grid reduction helpers shared by the Krylov kernels (solver.cu, pstokes.cu).
#pragma once
#include "internal.cuh"

namespace fs {

constexpr int kBlock = 256;
constexpr int kMaxBlocks = 1024;   // partial arrays are sized for this

template <int R>
__device__ __forceinline__ void load_vec(const double* __restrict__ p, int64_t i, double (&o)[R]) {
  if (R == 1) o[0] = p[i];
  else { double2 t = reinterpret_cast<const double2*>(p)[i]; o[0] = t.x; o[R - 1] = t.y; }
}
template <int R>
__device__ __forceinline__ void load_vec_ldg(const double* __restrict__ p, int64_t i, double (&o)[R]) {
  if (R == 1) o[0] = __ldg(p + i);
  else { double2 t = __ldg(reinterpret_cast<const double2*>(p) + i); o[0] = t.x; o[R - 1] = t.y; }
}
template <int R>
__device__ __forceinline__ void store_vec(double* __restrict__ p, int64_t i, const double (&o)[R]) {
  if (R == 1) p[i] = o[0];
  else reinterpret_cast<double2*>(p)[i] = make_double2(o[0], o[R - 1]);
}

// block-wide deterministic sum of K values per thread -> out[0..K) valid in thread 0
template <int K>
__device__ __forceinline__ void block_reduce(double (&v)[K], double* smem /* K*32 */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < K; ++k) smem[k * 32 + warp] = v[k];
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double t = (lane < nw) ? smem[k * 32 + lane] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      v[k] = t;
    }
  }
  __syncthreads();
}

// every block re-reduces the partial array (nblk x K) in the same fixed order (thread-strided
// sums with the loads issued together, warp xor trees, then the warps in ascending order);
// result broadcast to all threads through shared memory.
template <int K>
__device__ __forceinline__ void reduce_partials(const double* __restrict__ part, int nblk, double (&out)[K],
                                                double* smem /* K */) {
  __shared__ double scr[32 * K];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5, nw = blockDim.x >> 5;
  double acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = 0.0;
  for (int b0 = t; b0 < nblk; b0 += 4 * blockDim.x) {
    double v[4][K];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int b = b0 + j * blockDim.x;
#pragma unroll
      for (int k = 0; k < K; ++k) v[j][k] = (b < nblk) ? part[(size_t)b * K + k] : 0.0;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < K; ++k) acc[k] += v[j][k];
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    if (lane == 0) scr[w * K + k] = acc[k];
  }
  __syncthreads();
  if (t < K) {
    double s = 0.0;
    for (int q = 0; q < nw; ++q) s += scr[q * K + t];
    smem[t] = s;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) out[k] = smem[k];
  __syncthreads();
}

// The convergence flag as ONE value per CTA: thread 0 reads it, everybody branches on the shared
// copy.  A per-thread read could split a CTA around the __syncthreads of the reductions below when
// another CTA of the same launch sets the flag meanwhile.
__device__ __forceinline__ bool block_done(const int* flags) {
  __shared__ int s_done;
  if (threadIdx.x == 0) s_done = *reinterpret_cast<const volatile int*>(flags);
  __syncthreads();
  return s_done != 0;
}

}  // namespace fs
