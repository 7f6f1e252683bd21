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
// Split form for y = A [x; x2]: the entries with column < nsplit come first, padded to the slice's
// longest first part, then the rest with columns relative to nsplit -- the gather source is then
// uniform per step and there is no per-entry select.
// The CSR-stream kernels measured ~1.0-1.3 nonzeros / cycle / SM whatever the value width (L1TEX
// bound: scattered gathers + shared-memory staging), so fp32 values bought nothing there.
// y = A x or A [x; x2] (columns >= nsplit read x2), optional per-CTA partials of x.y.
#include <climits>
#include <cub/cub.cuh>
#include <cuda_fp16.h>

#include "dist.cuh"

namespace fs {

constexpr int kST = 256;        // threads per CTA
constexpr int kSW = kST / 32;   // warps (slices in flight) per CTA

struct SellArgs {
  int n, nslices;
  const long long* sptr;
  const int* wg;        // split form: width of the first part of every slice
  const int* cols;
  const float* v32;
  const double* v64;
  const double* x;
  const double* x2;
  double* y;
  double* part;
  int l2hint;
  const int* perm;      // SELL-C-sigma row of every slot (null: identity)
  const unsigned* pk;   // packed form (fp16 value << 16 | column offset), null: none
  const int2* cbase;    // per slice: base column of each part, .x == INT_MIN: slice kept in cols / v32
  double pk_inv;        // 1 / (power-of-two scale of the packed values)
  const float* xf;      // FMT 4: fp32 mirrors of x / x2 (the gather sources of packed slices)
  const float* x2f;
  float* yf;            // optional fp32 mirror of y (any format)
};

// streamed once per launch: read-only path, no L1 allocation (the L1 is for the gathered vectors), and an L2
// evict-first hint: one PCG iteration streams ~630 MB of matrix entries through a 126 MB L2 that should rather keep
// the ~100 MB of CG / multigrid vectors between the kernels that produce and consume them (FS_L2_HINT=0: no hint)
__device__ __forceinline__ uint64_t sell_policy(int hint) {
  uint64_t p;
  if (hint) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ float ld_stream(const float* a, uint64_t pol) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(a), "l"(pol));
  return v;
}
__device__ __forceinline__ double ld_stream(const double* a, uint64_t pol) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(a), "l"(pol));
  return v;
}
__device__ __forceinline__ int ld_stream(const int* a, uint64_t pol) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(pol));
  return v;
}
__device__ __forceinline__ unsigned ld_stream(const unsigned* a, uint64_t pol) {
  unsigned v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(pol));
  return v;
}

// One part of a slice: W entries per lane at stride 32, all gathered from x.  Unguarded batches of 8
// (loads first, then the gathers, then the sums), then a guarded tail.
// The fp32 (preconditioner) form uses fused multiply-adds; the fp64 form keeps multiply-then-add so
// that its sums have the same bits as the CSR kernels' (the library is built with -fmad=false).
template <class VT>
__device__ __forceinline__ double sell_mac(VT v, double xv, double acc) {
  if (sizeof(VT) == 4) return fma((double)v, xv, acc);
  return acc + (double)v * xv;
}

template <int U, class VT>
__device__ __forceinline__ double sell_batch(const VT* __restrict__ vp, const int* __restrict__ cp, const double* __restrict__ x,
                                             double acc, uint64_t pol) {
  VT vv[U];
  int cc[U];
#pragma unroll
  for (int j = 0; j < U; ++j) { vv[j] = ld_stream(vp + (j << 5), pol); cc[j] = ld_stream(cp + (j << 5), pol); }
  double xx[U];
#pragma unroll
  for (int j = 0; j < U; ++j) xx[j] = __ldg(x + cc[j]);
#pragma unroll
  for (int j = 0; j < U; ++j) acc = sell_mac<VT>(vv[j], xx[j], acc);
  return acc;
}

// guarded batch for the last rem < U entries
template <int U, class VT>
__device__ __forceinline__ double sell_tail(const VT* __restrict__ vp, const int* __restrict__ cp, int rem,
                                            const double* __restrict__ x, double acc, uint64_t pol) {
  VT vv[U];
  int cc[U];
#pragma unroll
  for (int j = 0; j < U; ++j) {
    const bool ok = j < rem;
    vv[j] = ok ? ld_stream(vp + (j << 5), pol) : VT(0);
    cc[j] = ok ? ld_stream(cp + (j << 5), pol) : 0;
  }
  double xx[U];
#pragma unroll
  for (int j = 0; j < U; ++j) xx[j] = __ldg(x + cc[j]);
#pragma unroll
  for (int j = 0; j < U; ++j) acc = sell_mac<VT>(vv[j], xx[j], acc);
  return acc;
}

template <class VT>
__device__ __forceinline__ double sell_part(const VT* __restrict__ vp, const int* __restrict__ cp, int W,
                                            const double* __restrict__ x, double acc, uint64_t pol) {
  int k = 0;
  for (; k + 8 <= W; k += 8, vp += 256, cp += 256) acc = sell_batch<8, VT>(vp, cp, x, acc, pol);
  if (sizeof(VT) == 8) {
    // fp64 (the CG's A*p, rows of ~7): the last 1..7 entries as ONE guarded batch -- a second dependent
    // load / gather round costs more than the guards
    if (k < W) acc = sell_tail<7, VT>(vp, cp, W - k, x, acc, pol);
  } else {
    // fp32 two-part operators: 4 unguarded + up to 3 guarded keeps the kernel inside 40 registers
    if (k + 4 <= W) { acc = sell_batch<4, VT>(vp, cp, x, acc, pol); k += 4; vp += 128; cp += 128; }
    if (k < W) acc = sell_tail<3, VT>(vp, cp, W - k, x, acc, pol);
  }
  return acc;
}

// finest up-sweep (split fp32 form with the fused dot, rows of ~7 + ~7): batches of 4 only, so that the
// kernel fits 32 registers and all 64 warps of an SM are resident (measured: 50.9 us against 53.8 us at 40
// registers / 48 warps; the deeper rows of the coarser levels prefer the 8-wide batches)
__device__ __forceinline__ double sell_part4(const float* __restrict__ vp, const int* __restrict__ cp, int W,
                                             const double* __restrict__ x, double acc, uint64_t pol) {
  int k = 0;
  for (; k + 4 <= W; k += 4, vp += 128, cp += 128) acc = sell_batch<4, float>(vp, cp, x, acc, pol);
  if (k < W) acc = sell_tail<3, float>(vp, cp, W - k, x, acc, pol);
  return acc;
}

// ---- packed entries: 4 bytes instead of 8 per nonzero ----------------------------------------------------
// The preconditioner's operators only have to be a fixed SPD map, so their values are stored as fp16 (scaled by a power of
// two per matrix: no iteration more than with fp32 values on the bench meshes, against +2 for bf16 --
// scripts/proto_operator_precision.py) and the column as a 16-bit offset from the slice's base column: the rows of a slice are
// neighbours in the mesh numbering, their columns sit in a narrow band.  One 32-bit stream load per entry instead of two,
// half the bytes; xb = x + base.
// gather x[base + 16-bit offset]: one LOP + one IMAD.WIDE.U32 + the load
__device__ __forceinline__ double pk_gather(const double* xb, unsigned w) {
  const double* p;
  asm("mad.wide.u32 %0, %1, 8, %2;" : "=l"(p) : "r"(w & 0xffffu), "l"(xb));
  return __ldg(p);
}
__device__ __forceinline__ double pk_mac(unsigned w, double xv, double acc) {
  return fma((double)__half2float(__ushort_as_half((unsigned short)(w >> 16))), xv, acc);
}

template <int U>
__device__ __forceinline__ double pk_batch(const unsigned* __restrict__ wp, const double* __restrict__ xb, double acc, uint64_t pol) {
  unsigned ww[U];
#pragma unroll
  for (int j = 0; j < U; ++j) ww[j] = ld_stream(wp + (j << 5), pol);
  double xx[U];
#pragma unroll
  for (int j = 0; j < U; ++j) xx[j] = pk_gather(xb, ww[j]);
#pragma unroll
  for (int j = 0; j < U; ++j) acc = pk_mac(ww[j], xx[j], acc);
  return acc;
}

template <int U>
__device__ __forceinline__ double pk_tail(const unsigned* __restrict__ wp, int rem, const double* __restrict__ xb, double acc,
                                          uint64_t pol) {
  unsigned ww[U];
#pragma unroll
  for (int j = 0; j < U; ++j) ww[j] = j < rem ? ld_stream(wp + (j << 5), pol) : 0u;   // value 0 times xb[0]
  double xx[U];
#pragma unroll
  for (int j = 0; j < U; ++j) xx[j] = pk_gather(xb, ww[j]);
#pragma unroll
  for (int j = 0; j < U; ++j) acc = pk_mac(ww[j], xx[j], acc);
  return acc;
}

template <int U>
__device__ __forceinline__ double pk_part(const unsigned* __restrict__ wp, int W, const double* __restrict__ xb, double acc,
                                          uint64_t pol) {
  int k = 0;
  for (; k + U <= W; k += U, wp += U * 32) acc = pk_batch<U>(wp, xb, acc, pol);
  if (U == 8) {
    if (k + 4 <= W) { acc = pk_batch<4>(wp, xb, acc, pol); k += 4; wp += 128; }
  }
  if (k < W) acc = pk_tail<3>(wp, W - k, xb, acc, pol);
  return acc;
}

// fp32 gathers (FMT 4).  The gathered vector costs two registers per entry in flight as fp64 and one as fp32, and half the
// L1 / L2 sectors: with an fp32 mirror of the input vectors (written by their producers) a batch of 8 entries fits the
// 32-register budget of 64 resident warps.  Products and the row sum are fp32 (13 terms, operators already rounded to
// fp16); the row result is widened and scaled in fp64.  The preconditioner stays a fixed map to 1e-7.
__device__ __forceinline__ float pk_gather32(const float* xb, unsigned w) {
  const float* p;
  asm("mad.wide.u32 %0, %1, 4, %2;" : "=l"(p) : "r"(w & 0xffffu), "l"(xb));
  return __ldg(p);
}

template <int U, bool GUARD>
__device__ __forceinline__ float pk32_batch(const unsigned* __restrict__ wp, int rem, const float* __restrict__ xb, float acc, uint64_t pol) {
  unsigned ww[U];
#pragma unroll
  for (int j = 0; j < U; ++j) ww[j] = (!GUARD || j < rem) ? ld_stream(wp + (j << 5), pol) : 0u;
  float xx[U];
#pragma unroll
  for (int j = 0; j < U; ++j) xx[j] = pk_gather32(xb, ww[j]);
#pragma unroll
  for (int j = 0; j < U; ++j) acc = fmaf(__half2float(__ushort_as_half((unsigned short)(ww[j] >> 16))), xx[j], acc);
  return acc;
}

__device__ __forceinline__ float pk32_part(const unsigned* __restrict__ wp, int W, const float* __restrict__ xb, float acc, uint64_t pol) {
  int k = 0;
  for (; k + 8 <= W; k += 8, wp += 256) acc = pk32_batch<8, false>(wp, 8, xb, acc, pol);
  if (k < W) acc = pk32_batch<8, true>(wp, W - k, xb, acc, pol);     // the last 1..7 entries as one guarded batch
  return acc;
}

// These kernels are latency bound, not bandwidth bound: a warp walks its slice as a chain of dependent memory rounds
// (slice header -> stream loads -> gathers -> next batch ...), and the time is (rounds per slice) / (resident warps) --
// measured: the same code at 48 instead of 64 warps per SM runs 1.4x longer.  Tried and dropped: all 16 packed words of a
// slice loaded at once with the next header prefetched (48 registers, 40 warps per SM): ptxas sinks the second group of
// loads behind the first group's gathers whatever the source order, 46.9 us against 43.0 us for the finest up-sweep
// (profiles/r02_ab_round2b.txt, FS_PK_MODE=3).  What helps is a smaller footprint per entry in flight: fp32 gathers.

// DIST (partitioned step, pstokes.cu): x / x2 are [own | halo] vectors whose halo entries are written by the
// neighbouring ranks; the kernel returns at once when the running solve has converged and otherwise
// waits for the halo flags of its input channels before the first gather (dist.cuh).
// Boundary first: the few slices that read halo entries or whose rows other ranks read ("boundary slices", listed in
// blist) are done at the START of the kernel -- wait for the input flags, compute, store the rows into the
// neighbours' halo slots, and the last boundary CTA releases the output channel's flags -- while all other CTAs are
// already streaming the interior.  The flags are thus on their way one whole kernel before the consumer needs them,
// and the consumer's own wait (same scheme) finds them set: the transfer latency hides behind the interior work.
// A producer for a replicated level (gather) sends every row; it releases its flags at the end.
struct DistSell {
  Comm c;
  HaloWait w;
  PushSpec ps;
  const unsigned* bmask = nullptr;
  const int* blist = nullptr;
  const int2* btab = nullptr;
  int n_blist = 0;
  int nb_cta = 0;
  int tag = 0;
};

// FMT: 0 fp64 values, 1 fp32 values, 2 packed entries in batches of 4 with fp64 gathers, 4 packed entries in batches of 8
// with fp32 gathers (both 32 registers, 64 warps per SM); slices that could not be packed take the fp32-value path
template <bool SPLIT, bool DOT, int FMT, bool DIST>
__global__ void __launch_bounds__(kST, ((SPLIT && FMT == 1 && DOT && !DIST) || FMT == 2 || FMT == 4) ? 8 : 6) k_spmv_sell(SellArgs a, DistSell d) {
  constexpr bool F32 = FMT >= 1;
  __shared__ double red[kSW];
  if (!DIST) pdl_launch();
  if (DIST) {
    if (d.c.done && *d.c.done) return;
    dist_trace(d.c, d.tag * 10 + 0);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t pol = sell_policy(a.l2hint);
  double dacc = 0.0;
  bool pushed = false;
  // the first nb CTAs do the boundary slices and nothing else, the others share the interior
  const bool two = DIST && d.n_blist > 0;
  const int nb = two ? d.nb_cta : 0;
  const bool bcta = two && (int)blockIdx.x < nb;
  if (bcta) {
    halo_wait(d.c, d.w);
    dist_trace(d.c, d.tag * 10 + 1);
  }
  const int count = bcta ? d.n_blist : a.nslices;
  const int first = (bcta ? blockIdx.x : blockIdx.x - nb) * kSW + warp;
  const int nwarps = (bcta ? nb : (int)gridDim.x - nb) * kSW;
  {
  // programmatic dependent launch: everything above ran while the previous kernel was finishing.  (Reading the first
  // header and prefetching its entry lines into L2 before the wait was measured: no gain, and the header kept live
  // across the loop cost registers -- spills at the 32 / 40 register budgets.)
  if (!DIST) pdl_wait();
  for (int si = first; si < count; si += nwarps) {
    int s = si;
    int2 dst = make_int2(-1, -1);
    if (two) {
      if (bcta) {
        s = __ldg(d.blist + si);
        if (d.btab) dst = __ldg(d.btab + ((size_t)si << 5) + lane);
      } else if ((__ldg(d.bmask + (si >> 5)) >> (si & 31)) & 1u) continue;
    }
    const long long off = __ldg(a.sptr + s);
    const int W = (int)((__ldg(a.sptr + s + 1) - off) >> 5);
    const int Wg = SPLIT ? __ldg(a.wg + s) : W;
    const int* __restrict__ cp = a.cols + off + lane;
    double acc = 0.0;
    int2 cb = make_int2(INT_MIN, 0);
    if (FMT >= 2) cb = __ldg(a.cbase + s);
    if (FMT >= 2 && cb.x != INT_MIN) {
      const unsigned* __restrict__ wp = a.pk + off + lane;
      const double* xb = a.x + cb.x;
      const double* xb2 = SPLIT ? a.x2 + cb.y : nullptr;
      if (FMT == 4) {
        float f = pk32_part(wp, Wg, a.xf + cb.x, 0.0f, pol);
        if (SPLIT) f = pk32_part(wp + ((long long)Wg << 5), W - Wg, a.x2f + cb.y, f, pol);
        acc = (double)f;
      } else {
        acc = pk_part<4>(wp, Wg, xb, acc, pol);
        if (SPLIT) acc = pk_part<4>(wp + ((long long)Wg << 5), W - Wg, xb2, acc, pol);
      }
      acc *= a.pk_inv;
    } else if (F32) {
      const float* __restrict__ vp = a.v32 + off + lane;
      if ((SPLIT && DOT) || FMT >= 2) {   // the finest up-sweep; unpackable slices of a packed matrix
        acc = sell_part4(vp, cp, Wg, a.x, acc, pol);
        if (SPLIT) acc = sell_part4(vp + ((long long)Wg << 5), cp + ((long long)Wg << 5), W - Wg, a.x2, acc, pol);
      } else {
        acc = sell_part<float>(vp, cp, Wg, a.x, acc, pol);
        if (SPLIT) acc = sell_part<float>(vp + ((long long)Wg << 5), cp + ((long long)Wg << 5), W - Wg, a.x2, acc, pol);
      }
    } else {
      const double* __restrict__ vp = a.v64 + off + lane;
      acc = sell_part<double>(vp, cp, Wg, a.x, acc, pol);
      if (SPLIT) acc = sell_part<double>(vp + ((long long)Wg << 5), cp + ((long long)Wg << 5), W - Wg, a.x2, acc, pol);
    }
    const int slot = (s << 5) + lane;
    if (slot < a.n) {
      const int row = a.perm ? __ldg(a.perm + slot) : slot;
      a.y[row] = acc;
      if (a.yf) a.yf[row] = (float)acc;
      if (DOT) dacc += __ldg(a.x + row) * acc;
      if (DIST && d.ps.enabled) {
        if (d.ps.gather) pushed |= push_row(d.ps, row, acc);
        else if (bcta) {
          if (dst.y == -2) pushed |= push_row(d.ps, row, acc);
          else {
            if (dst.x >= 0) dist_st_sys_f64(reinterpret_cast<double*>(d.ps.peer_base[dst.x >> 28] + d.ps.vec_off) + (dst.x & 0x0fffffff), acc);
            if (dst.y >= 0) dist_st_sys_f64(reinterpret_cast<double*>(d.ps.peer_base[dst.y >> 28] + d.ps.vec_off) + (dst.y & 0x0fffffff), acc);
          }
        }
      }
    }
  }
  if (bcta && d.ps.enabled && !d.ps.gather)
    push_finish(d.c, d.ps, pushed, dist_seq(d.c), nb, d.tag * 10 + 4);      // the send rows all sit in boundary slices
  if (bcta) dist_trace(d.c, d.tag * 10 + 2);
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
  if (DIST && d.ps.enabled && d.ps.gather) push_finish(d.c, d.ps, pushed, dist_seq(d.c), (int)gridDim.x);
  if (DIST) { dist_trace(d.c, d.tag * 10 + 3); dist_trace_last(d.c, d.tag * 10 + 6); }
}

// ---- bulk-async (TMA) staged variant -----------------------------------------------------------------
// The register-staged kernel above issues, per batch, a round of stream loads (values + columns), waits, issues the
// gathers, waits again: two dependent memory round trips per batch, and the registers that hold a batch in flight cap
// the occupancy.  Here every warp owns a ring of kBS stages in shared memory; lane 0 keeps kBS chunks of the warp's
// slice stream (8 entries per lane: 1 KB of columns + 1 or 2 KB of values, contiguous in the SELL layout) in flight with
// cp.async.bulk (1-D TMA, L2 evict-first hint) completing on one mbarrier per stage, and the lanes only read the landed
// chunk from shared memory, gather and accumulate -- one round trip on the critical path, no staging registers.
constexpr int kBK = 8;          // entries per lane and chunk

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}

template <bool SPLIT, bool DOT, bool F32, int kBS, int MINB>
__global__ void __launch_bounds__(kST, MINB) k_spmv_sell_bulk(SellArgs a) {
  using VT = typename std::conditional<F32, float, double>::type;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ double red[kSW];
  __shared__ __align__(8) unsigned long long bars[kSW * kBS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = gridDim.x * kSW;
  constexpr int kColB = kBK * 32 * 4, kValB = kBK * 32 * (int)sizeof(VT), kStageB = kColB + kValB;
  unsigned char* ring = smem_raw + (size_t)warp * kBS * kStageB;
  const unsigned ring_s = smem_u32(ring);
  const unsigned bar0 = smem_u32(bars + warp * kBS);
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  if (lane == 0) {
    for (int k = 0; k < kBS; ++k) mbar_init(bar0 + 8 * k, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
  // producer cursor (kept identical in all lanes; lane 0 issues)
  int p_slice = blockIdx.x * kSW + warp, p_k = 0, p_W = 0, p_n = 0;
  long long p_off = 0;
  auto issue = [&]() -> bool {
    while (p_slice < a.nslices) {
      if (p_k == 0) {
        p_off = __ldg(a.sptr + p_slice);
        p_W = (int)((__ldg(a.sptr + p_slice + 1) - p_off) >> 5);
      }
      if (p_k < p_W) break;
      p_slice += nwarps;       // empty slice: nothing to stream
      p_k = 0;
    }
    if (p_slice >= a.nslices) return false;
    const int cnt = min(kBK, p_W - p_k);
    const int st = p_n % kBS;
    if (lane == 0) {
      const unsigned bar = bar0 + 8 * st, dst = ring_s + st * kStageB;
      const long long e0 = p_off + ((long long)p_k << 5);
      mbar_expect_tx(bar, (unsigned)(cnt * 32 * (4 + (int)sizeof(VT))));
      bulk_g2s(dst, a.cols + e0, (unsigned)(cnt * 128), bar, pol);
      bulk_g2s(dst + kColB, (F32 ? (const void*)(a.v32 + e0) : (const void*)(a.v64 + e0)), (unsigned)(cnt * 32 * (int)sizeof(VT)), bar, pol);
    }
    p_k += cnt;
    if (p_k >= p_W) { p_slice += nwarps; p_k = 0; }
    ++p_n;
    return true;
  };
  for (int k = 0; k < kBS; ++k) issue();
  double dacc = 0.0;
  int c_n = 0;
  for (int s = blockIdx.x * kSW + warp; s < a.nslices; s += nwarps) {
    const long long off = __ldg(a.sptr + s);
    const int W = (int)((__ldg(a.sptr + s + 1) - off) >> 5);
    const int Wg = SPLIT ? __ldg(a.wg + s) : W;
    double acc = 0.0;
    for (int k0 = 0; k0 < W; k0 += kBK) {
      const int st = c_n % kBS;
      mbar_wait(bar0 + 8 * st, (unsigned)((c_n / kBS) & 1));
      const int* sc = reinterpret_cast<const int*>(ring + st * kStageB) + lane;
      const VT* sv = reinterpret_cast<const VT*>(ring + st * kStageB + kColB) + lane;
      const int cnt = min(kBK, W - k0);
      double xx[kBK];
      if (cnt == kBK) {
#pragma unroll
        for (int j = 0; j < kBK; ++j) xx[j] = __ldg(((!SPLIT || k0 + j < Wg) ? a.x : a.x2) + sc[j << 5]);
#pragma unroll
        for (int j = 0; j < kBK; ++j) acc = sell_mac<VT>(sv[j << 5], xx[j], acc);     // values straight from shared memory
      } else {
#pragma unroll
        for (int j = 0; j < kBK; ++j) xx[j] = __ldg(((!SPLIT || k0 + j < Wg) ? a.x : a.x2) + (j < cnt ? sc[j << 5] : 0));
#pragma unroll
        for (int j = 0; j < kBK; ++j) acc = sell_mac<VT>(j < cnt ? sv[j << 5] : VT(0), xx[j], acc);
      }
      __syncwarp();                                                     // every lane is done with the stage
      if (lane == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue();                                                          // refill it
      ++c_n;
    }
    const int slot = (s << 5) + lane;
    if (slot < a.n) {
      const int row = a.perm ? __ldg(a.perm + slot) : slot;
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

// two interleaved right-hand sides (x, y are (n,2) row-major), fp64 values: the viscous 2-RHS CG
template <bool DOT, bool DIST>
__global__ void __launch_bounds__(kST, 4) k_spmv_sell2(SellArgs a, const int* __restrict__ done, DistSell d) {
  __shared__ double red[2 * kSW];
  if (done && *done) return;
  if (DIST) halo_wait(d.c, d.w);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = gridDim.x * kSW;
  const uint64_t pol = sell_policy(a.l2hint);
  const double2* __restrict__ x2v = reinterpret_cast<const double2*>(a.x);
  double d0 = 0.0, d1 = 0.0;
  for (int s = blockIdx.x * kSW + warp; s < a.nslices; s += nwarps) {
    const long long off = __ldg(a.sptr + s);
    const int W = (int)((__ldg(a.sptr + s + 1) - off) >> 5);
    const int* __restrict__ cp = a.cols + off + lane;
    const double* __restrict__ vp = a.v64 + off + lane;
    double a0 = 0.0, a1 = 0.0;
    for (int k0 = 0; k0 < W; k0 += 4, vp += 128, cp += 128) {
      double vv[4];
      int cc[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool ok = k0 + j < W;
        vv[j] = ok ? ld_stream(vp + (j << 5), pol) : 0.0;
        cc[j] = ok ? ld_stream(cp + (j << 5), pol) : 0;
      }
      double2 xx[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) xx[j] = __ldg(x2v + cc[j]);
#pragma unroll
      for (int j = 0; j < 4; ++j) { a0 += vv[j] * xx[j].x; a1 += vv[j] * xx[j].y; }
    }
    const int row = (s << 5) + lane;
    if (row < a.n) {
      reinterpret_cast<double2*>(a.y)[row] = make_double2(a0, a1);
      if (DOT) { const double2 xr = __ldg(x2v + row); d0 += xr.x * a0; d1 += xr.y * a1; }
    }
  }
  if (DOT) {
    for (int o = 16; o > 0; o >>= 1) { d0 += __shfl_xor_sync(0xffffffffu, d0, o); d1 += __shfl_xor_sync(0xffffffffu, d1, o); }
    if (lane == 0) { red[2 * warp] = d0; red[2 * warp + 1] = d1; }
    __syncthreads();
    if (threadIdx.x < 2) {
      double sum = 0.0;
      for (int k = 0; k < kSW; ++k) sum += red[2 * k + threadIdx.x];
      a.part[2 * blockIdx.x + threadIdx.x] = sum;
    }
  }
}

// ---- build -------------------------------------------------------------------------------
// number of leading entries of a (column-sorted) row with column < nsplit
__device__ __forceinline__ int row_split(const CsrView& A, int rs, int len, int nsplit) {
  int g = 0;
  while (g < len && A.colidx[rs + g] < nsplit) ++g;
  return g;
}

// per slice: 32 * (max first-part length + max second-part length); nsplit = INT_MAX: one part
__global__ void k_sell_width(CsrView A, int nslices, int nsplit, const int* __restrict__ perm, long long* __restrict__ w32,
                             int* __restrict__ wg) {
  const int s = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (s > nslices) return;
  int lg = 0, lp = 0;
  const int slot = (s << 5) + lane;
  const int row = (s < nslices && slot < A.n) ? (perm ? perm[slot] : slot) : A.n;
  if (s < nslices && row < A.n) {
    const int rs = A.rowptr[row], len = A.rowptr[row + 1] - rs;
    lg = row_split(A, rs, len, nsplit);
    lp = len - lg;
  }
  for (int o = 16; o > 0; o >>= 1) {
    lg = max(lg, __shfl_xor_sync(0xffffffffu, lg, o));
    lp = max(lp, __shfl_xor_sync(0xffffffffu, lp, o));
  }
  if (lane == 0) {
    w32[s] = (long long)(lg + lp) << 5;   // entry nslices = 0: the scan's total lands there
    if (wg && s < nslices) wg[s] = lg;
  }
}

template <bool F32>
__global__ void k_sell_fill(CsrView A, int nslices, int nsplit, const int* __restrict__ perm, const long long* __restrict__ sptr,
                            const int* __restrict__ wg, int* __restrict__ cols, float* __restrict__ v32, double* __restrict__ v64) {
  const int s = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (s >= nslices) return;
  const long long off = sptr[s];
  const int W = (int)((sptr[s + 1] - off) >> 5);
  const int Wg = wg ? wg[s] : W;
  const int slot = (s << 5) + lane;
  const int row = slot < A.n ? (perm ? perm[slot] : slot) : A.n;
  int rs = 0, len = 0, lg = 0;
  if (row < A.n) { rs = A.rowptr[row]; len = A.rowptr[row + 1] - rs; lg = row_split(A, rs, len, nsplit); }
  for (int k = 0; k < W; ++k) {
    const long long idx = off + ((long long)k << 5) + lane;
    // source entry: first part k < lg, second part (k - Wg) < len - lg; padding: value 0 times x[0]
    int src = -1;
    if (k < Wg) { if (k < lg) src = rs + k; }
    else if (k - Wg < len - lg) src = rs + lg + (k - Wg);
    const int col = src >= 0 ? A.colidx[src] : 0;
    cols[idx] = (src >= 0 && k >= Wg) ? col - nsplit : col;
    const double v = src >= 0 ? A.vals[src] : 0.0;
    if (F32) v32[idx] = (float)v; else v64[idx] = v;
  }
}

// ---- packed form ---------------------------------------------------------------------------------------
__global__ void k_maxabs_bits(const double* __restrict__ v, int64_t n, unsigned long long* __restrict__ out) {
  unsigned long long m = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double a = fabs(v[i]);
    if (a == a) m = max(m, (unsigned long long)__double_as_longlong(a));   // non-negative doubles order like their bits
  }
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

double csr_maxabs(const fs_csr& A) {
  if (!A.nnz || !A.vals.p) return 0.0;
  DBuf<unsigned long long> m(1);
  m.zero();
  k_maxabs_bits<<<(int)std::min<int64_t>(div_up(A.nnz, 256), 2048), 256, 0, stream()>>>(A.vals.p, A.nnz, m.p);
  FS_LAUNCH_CHECK();
  const unsigned long long bits = m.to_host()[0];
  double r;
  std::memcpy(&r, &bits, sizeof r);
  return r;
}

static int sell_pack_mode() {  // read at every set-up (not cached): tests switch it between hierarchies
  const char* e = std::getenv("FS_SELL_PACK");
  return e ? std::atoi(e) : 1;
}
// FS_SELL_PACK=0: fp32 value + 32-bit column per entry; 1 (default): packed; 2: the packed form's rounding with every
// slice left in the 32-bit encoding (test: both encodings must give the same bits)
bool sell_pack_enabled() { return sell_pack_mode() != 0; }

// One warp per slice.  Rounds the slice's fp32 values to fp16 at the matrix scale (v32 keeps the rounded value, so the
// 32-bit encoding of an unpackable slice gives the same products), finds the column range of each part over the real
// entries and, if both ranges fit 16 bits, writes the packed words.  Padding (value 0) becomes word 0: 0 * x[base].
__global__ void k_sell_pack(int nslices, const long long* __restrict__ sptr, const int* __restrict__ wg, const int* __restrict__ cols,
                            float* __restrict__ v32, float scale, float inv_scale, unsigned* __restrict__ pk, int2* __restrict__ cbase,
                            unsigned long long* __restrict__ n_unpacked, int force_unpacked) {
  const int s = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (s >= nslices) return;
  const long long off = sptr[s];
  const int W = (int)((sptr[s + 1] - off) >> 5);
  const int Wg = wg ? wg[s] : W;
  int lo1 = INT_MAX, hi1 = INT_MIN, lo2 = INT_MAX, hi2 = INT_MIN;
  for (int k = 0; k < W; ++k) {
    const long long idx = off + ((long long)k << 5) + lane;
    const float h = __half2float(__float2half_rn(v32[idx] * scale));
    v32[idx] = h * inv_scale;
    if (h != 0.0f) {
      const int c = cols[idx];
      if (k < Wg) { lo1 = min(lo1, c); hi1 = max(hi1, c); }
      else { lo2 = min(lo2, c); hi2 = max(hi2, c); }
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo1 = min(lo1, __shfl_xor_sync(0xffffffffu, lo1, o)); hi1 = max(hi1, __shfl_xor_sync(0xffffffffu, hi1, o));
    lo2 = min(lo2, __shfl_xor_sync(0xffffffffu, lo2, o)); hi2 = max(hi2, __shfl_xor_sync(0xffffffffu, hi2, o));
  }
  if (hi1 < lo1) lo1 = hi1 = 0;     // a part without real entries
  if (hi2 < lo2) lo2 = hi2 = 0;
  const bool ok = !force_unpacked && (long long)hi1 - lo1 <= 65535 && (long long)hi2 - lo2 <= 65535;
  if (lane == 0) {
    cbase[s] = ok ? make_int2(lo1, lo2) : make_int2(INT_MIN, 0);
    if (!ok) atomicAdd(n_unpacked, 1ull);
  }
  for (int k = 0; k < W; ++k) {
    const long long idx = off + ((long long)k << 5) + lane;
    unsigned w = 0;
    if (ok) {
      const __half h = __float2half_rn(v32[idx] * scale);     // exact: v32 is already an fp16 value / scale
      const unsigned hb = __half_as_ushort(h);
      if (hb & 0x7fffu) w = (hb << 16) | (unsigned)(cols[idx] - (k < Wg ? lo1 : lo2));
    }
    pk[idx] = w;
  }
}

void sell_free(fs_sell* s) { delete s; }

// nsplit >= 0: two-part slices (columns < nsplit first, padded per slice; the second part's columns
// are stored relative to nsplit), for y = A [x; x2] without a per-entry select.
__global__ void k_sigma_keys(const int* __restrict__ rowptr, int n, int sigma, unsigned long long* __restrict__ keys, int* __restrict__ rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned len = (unsigned)(rowptr[i + 1] - rowptr[i]);
  keys[i] = ((unsigned long long)(unsigned)(i / sigma) << 32) | (0xffffffffu - len);   // window, then longest first (stable: row order)
  rows[i] = i;
}

void sell_build(const fs_csr& A, bool f32, fs_sell& out, int nsplit, int sigma, double pk_maxabs) {
  cudaStream_t st = stream();
  const int n = (int)A.n;
  const int nslices = div_up(n, 32);
  const bool split = nsplit >= 0;
  const int* perm = nullptr;
  if (sigma > 32 && n > sigma) {
    DBuf<unsigned long long> keys(n), keys2(n);
    DBuf<int> rows(n);
    out.perm.alloc(n);
    k_sigma_keys<<<div_up(n, 256), 256, 0, st>>>(A.rowptr, n, sigma, keys.p, rows.p);
    FS_LAUNCH_CHECK();
    size_t bytes = 0;
    FS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys.p, keys2.p, rows.p, out.perm.p, n, 0, 64, st));
    DBuf<char> tmp(bytes);
    FS_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, keys.p, keys2.p, rows.p, out.perm.p, n, 0, 64, st));
    count_launch(2);
    FS_CUDA(cudaStreamSynchronize(st));
    perm = out.perm.p;
  }
  DBuf<long long> w32((size_t)nslices + 1);
  out.sptr.alloc((size_t)nslices + 1);
  if (split) out.wg.alloc((size_t)nslices);
  const int g = (int)div_up(((int64_t)nslices + 1) * 32, 256);
  k_sell_width<<<g, 256, 0, st>>>(A.view(), nslices, split ? nsplit : 0x7fffffff, perm, w32.p, split ? out.wg.p : nullptr);
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
  out.nsplit = nsplit;
  out.cols.alloc((size_t)total);
  if (f32) out.v32.alloc((size_t)total); else out.v64.alloc((size_t)total);
  const int ns = split ? nsplit : 0x7fffffff;
  if (f32) k_sell_fill<true><<<g, 256, 0, st>>>(A.view(), nslices, ns, perm, out.sptr.p, out.wg.p, out.cols.p, out.v32.p, nullptr);
  else k_sell_fill<false><<<g, 256, 0, st>>>(A.view(), nslices, ns, perm, out.sptr.p, out.wg.p, out.cols.p, nullptr, out.v64.p);
  FS_LAUNCH_CHECK();
  FS_CUDA(cudaStreamSynchronize(st));
  if (f32 && sell_pack_enabled() && total > 0) {
    const double m = pk_maxabs > 0.0 ? pk_maxabs : csr_maxabs(A);
    if (m > 0.0 && m < 1e30 && m > 1e-30) {
      // largest value lands in [2^14, 2^15): inside fp16's range, 24 binades above its smallest subnormal
      const double scale = std::ldexp(1.0, 14 - std::ilogb(m));
      out.pk.alloc((size_t)total);
      out.cbase.alloc((size_t)nslices);
      DBuf<unsigned long long> cnt(1);
      cnt.zero();
      k_sell_pack<<<(int)div_up((int64_t)nslices * 32, 256), 256, 0, st>>>(nslices, out.sptr.p, split ? out.wg.p : nullptr, out.cols.p,
                                                                           out.v32.p, (float)scale, (float)(1.0 / scale), out.pk.p,
                                                                           out.cbase.p, cnt.p, sell_pack_mode() == 2 ? 1 : 0);
      FS_LAUNCH_CHECK();
      out.pk_inv = 1.0 / scale;
      out.pk_unpacked = (long long)cnt.to_host()[0];
    }
  }
}

static int l2_hint() {
  static const int v = [] { const char* e = std::getenv("FS_L2_HINT"); return e ? std::atoi(e) : 1; }();
  return v;
}

static int pk_mode() {   // FS_PK_MODE: kernel for packed matrices (2 / 4, see k_spmv_sell); 4 needs the fp32 mirrors
  static const int v = [] { const char* e = std::getenv("FS_PK_MODE"); return e ? std::atoi(e) : 4; }();
  return v;
}

template <bool SPLIT, bool DOT>
static void launch_sell(const SellArgs& args, int grid, const DistSell* d) {
  static const DistSell none{};
  if (d) {
    if (args.pk) k_spmv_sell<SPLIT, DOT, 2, true><<<grid, kST, 0, stream()>>>(args, *d);
    else if (args.v32) k_spmv_sell<SPLIT, DOT, 1, true><<<grid, kST, 0, stream()>>>(args, *d);
    else k_spmv_sell<SPLIT, DOT, 0, true><<<grid, kST, 0, stream()>>>(args, *d);
  } else {
    if (args.pk && args.xf) launch_pdl(k_spmv_sell<SPLIT, DOT, 4, false>, grid, kST, 0, args, none);
    else if (args.pk) launch_pdl(k_spmv_sell<SPLIT, DOT, 2, false>, grid, kST, 0, args, none);
    else if (args.v32) launch_pdl(k_spmv_sell<SPLIT, DOT, 1, false>, grid, kST, 0, args, none);
    else launch_pdl(k_spmv_sell<SPLIT, DOT, 0, false>, grid, kST, 0, args, none);
  }
}

// y = S x, or S [x; x2] for a matrix built in split form.  Returns the grid (= number of dot
// partials when asked for), 0 if S is empty.
// FS_SELL_TMA=1: the bulk-async staged kernel for the single-GPU SpMVs
template <bool SPLIT, bool DOT, bool F32, int kBS, int MINB>
static int launch_bulk_cfg(const SellArgs& args) {
  constexpr int vb = F32 ? 4 : 8;
  const size_t smem = (size_t)kSW * kBS * (kBK * 32 * (4 + vb));
  static int per_sm = -1;
  if (per_sm < 0) {
    FS_CUDA(cudaFuncSetAttribute(k_spmv_sell_bulk<SPLIT, DOT, F32, kBS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_spmv_sell_bulk<SPLIT, DOT, F32, kBS, MINB>, kST, smem));
    per_sm = std::max(per_sm, 1);
  }
  const int grid = std::max(1, std::min(div_up(args.nslices, kSW), sm_count() * per_sm));
  k_spmv_sell_bulk<SPLIT, DOT, F32, kBS, MINB><<<grid, kST, smem, stream()>>>(args);
  FS_LAUNCH_CHECK();
  return grid;
}

static int sell_tma_mode();
template <bool SPLIT, bool DOT, bool F32>
static int launch_bulk(const SellArgs& args) {
  switch (sell_tma_mode()) {
    case 2: return launch_bulk_cfg<SPLIT, DOT, F32, 2, (F32 ? 6 : 4)>(args);      // 2 stages, many warps
    case 3: return launch_bulk_cfg<SPLIT, DOT, F32, 3, (F32 ? 4 : 3)>(args);
    default: return launch_bulk_cfg<SPLIT, DOT, F32, 4, (F32 ? 3 : 2)>(args);     // 4 stages, 24 / 16 warps per SM
  }
}

static int sell_tma_mode() {
  static const int v = [] { const char* e = std::getenv("FS_SELL_TMA"); return e ? std::atoi(e) : 0; }();
  return v;
}

static int spmv_sell_impl(const fs_sell& S, const double* x, double* y, const double* x2, double* dot_partials,
                          DistSell* d, const SellF32* f32v = nullptr) {
  if (!S.nslices) return 0;
  FS_REQUIRE((S.nsplit >= 0) == (x2 != nullptr), "spmv_sell: split form and second vector must come together");
  if (!d && sell_tma_mode() && S.nslices >= 4096) {
    SellArgs args{S.n, S.nslices, S.sptr.p, S.wg.p, S.cols.p, S.v32.p, S.v64.p, x, x2, y, dot_partials, 1, S.perm.p, nullptr, nullptr, 1.0,
                  nullptr, nullptr, nullptr};
    const bool f32 = S.v32.p != nullptr;
    if (x2) {
      if (dot_partials) return f32 ? launch_bulk<true, true, true>(args) : launch_bulk<true, true, false>(args);
      return f32 ? launch_bulk<true, false, true>(args) : launch_bulk<true, false, false>(args);
    }
    if (dot_partials) return f32 ? launch_bulk<false, true, true>(args) : launch_bulk<false, true, false>(args);
    return f32 ? launch_bulk<false, false, true>(args) : launch_bulk<false, false, false>(args);
  }
  // fp32 gathers: every gather source needs its mirror (and the kernel is only built for the single-GPU path)
  const bool g32 = S.pk.p && !d && pk_mode() == 4 && f32v && f32v->xf && (!x2 || f32v->x2f);
  const int per_sm = ((x2 && S.v32.p && dot_partials && !d) || S.pk.p) ? 8 : 6;   // 8: kernels that fit 32 registers
  int grid = std::max(1, std::min(div_up(S.nslices, kSW), sm_count() * per_sm));
  SellArgs args{S.n, S.nslices, S.sptr.p, S.wg.p, S.cols.p, S.v32.p, S.v64.p, x, x2, y, dot_partials, l2_hint(), S.perm.p,
                S.pk.p, S.cbase.p, S.pk_inv, g32 ? f32v->xf : nullptr, g32 ? f32v->x2f : nullptr, f32v ? f32v->yf : nullptr};
  FS_REQUIRE(!(d && S.perm.p), "SELL-C-sigma matrices are not used by the partitioned kernels");
  int grid_add = 0;
  if (d && (d->w.nch || (d->ps.enabled && !d->ps.gather)) && S.n_blist > 0) {
    d->bmask = S.bmask.p; d->blist = S.blist.p; d->n_blist = S.n_blist;
    d->btab = (d->ps.enabled && !d->ps.gather) ? S.btab.p : nullptr;
    // dedicated boundary CTAs, scheduled first.  Their warps take about half as many slices as an interior warp, so the
    // flags leave in the first half of the kernel (the consumer is a later kernel) without idling a third of the SMs
    const int per_warp = std::max(1, S.nslices / (2 * grid * kSW));
    d->nb_cta = std::max(1, std::min(div_up(S.n_blist, kSW * per_warp), grid / 2));
    grid_add = d->nb_cta;
  }
  if (d && d->ps.enabled && !d->ps.gather && S.n_blist == 0) d->ps.enabled = 0;   // nothing to send from these rows
  if (grid_add) grid = std::max(grid, grid_add + 1);      // at least one interior CTA; the total stays one resident wave
  if (x2) {
    if (dot_partials) launch_sell<true, true>(args, grid, d);
    else launch_sell<true, false>(args, grid, d);
  } else {
    if (dot_partials) launch_sell<false, true>(args, grid, d);
    else launch_sell<false, false>(args, grid, d);
  }
  FS_LAUNCH_CHECK();
  return grid;
}

int spmv_sell(const fs_sell& S, const double* x, double* y, const double* x2, double* dot_partials, const SellF32* f32v) {
  return spmv_sell_impl(S, x, y, x2, dot_partials, nullptr, f32v);
}

int spmv_sell_dist(const fs_sell& S, const double* x, double* y, const double* x2, double* dot_partials, const Comm& c,
                   const HaloWait& w, const PushSpec* push, int trace_tag) {
  DistSell d;
  d.c = c;
  d.w = w;
  d.tag = trace_tag;
  if (push) d.ps = *push;
  return spmv_sell_impl(S, x, y, x2, dot_partials, &d);
}

// ---- boundary slices ---------------------------------------------------------------------------
__global__ void k_mark_boundary(CsrView A, int n_own_a, int nsplit, int n_own_b, const unsigned* __restrict__ out_smask,
                                unsigned* __restrict__ bmask) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= A.n) return;
  bool halo = out_smask && ((out_smask[row >> 5] >> (row & 31)) & 1u);    // the row itself is read by another rank
  for (int k = A.rowptr[row]; k < A.rowptr[row + 1]; ++k) {
    const int c = A.colidx[k];
    if (c < nsplit) halo |= (c >= n_own_a);
    else halo |= (n_own_b >= 0 && c - nsplit >= n_own_b);
  }
  if (halo) atomicOr(bmask + (row >> 10), 1u << ((row >> 5) & 31));
}

void sell_mark_boundary(fs_sell& S, const fs_csr& loc, int n_own_a, int nsplit, int n_own_b, const Space* out) {
  const int nsl = S.nslices;
  if (!nsl) return;
  const int nw = div_up(nsl, 32);
  const bool sends = out && !out->gather && out->n_send > 0;
  S.bmask.alloc(nw);
  S.bmask.zero();
  k_mark_boundary<<<div_up(loc.n, 256), 256, 0, stream()>>>(loc.view(), n_own_a, nsplit, n_own_b, sends ? out->smask.p : nullptr, S.bmask.p);
  FS_LAUNCH_CHECK();
  std::vector<unsigned> h = S.bmask.to_host();
  std::vector<int> list;
  for (int s = 0; s < nsl; ++s) if ((h[s >> 5] >> (s & 31)) & 1u) list.push_back(s);
  S.n_blist = (int)list.size();
  if (S.n_blist) { S.blist.alloc(list.size()); S.blist.upload(list.data(), list.size()); }
  if (sends && S.n_blist) {
    std::vector<int2> tab((size_t)S.n_blist * 32, make_int2(-1, -1));
    for (int j = 0; j < S.n_blist; ++j)
      for (int l = 0; l < 32; ++l) {
        const int row = list[j] * 32 + l;
        auto it = std::lower_bound(out->h_urow.begin(), out->h_urow.end(), row);
        if (it == out->h_urow.end() || *it != row) continue;
        const size_t u = it - out->h_urow.begin();
        const int k0 = out->h_uptr[u], k1 = out->h_uptr[u + 1];
        int2& e = tab[(size_t)j * 32 + l];
        if (k1 - k0 > 2 || out->h_udst[k0] >= (1 << 28)) { e.y = -2; continue; }
        e.x = (out->h_upeer[k0] << 28) | out->h_udst[k0];
        if (k1 - k0 == 2) e.y = (out->h_upeer[k0 + 1] << 28) | out->h_udst[k0 + 1];
      }
    S.btab.alloc(tab.size());
    S.btab.upload(tab.data(), tab.size());
  }
  FS_CUDA(cudaStreamSynchronize(stream()));
}


int spmv_sell_grid(const fs_sell& S) { return std::max(1, std::min(div_up(S.nslices, kSW), sm_count() * 4)); }

// (n,2) interleaved x, y; fp64 one-part matrix; optional partials of x.y per column (2 per CTA) and early-exit flag
void spmv_sell2(const fs_sell& S, const double* x, double* y, double* dot_partials, const int* done) {
  FS_REQUIRE(S.nslices && S.v64.p && S.nsplit < 0, "spmv_sell2: needs a one-part fp64 SELL matrix");
  const int grid = spmv_sell_grid(S);
  SellArgs args{S.n, S.nslices, S.sptr.p, nullptr, S.cols.p, nullptr, S.v64.p, x, nullptr, y, dot_partials, l2_hint(), nullptr, nullptr, nullptr, 1.0, nullptr, nullptr, nullptr};
  static const DistSell none{};
  if (dot_partials) k_spmv_sell2<true, false><<<grid, kST, 0, stream()>>>(args, done, none);
  else k_spmv_sell2<false, false><<<grid, kST, 0, stream()>>>(args, done, none);
  FS_LAUNCH_CHECK();
}

void spmv_sell2_dist(const fs_sell& S, const double* x, double* y, double* dot_partials, const int* done, const Comm& c,
                     const HaloWait& w) {
  FS_REQUIRE(S.nslices && S.v64.p && S.nsplit < 0, "spmv_sell2: needs a one-part fp64 SELL matrix");
  const int grid = spmv_sell_grid(S);
  SellArgs args{S.n, S.nslices, S.sptr.p, nullptr, S.cols.p, nullptr, S.v64.p, x, nullptr, y, dot_partials, l2_hint(), nullptr, nullptr, nullptr, 1.0, nullptr, nullptr, nullptr};
  DistSell d;
  d.c = c;
  d.w = w;
  if (dot_partials) k_spmv_sell2<true, true><<<grid, kST, 0, stream()>>>(args, done, d);
  else k_spmv_sell2<false, true><<<grid, kST, 0, stream()>>>(args, done, d);
  FS_LAUNCH_CHECK();
}

}  // namespace fs
