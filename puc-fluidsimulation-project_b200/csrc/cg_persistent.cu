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
  int* flags;           // [0]=converged [1]=iterations [2]=multi-GPU wait error (0 ok) [4],[5]=its committed copies
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

// ---- multi-GPU (one rank per GPU, peer memory over NVLink) ---------------------------
// Every rank owns a contiguous row block; p is stored as [own rows | halo rows] inside an
// IPC-shared "mailbox" allocation so that neighbours write their boundary values of p
// straight into this rank's halo slots (peer stores) from inside pass C.  The two dot
// product reductions per iteration are an all-to-all of per-rank partial sums through
// the same mailboxes, summed in rank order (deterministic).  Flags are monotonically
// increasing epochs written with release.sys / polled with acquire.sys.
constexpr int kMaxRanks = 8;
struct DistArgs {
  int rank, world, n_nbr;
  int nbr[kMaxRanks];                              // ranks this rank exchanges halos with
  double* red_local;                               // [2][world][4]
  unsigned long long* redflag_local;               // [2][world]
  unsigned long long* haloflag_local;              // [world]
  double* red_peer[kMaxRanks];
  unsigned long long* redflag_peer[kMaxRanks];
  unsigned long long* haloflag_peer[kMaxRanks];
  double* p_peer[kMaxRanks];                       // peers' p = [own | halo]
  const int* send_row;                             // own row to send
  const int* send_peer;                            // destination rank
  const int* send_dst;                             // slot in the destination's p
  const int* send_begin;                           // [grid+1] CTA b pushes entries [send_begin[b], send_begin[b+1])
  const int* needs_halo;                           // [grid] 1 if CTA b's rows reference a halo column
  unsigned long long red_epoch0, halo_epoch0;      // epochs already consumed by earlier solves
};

__device__ __forceinline__ void st_release_sys(unsigned long long* a, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* a) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(a) : "memory");
  return v;
}
__device__ __forceinline__ void st_sys_f64(double* a, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(a), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_sys_f64(const double* a) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(a) : "memory");
  return v;
}
// bounded spin: a lost peer must end in an error code, not in a hung GPU.  After a
// time-out (or once any thread of this rank has timed out) every wait falls through, the
// kernel runs to its end with garbage and the host reports FS_ERR_INTERNAL.
// (FS_DIST_TIMEOUT_MS, default 20 s, measured with %globaltimer: a slow start of a peer does not trip it)
__device__ unsigned long long g_wait_timeout_ns = 20000000000ull;
__device__ __forceinline__ void wait_flag(const unsigned long long* a, unsigned long long want, int* err, int code) {
  unsigned spins = 0;
  unsigned long long t0 = 0;
  while (ld_acquire_sys(a) < want) {
    if ((++spins & 255u) == 0) {
      if (*(volatile int*)err != 0) return;
      if (t0 == 0) t0 = gtime();
      else if (gtime() - t0 > g_wait_timeout_ns) { atomicCAS(err, 0, code); return; }
    }
  }
}

// all ranks: v[k] (already summed over the local CTAs, identical in every CTA) -> sum over ranks.
// Low-latency protocol: every double travels as two 8-byte words {32 data bits | 32-bit epoch};
// an 8-byte store is atomic, so data and "ready" flag arrive together and no fence or separate
// flag write is needed.  CTA 0 posts this rank's sums into every rank's mailbox (peer stores),
// every CTA polls its own mailbox and adds the world's contributions in rank order.
__device__ __forceinline__ void st_sys_u64(unsigned long long* a, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(a), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long* a) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(a) : "memory");
  return v;
}

template <int K>
__device__ __forceinline__ void rank_allreduce(const DistArgs& d, double (&v)[K], unsigned long long epoch, double* sm, int* err) {
  static_assert(K <= 2, "mailbox slot holds two doubles");
  const int t = threadIdx.x;
  const int par = (int)(epoch & 1ull);
  const unsigned long long e32 = (epoch & 0xffffffffull) << 32;
  if (blockIdx.x == 0 && t < d.world) {
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(d.red_peer[t]) + ((size_t)par * d.world + d.rank) * 4;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const unsigned long long bits = (unsigned long long)__double_as_longlong(v[k]);
      st_sys_u64(dst + 2 * k, (bits & 0xffffffffull) | e32);
      st_sys_u64(dst + 2 * k + 1, (bits >> 32) | e32);
    }
  }
  __shared__ double got[kMaxRanks][2];
  if (t < d.world) {
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(d.red_local) + ((size_t)par * d.world + t) * 4;
    unsigned long long w[2 * K];
    unsigned spins = 0;
    unsigned long long t0 = 0;
    bool ok = false;
    while (!ok) {
      ok = true;
#pragma unroll
      for (int j = 0; j < 2 * K; ++j) { w[j] = ld_sys_u64(src + j); ok = ok && ((w[j] & 0xffffffff00000000ull) == e32); }
      if (!ok && (++spins & 255u) == 0) {
        if (*(volatile int*)err != 0) break;
        if (t0 == 0) t0 = gtime();
        else if (gtime() - t0 > g_wait_timeout_ns) { atomicCAS(err, 0, 0x200 | t); break; }
      }
    }
#pragma unroll
    for (int k = 0; k < K; ++k)
      got[t][k] = __longlong_as_double((long long)((w[2 * k] & 0xffffffffull) | (w[2 * k + 1] << 32)));
  }
  __syncthreads();
  if (t == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double s = 0.0;
      for (int q = 0; q < d.world; ++q) s += got[q][k];
      sm[k] = s;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = sm[k];
  __syncthreads();
}

// CTA b pushes the boundary values of p it owns into the neighbours' halo slots
__device__ __forceinline__ void push_halo(const DistArgs& d, const double* p_own) {
  const int s0 = d.send_begin[blockIdx.x], s1 = d.send_begin[blockIdx.x + 1];
  if (s0 == s1) return;
  __syncthreads();
  for (int k = s0 + threadIdx.x; k < s1; k += blockDim.x)
    st_sys_f64(d.p_peer[d.send_peer[k]] + d.send_dst[k], __ldcg(p_own + d.send_row[k]));
  __threadfence_system();
}

template <bool DIST>
__global__ void __launch_bounds__(kPT, 2) k_cg_persistent(CgPersistArgs a, DistArgs d) {
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
  unsigned long long red_epoch = d.red_epoch0, halo_epoch = d.halo_epoch0;
  if (DIST) {
    // p = z0 was written by the init kernels: publish its boundary values
    push_halo(d, a.p);
    grid.sync();
    ++halo_epoch;
    if (b == 0 && t < d.n_nbr) { __threadfence_system(); st_release_sys(d.haloflag_peer[d.nbr[t]] + d.rank, halo_epoch); }
  }

  while (it < a.maxit) {
    if (DIST) {
      if (d.needs_halo[b]) {      // only the CTAs whose rows reference halo columns have to wait
        if (t < d.n_nbr) wait_flag(d.haloflag_local + d.nbr[t], halo_epoch, a.flags + 2, 0x100 | d.nbr[t]);
        __syncthreads();
      }
    }
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
    // A time-out must end the solve in ALL CTAs at the same barrier (a CTA that left the loop alone would leave the others
    // hanging in the next grid.sync): CTA 0 copies the live error word into a per-iteration-parity slot BEFORE the barrier,
    // every CTA reads that slot AFTER it -- one value for the whole grid, and a slow reader is never overtaken (two slots).
    if (DIST && b == 0 && t == 0) __stcg(a.flags + 4 + (it & 1), *(volatile int*)(a.flags + 2));
    grid.sync();
    if (DIST && __ldcg(a.flags + 4 + (it & 1)) != 0) break;
    if (timer) { unsigned long long t1 = gtime(); nsA += t1 - t0; t0 = t1; }

    // ------------------------------------------------------------------ pass B
    double pAp[1];
    sum_partials<1>(a.partA, nb, pAp, sm);
    if (DIST) rank_allreduce<1>(d, pAp, ++red_epoch, sm, a.flags + 2);
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
    if (DIST) rank_allreduce<2>(d, sB, ++red_epoch, sm, a.flags + 2);
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
    if (DIST) push_halo(d, a.p);
    grid.sync();
    if (DIST) {
      ++halo_epoch;
      if (b == 0 && t < d.n_nbr) { __threadfence_system(); st_release_sys(d.haloflag_peer[d.nbr[t]] + d.rank, halo_epoch); }
    }
    if (timer) { unsigned long long t1 = gtime(); nsC += t1 - t0; t0 = t1; }
  }
  if (timer) {
    a.scal[0] = rz_old;
    a.scal[3] = rr;
    a.scal[8] = (double)nsA;
    a.scal[9] = (double)nsB;
    a.scal[10] = (double)nsC;
    a.scal[11] = (double)it;
    a.scal[12] = (double)red_epoch;
    a.scal[13] = (double)halo_epoch;
    a.flags[0] = converged ? 1 : 0;
    a.flags[1] = it;
  }
}

static int g_persist_state = -1;   // -1 unknown, 0 unsupported, 1 ready

static size_t persist_smem(const CsrView& A) {
  const size_t wcap = ((size_t)A.wtile_nnz_max + 31) / 32 * 32;
  return (size_t)kPW * wcap * sizeof(double);
}

// Returns false when this matrix cannot use the persistent kernel (a 32-row mini-tile too
// large for shared memory, cooperative launch unsupported); the caller then runs the
// 3-kernels-per-iteration path.
bool cg_persistent_supported(const CsrView& A, size_t* smem_out) {
  if (A.wtile_nnz_max <= 0) return false;
  const size_t smem = persist_smem(A);
  if (smem > 100 * 1024) return false;
  if (g_persist_state < 0) {
    int dev = 0, coop = 0;
    FS_CUDA(cudaGetDevice(&dev));
    FS_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    if (!coop) { g_persist_state = 0; return false; }
    FS_CUDA(cudaFuncSetAttribute(k_cg_persistent<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    FS_CUDA(cudaFuncSetAttribute(k_cg_persistent<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    g_persist_state = 1;
  }
  if (g_persist_state == 0) return false;
  if (smem_out) *smem_out = smem;
  return true;
}

int cg_persistent_grid(const CsrView& A) {
  size_t smem = 0;
  FS_REQUIRE(cg_persistent_supported(A, &smem), "persistent CG not supported for this matrix");
  int per_sm = 0;
  FS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cg_persistent<true>, kPT, smem));
  FS_REQUIRE(per_sm >= 1, "persistent CG kernel does not fit on an SM");
  per_sm = std::min(per_sm, 2);
  return std::max(1, std::min(sm_count() * per_sm, div_up(A.n, kPT)));
}

// first row of CTA b's block for a grid of nb CTAs (same arithmetic as the kernel)
int cg_persistent_block_row0(const CsrView& A, int b, int nb) {
  const int ntiles = div_up(A.n, kPT);
  return std::min<long long>(A.n, (((long long)ntiles * b) / nb) * kPT);
}

// Launch the persistent solve.  scal/flags as documented in CgPersistArgs.
void cg_persistent_launch(const CsrView& A, double* x, double* r, double* p, double* Ap, const double* dinv,
                          double* partA, double* partB, double* scal, int* flags, int maxit, double tol2,
                          const DistArgs* dist) {
  const int grid = cg_persistent_grid(A);
  const size_t smem = persist_smem(A);
  const int wcap = (A.wtile_nnz_max + 31) / 32 * 32;
  CgPersistArgs args{A, x, r, p, Ap, dinv, partA, partB, scal, flags, maxit, tol2, div_up(A.n, kPT), wcap};
  DistArgs d{};
  if (dist) d = *dist;
  void* kargs[] = {&args, &d};
  void* fn = dist ? (void*)k_cg_persistent<true> : (void*)k_cg_persistent<false>;
  FS_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kPT), kargs, smem, stream()));
  count_launch();
}

void cg_persistent_launch(const CsrView& A, double* x, double* r, double* p, double* Ap, const double* dinv,
                          double* partA, double* partB, double* scal, int* flags, int maxit, double tol2) {
  cg_persistent_launch(A, x, r, p, Ap, dinv, partA, partB, scal, flags, maxit, tol2, nullptr);
}

// needs[b] = 1 iff some nonzero of CTA b's row block has a halo column (>= n_own)
__global__ void k_needs_halo(const int* __restrict__ rowptr, const int* __restrict__ colidx, const int* __restrict__ row0,
                             int n_own, int* __restrict__ needs) {
  const int b = blockIdx.x;
  int any = 0;
  for (int k = rowptr[row0[b]] + threadIdx.x; k < rowptr[row0[b + 1]]; k += blockDim.x) any |= (colidx[k] >= n_own);
  any = __syncthreads_or(any);
  if (threadIdx.x == 0) needs[b] = any;
}

// ---- init of a partitioned solve with x0 = 0: r = b, p[own] = Dinv b, local (b.b, r.z) ----
__global__ void __launch_bounds__(256)
k_dist_init(int n, const double* __restrict__ b, const double* __restrict__ dinv, double* __restrict__ x,
            double* __restrict__ r, double* __restrict__ p, double* __restrict__ part) {
  __shared__ double red[2][8];
  double s0 = 0.0, s1 = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double bv = b[i], z = (dinv ? dinv[i] : 1.0) * bv;
    x[i] = 0.0; r[i] = bv; p[i] = z;
    s0 += bv * bv; s1 += bv * z;
  }
  for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s0; red[1][threadIdx.x >> 5] = s1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a0 = 0.0, a1 = 0.0;
    for (int w = 0; w < 8; ++w) { a0 += red[0][w]; a1 += red[1][w]; }
    part[2 * blockIdx.x] = a0; part[2 * blockIdx.x + 1] = a1;
  }
}

}  // namespace fs

// ======================================================================================
// fs_dist: the row-block partitioned pressure CG
// ======================================================================================
struct fs_dist {
  int rank = 0, world = 1;
  int64_t n_own = 0, n_halo = 0;
  fs_csr mat;                      // local rows, columns in [0, n_own + n_halo)
  fs::DBuf<double> x, r, Ap, b;    // n_own
  // IPC-shared allocation.  The control words sit at FIXED offsets (the same on every rank,
  // whatever its row count) so that a rank can address its peers' flags; p follows them.
  fs::DBuf<char> mailbox;          // [red 1 KB | redflag 1 KB | haloflag 2 KB | p (n_own+n_halo)]
  static constexpr size_t off_red = 0, off_redflag = 1024, off_haloflag = 2048, off_p = 4096;
  size_t mailbox_bytes = 0;
  void* peer_base[fs::kMaxRanks] = {nullptr};
  bool peer_opened[fs::kMaxRanks] = {false};
  fs::DBuf<int> send_row, send_peer, send_dst, send_begin, needs_halo;
  std::vector<int> nbr;
  unsigned long long red_epoch = 0, halo_epoch = 0;
  bool connected = false;
  double* p() { return reinterpret_cast<double*>(mailbox.p + off_p); }
  ~fs_dist() {
    for (int q = 0; q < fs::kMaxRanks; ++q)
      if (peer_opened[q]) cudaIpcCloseMemHandle(peer_base[q]);
  }
};

using namespace fs;

extern "C" {

int fs_dist_create(int rank, int world, int64_t n_own, int64_t n_halo, int64_t nnz, const int32_t* rowptr,
                   const int32_t* colidx, const double* vals, fs_dist** out) {
  FS_API_BEGIN
  FS_REQUIRE(out, "out is NULL");
  *out = nullptr;
  FS_REQUIRE(world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world, "bad rank/world (max 8 ranks)");
  FS_REQUIRE(n_own > 0 && n_halo >= 0 && rowptr && colidx && vals, "bad arguments");
  std::unique_ptr<fs_dist> d(new fs_dist());
  d->rank = rank; d->world = world; d->n_own = n_own; d->n_halo = n_halo;
  d->mat.n = n_own; d->mat.nnz = nnz;
  d->mat.rowptr_own.alloc(n_own + 1); d->mat.rowptr_own.upload(rowptr, n_own + 1);
  d->mat.colidx_own.alloc(nnz); d->mat.colidx_own.upload(colidx, nnz);
  d->mat.vals.alloc(nnz); d->mat.vals.upload(vals, nnz);
  d->mat.rowptr = d->mat.rowptr_own.p; d->mat.colidx = d->mat.colidx_own.p;
  d->x.alloc(n_own); d->r.alloc(n_own); d->Ap.alloc(n_own); d->b.alloc(n_own);
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  static_assert(2 * kMaxRanks * 4 * sizeof(double) <= 1024 && 2 * kMaxRanks * 8 <= 1024, "control block layout");
  d->mailbox_bytes = fs_dist::off_p + up((size_t)(n_own + n_halo) * sizeof(double));
  d->mailbox.alloc(d->mailbox_bytes);
  d->mailbox.zero();
  fs::sync();
  *out = d.release();
  FS_API_END
}

int fs_dist_destroy(fs_dist* d) {
  FS_API_BEGIN
  if (d) { cudaStreamSynchronize(stream()); delete d; }
  FS_API_END
}

int fs_dist_ipc_handle(fs_dist* d, void* handle64) {
  FS_API_BEGIN
  FS_REQUIRE(d && handle64, "NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  FS_CUDA(cudaIpcGetMemHandle(&h, d->mailbox.p));
  std::memcpy(handle64, &h, 64);
  FS_API_END
}

int fs_dist_connect(fs_dist* d, const void* all_handles, const int32_t* send_row, const int32_t* send_peer,
                    const int32_t* send_dst, int64_t n_send) {
  FS_API_BEGIN
  FS_REQUIRE(d && all_handles, "NULL argument");
  for (int q = 0; q < d->world; ++q) {
    if (q == d->rank) { d->peer_base[q] = d->mailbox.p; continue; }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, (const char*)all_handles + 64 * q, 64);
    FS_CUDA(cudaIpcOpenMemHandle(&d->peer_base[q], h, cudaIpcMemLazyEnablePeerAccess));
    d->peer_opened[q] = true;
  }
  // send list must be sorted by own row so that every CTA pushes a contiguous range
  std::vector<int> rows(send_row, send_row + n_send), peers(send_peer, send_peer + n_send), dsts(send_dst, send_dst + n_send);
  for (int64_t k = 1; k < n_send; ++k) FS_REQUIRE(rows[k - 1] <= rows[k], "send list must be sorted by row");
  std::vector<char> isn(d->world, 0);
  for (int64_t k = 0; k < n_send; ++k) {
    FS_REQUIRE(peers[k] >= 0 && peers[k] < d->world && peers[k] != d->rank, "bad destination rank");
    FS_REQUIRE(rows[k] >= 0 && rows[k] < d->n_own, "send row out of range");
    isn[peers[k]] = 1;
  }
  d->nbr.clear();
  for (int q = 0; q < d->world; ++q) if (isn[q]) d->nbr.push_back(q);
  ensure_tiles(&d->mat);
  const CsrView A = d->mat.view();
  const int grid = cg_persistent_grid(A);
  std::vector<int> begin(grid + 1, 0);
  for (int b = 0; b <= grid; ++b) {
    const int row0 = (b == grid) ? (int)d->n_own : cg_persistent_block_row0(A, b, grid);
    begin[b] = (int)(std::lower_bound(rows.begin(), rows.end(), row0) - rows.begin());
  }
  d->send_row.alloc(std::max<int64_t>(n_send, 1)); d->send_peer.alloc(std::max<int64_t>(n_send, 1));
  d->send_dst.alloc(std::max<int64_t>(n_send, 1)); d->send_begin.alloc(grid + 1);
  if (n_send) { d->send_row.upload(rows.data(), n_send); d->send_peer.upload(peers.data(), n_send); d->send_dst.upload(dsts.data(), n_send); }
  d->send_begin.upload(begin.data(), grid + 1);
  {
    std::vector<int> row0(grid + 1);
    for (int b = 0; b <= grid; ++b) row0[b] = (b == grid) ? (int)d->n_own : cg_persistent_block_row0(A, b, grid);
    DBuf<int> drow0(grid + 1);
    drow0.upload(row0.data(), grid + 1);
    d->needs_halo.alloc(grid);
    k_needs_halo<<<grid, 256, 0, stream()>>>(d->mat.rowptr, d->mat.colidx, drow0.p, (int)d->n_own, d->needs_halo.p);
    FS_LAUNCH_CHECK();
    fs::sync();
  }
  fs::sync();
  d->connected = true;
  FS_API_END
}

// local part of the init (x0 = 0): returns this rank's (b.b, b.Dinv b); the caller sums them over ranks
int fs_dist_cg_begin(fs_dist* d, const double* b_own, int precond, double* local_sums2) {
  FS_API_BEGIN
  FS_REQUIRE(d && b_own && local_sums2 && d->connected, "bad arguments / not connected");
  const double* dinv = nullptr;
  if (precond == FS_PRECOND_JACOBI) { jacobi_prepare(&d->mat); dinv = d->mat.dinv.p; }
  In<double> ib(b_own, d->n_own);
  const int g = std::max(1, std::min(div_up(d->n_own, 256), 512));
  DBuf<double> part(2 * g);
  k_dist_init<<<g, 256, 0, stream()>>>((int)d->n_own, ib.d, dinv, d->x.p, d->r.p, d->p(), part.p);
  FS_LAUNCH_CHECK();
  std::vector<double> h = part.to_host();
  local_sums2[0] = local_sums2[1] = 0.0;
  for (int k = 0; k < g; ++k) { local_sums2[0] += h[2 * k]; local_sums2[1] += h[2 * k + 1]; }
  FS_API_END
}

int fs_dist_cg_run(fs_dist* d, double bb_global, double rz_global, double* x_own, double rtol, int maxit, int precond,
                   int* iters, double* relres, double* ns_pass3) {
  FS_API_BEGIN
  FS_REQUIRE(d && x_own && d->connected, "bad arguments / not connected");
  const double* dinv = (precond == FS_PRECOND_JACOBI) ? d->mat.dinv.p : nullptr;
  ensure_tiles(&d->mat);
  const CsrView A = d->mat.view();
  if (!d->mat.partials.n) d->mat.partials.alloc(4096 * 4);
  if (!d->mat.scal.n) d->mat.scal.alloc(64);
  double* scal = d->mat.scal.p;
  int* flags = reinterpret_cast<int*>(scal + 32);
  double hs[16] = {0};
  hs[0] = rz_global; hs[2] = bb_global; hs[3] = bb_global;
  FS_CUDA(cudaMemcpyAsync(scal, hs, sizeof(hs), cudaMemcpyHostToDevice, stream()));
  FS_CUDA(cudaMemsetAsync(flags, 0, 8 * sizeof(int), stream()));   // [4], [5]: committed copies of the error word (uniform exit)
  static bool timeout_set = false;
  if (!timeout_set) {
    if (const char* e = std::getenv("FS_DIST_TIMEOUT_MS")) {
      const unsigned long long ns = (unsigned long long)std::atof(e) * 1000000ull;
      FS_CUDA(cudaMemcpyToSymbol(g_wait_timeout_ns, &ns, sizeof(ns)));
    }
    timeout_set = true;
  }
  Out<double> ox(x_own, d->n_own);
  int it = 0;
  double rr = bb_global;
  if (bb_global > 0.0) {
    DistArgs da{};
    da.rank = d->rank; da.world = d->world; da.n_nbr = (int)d->nbr.size();
    for (size_t k = 0; k < d->nbr.size(); ++k) da.nbr[k] = d->nbr[k];
    char* mb = d->mailbox.p;
    da.red_local = reinterpret_cast<double*>(mb + d->off_red);
    da.redflag_local = reinterpret_cast<unsigned long long*>(mb + d->off_redflag);
    da.haloflag_local = reinterpret_cast<unsigned long long*>(mb + d->off_haloflag);
    for (int q = 0; q < d->world; ++q) {
      char* pb = (char*)d->peer_base[q];
      da.red_peer[q] = reinterpret_cast<double*>(pb + d->off_red);
      da.redflag_peer[q] = reinterpret_cast<unsigned long long*>(pb + d->off_redflag);
      da.haloflag_peer[q] = reinterpret_cast<unsigned long long*>(pb + d->off_haloflag);
      da.p_peer[q] = reinterpret_cast<double*>(pb + fs_dist::off_p);
    }
    da.send_row = d->send_row.p; da.send_peer = d->send_peer.p; da.send_dst = d->send_dst.p; da.send_begin = d->send_begin.p; da.needs_halo = d->needs_halo.p;
    da.red_epoch0 = d->red_epoch; da.halo_epoch0 = d->halo_epoch;
    cg_persistent_launch(A, d->x.p, d->r.p, d->p(), d->Ap.p, dinv, d->mat.partials.p, d->mat.partials.p + 4096,
                         scal, flags, maxit, rtol * rtol, &da);
    int hf[4];
    FS_CUDA(cudaMemcpyAsync(hs, scal, sizeof(hs), cudaMemcpyDeviceToHost, stream()));
    FS_CUDA(cudaMemcpyAsync(hf, flags, sizeof(hf), cudaMemcpyDeviceToHost, stream()));
    fs::sync();
    if (hf[2] != 0) {
      char msg[160];
      std::snprintf(msg, sizeof(msg), "partitioned CG: rank %d timed out waiting for the %s flag of rank %d "
                    "(peer kernel not running or peer memory not reachable)", d->rank,
                    (hf[2] & 0x100) ? "halo" : "reduction", hf[2] & 0xff);
      throw Error(FS_ERR_INTERNAL, msg);
    }
    d->red_epoch = (unsigned long long)hs[12];
    d->halo_epoch = (unsigned long long)hs[13];
    it = hf[0] ? hf[1] : -hf[1] - 1;
    rr = hs[3];
    if (ns_pass3) { ns_pass3[0] = hs[8]; ns_pass3[1] = hs[9]; ns_pass3[2] = hs[10]; }
  } else {
    d->x.zero();
  }
  FS_CUDA(cudaMemcpyAsync(ox.d, d->x.p, d->n_own * sizeof(double), cudaMemcpyDeviceToDevice, stream()));
  ox.commit();
  fs::sync();
  if (iters) *iters = it >= 0 ? it : -it - 1;
  if (relres) *relres = bb_global > 0.0 ? std::sqrt(rr / bb_global) : 0.0;
  if (it < 0) throw Error(FS_ERR_NOCONV, "fs_dist_cg_run: no convergence within maxit");
  FS_API_END
}

}  // extern "C"
