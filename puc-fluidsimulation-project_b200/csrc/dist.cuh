// dist.cuh -- device-side communication primitives of the partitioned Stokes step (pstokes.cu):
// one rank per GPU, every rank's "arena" (control words + the vectors that have halo entries) is one
// CUDA-IPC shared allocation, and all data-path communication happens INSIDE the compute kernels or
// in single-CTA push kernels over NVLink peer memory -- no host or NCCL call inside a step.
//
//   sequence number   One device-side counter per rank (`seq`), bumped by every kernel that performs a
//                     cross-rank reduction.  All ranks execute the same kernel sequence and the reductions
//                     are global synchronisation points, so the counters advance in lock step.
//   halo exchange     A vector with halo entries lives as [own | halo] inside the arena.  Its producer is
//                     followed by k_halo_push: boundary values are stored straight into the neighbours'
//                     halo slots (st.relaxed.sys), then the channel's flag in each neighbour is released
//                     with the current sequence number.  Consumers spin on their own flags
//                     (ld.acquire.sys >= seq) before the first halo read.  Between two pushes of the same
//                     channel there is always a reduction, which the consumer joins only after it has
//                     finished reading: no double buffering is needed.
//   all-reduce        Every CTA owns the local sum (re-reduced partials); CTA 0 posts it into every rank's
//                     mailbox as 8-byte words {32 data bits | 32-bit epoch} (atomic: data and "ready" arrive
//                     together, no fence, no separate flag); every CTA polls its own mailbox and adds the
//                     contributions in rank order => deterministic and bit-identical on all ranks, which
//                     keeps the ranks' convergence decisions in lock step.  Two mailbox slots by epoch parity.
//   time-outs         Every spin is bounded by %globaltimer; on expiry an error word is set, all later
//                     waits fall through, kernels run to their end and the host reports FS_ERR_INTERNAL.
#pragma once
#include "internal.cuh"

namespace fs {

constexpr int kMaxRanks = 8;
constexpr int kMaxChannels = 48;
constexpr int kRedWords = 16;          // mailbox slot: 8 doubles as 16 tagged words

// control block at the start of every arena (same layout on all ranks)
constexpr size_t kCtlRedOff = 0;                                             // [2][kMaxRanks][kRedWords] u64
constexpr size_t kCtlFlagOff = 2 * kMaxRanks * kRedWords * 8;                // [kMaxChannels][kMaxRanks] u64
constexpr size_t kCtlBytes = 8192;
static_assert(kCtlFlagOff + (size_t)kMaxChannels * kMaxRanks * 8 <= kCtlBytes, "control block layout");

struct Comm {
  int rank = 0, world = 1;
  unsigned long long* seq = nullptr;        // local: reduction sequence number
  int* err = nullptr;                       // local: first time-out (0 = none)
  unsigned* done_ctr = nullptr;             // local: last-block arrival counters (one per reducing kernel family)
  const int* done = nullptr;                // local: convergence flag of the running solve (kernels return when set)
  unsigned long long* red_local = nullptr;
  unsigned long long* flag_local = nullptr;
  unsigned long long* red_peer[kMaxRanks] = {nullptr};
  unsigned long long* flag_peer[kMaxRanks] = {nullptr};
  unsigned long long timeout_ns = 20000000000ull;
  // optional event trace (FS_DIST_TRACE=1): CTA 0 / thread 0 of the partitioned kernels appends {tag, %globaltimer}
  unsigned long long* trace = nullptr;
  unsigned* trace_n = nullptr;
  unsigned trace_cap = 0;
};

// which channels a consumer kernel has to see before its first halo read
struct HaloWait {
  int nch = 0;
  int ch[2] = {0, 0};
  int n_from[2] = {0, 0};
  signed char from[2][kMaxRanks] = {{0}};
};

struct PushArgs {
  const double* src = nullptr;      // own values, `stride` doubles per row
  int stride = 1;
  int channel = 0;
  int n_send = 0;
  const int* send_row = nullptr;    // own row
  const int* send_peer = nullptr;   // destination rank
  const int* send_dst = nullptr;    // slot (row index) in the destination's vector
  size_t vec_off = 0;               // byte offset of the vector inside every arena
  char* peer_base[kMaxRanks] = {nullptr};
  int n_to = 0;
  signed char to[kMaxRanks] = {0};
  // contiguous form (gathered levels): rows [row0, row0 + n_rows) go to the same rows of every peer in `to`
  int gather = 0;
  int row0 = 0, n_rows = 0;
};

// Fused push: the producer kernel of a halo'd vector stores its boundary rows into the neighbours' halo slots
// itself (no separate push kernel, no launch gap) and the CTA that finishes last releases the channel's flags.
// The send rows of a space are described per 32-row slice: smask[s] has a bit for every row of the slice that
// is read by another rank, sbase[s] indexes the slice's first such row in uptr, and uptr / upeer / udst list the
// (rank, slot) destinations of every send row.
struct PushSpec {
  int enabled = 0;
  int gather = 0;                   // replicated space: row r goes to row row0 + r of every rank in `to`
  int row0 = 0;
  int channel = 0;
  const unsigned* smask = nullptr;
  const int* slist = nullptr;       // the slices with a non-zero mask
  int n_slist = 0;
  const int* sbase = nullptr;
  const int* uptr = nullptr;
  const int* upeer = nullptr;
  const int* udst = nullptr;
  size_t vec_off = 0;
  char* peer_base[kMaxRanks] = {nullptr};
  int n_to = 0;
  signed char to[kMaxRanks] = {0};
};

// ---- host side ------------------------------------------------------------------------------------
// An index space (mesh nodes, or the rows of one AMG level) cut into contiguous blocks, one per rank.
// Every rank knows the halo lists of ALL ranks (they are derived from replicated global matrices), so
// send lists and arena layouts need no exchange: only the 64-byte IPC handles travel between processes.
struct Space {
  std::vector<int64_t> split;               // world+1 block boundaries (global ids)
  int64_t own_lo = 0, n_own = 0, n_halo = 0;
  int64_t cap = 0;                          // max over ranks of n_own + n_halo: slot count of every vector
  std::vector<std::vector<int>> halo_all;   // per rank: sorted global ids outside its block that it reads
  DBuf<int> halo_dev;                       // this rank's list on the device (column remap)
  DBuf<int> send_row, send_peer, send_dst;  // own row -> slot send_dst of rank send_peer (sorted by peer, row)
  int n_send = 0;
  DBuf<unsigned> smask;                     // fused-push plan (see PushSpec), built from the send list
  DBuf<int> sbase, uptr, upeer, udst, slist;
  int n_slist = 0;
  std::vector<int> h_urow, h_uptr, h_upeer, h_udst;   // host copy of the plan: send rows ascending, their destinations
  int n_to = 0, n_from = 0;
  signed char to[kMaxRanks] = {0}, from[kMaxRanks] = {0};
  bool gather = false;                      // replicated space: vectors are full length, global numbering,
                                            // every rank computes its block and stores it into all peers
  int rank_of(int64_t g) const { return (int)(std::upper_bound(split.begin(), split.end(), g) - split.begin()) - 1; }
};

struct DVec {                               // a vector inside the arena: [own | halo], `stride` doubles per row
  double* p = nullptr;
  size_t off = 0;                           // byte offset inside every rank's arena
  int channel = -1;
  int stride = 1;
  const Space* sp = nullptr;
};

struct DistCtx {
  int rank = 0, world = 1;
  DBuf<char> arena;
  size_t arena_bytes = 0, arena_used = 0;
  void* peer_base[kMaxRanks] = {nullptr};
  bool peer_opened[kMaxRanks] = {false};
  DBuf<unsigned long long> seq;
  DBuf<int> err;
  DBuf<unsigned> done_ctr;
  DBuf<int> flags;                          // {converged, iterations, -, -} of the running solve
  DBuf<unsigned long long> trace;
  DBuf<unsigned> trace_n;
  int next_channel = 0;
  Comm comm;
  bool connected = false;
  ~DistCtx() {
    for (int q = 0; q < kMaxRanks; ++q)
      if (peer_opened[q]) cudaIpcCloseMemHandle(peer_base[q]);
  }
  void init(int rank_, int world_, size_t vector_bytes);
  DVec carve(const Space& sp, int stride);
  void connect(const void* all_handles);    // world x 64-byte IPC handles (own entry ignored)
  void push(const DVec& v);                 // k_halo_push: own boundary rows -> the neighbours' halo slots + flags
  PushSpec push_spec(const DVec& v) const;  // the same exchange, performed by the kernel that produces v (stride 1)
  void wait(const DVec& v);                 // stand-alone wait kernel (for consumers without a built-in wait)
  HaloWait wait_of(const DVec* a, const DVec* b = nullptr) const;
  void check(const char* where);            // throws FS_ERR_INTERNAL if a wait has timed out (synchronises)
};

// sorted unique (rank, global id) pairs read across block boundaries, accumulated from device matrices and host lists
struct HaloCollector {
  std::vector<unsigned long long> keys;     // rank << 32 | id
  // rows of M split by rsplit; columns [c_lo, c_hi) (shifted by -c_lo) live in a space split by csplit
  void add_matrix(const CsrView& M, const std::vector<int64_t>& rsplit, const std::vector<int64_t>& csplit, int64_t c_lo,
                  int64_t c_hi);
  void add(int rank, int64_t id) { keys.push_back(((unsigned long long)(unsigned)rank << 32) | (unsigned)id); }
  void finish(Space& sp, int rank, int world);   // fills halo_all, n_halo, cap, halo_dev, send lists, to / from
};

// rows [r0, r1) of a global CSR matrix as a local one.  Columns < nsplit_glob belong to space A, the others
// (shifted by -nsplit_glob) to space B (null: kept as they are, shifted to start at A's local width).  Entry order
// inside a row is preserved (same summation order as the global matrix).
void extract_rows(const CsrView& G, int64_t r0, int64_t r1, const Space& A, int64_t nsplit_glob, const Space* B, fs_csr& out);

}  // namespace fs
struct fs_stokes;
namespace fs {
void stokes_global_view(fs_stokes* s, fs_mesh** mesh, fs_csr** a_visc, fs_csr** k_red, double* DT, double* nu, std::vector<int>& dof);

// ---- partitioned AMG cycle (amg.cu) ----
struct AmgPartSpec {
  int rank = 0, world = 1;
  std::vector<int64_t> split0;    // world+1 row split of the fine operator
  int gather_rows = 100000;       // coarser levels with at most this many rows are replicated
};
int amg_part_levels(const Amg* amg);                                        // Lp: number of partitioned levels
const std::vector<int64_t>& amg_part_split(const Amg* amg, int level);      // row split of level 0..Lp
// halo requirements of the cycle's operators: level 0's go into c0 (the caller owns that space and adds its own
// matrices), the deeper spaces are finished here.  Returns the arena bytes the cycle's level vectors need.
size_t amg_part_collect(Amg* amg, HaloCollector& c0);
// cut this rank's rows out of the global operators (SELL-32 copies), carve the level vectors, free the global forms
void amg_part_finalize(Amg* amg, const Space& space0, DistCtx& ctx);
// one V-cycle z = M^-1 r on this rank's rows; r is an arena vector whose halo has been pushed
int amg_apply_dist(Amg* amg, const DVec& r, double* z, double* rz_part);

// mark the slices of S (built from the local CSR `loc`) that read halo entries: columns in [n_own_a, nsplit) of the
// first vector, or >= nsplit + n_own_b of the second (n_own_b < 0: the second vector has no halo part)
// ... or whose rows are send rows of the output space `out` (may be null); for those the per-lane destinations are
// tabulated next to the slice list (fs_sell::btab) so that the kernel needs no look-up chain before its remote stores
void sell_mark_boundary(fs_sell& S, const fs_csr& loc, int n_own_a, int nsplit, int n_own_b, const Space* out);
int spmv_sell_dist(const fs_sell& S, const double* x, double* y, const double* x2, double* dot_partials, const Comm& c,
                   const HaloWait& w, const PushSpec* push = nullptr, int trace_tag = 0);
void spmv_sub_dist(const CsrView& A, const double* x, double* y, const double* x2, int nsplit, const Comm& c, const HaloWait& w,
                   const PushSpec& ps);
void spmv_sell2_dist(const fs_sell& S, const double* x, double* y, double* dot_partials, const int* done, const Comm& c,
                     const HaloWait& w);

#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long dist_gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void dist_st_release_sys(unsigned long long* a, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long dist_ld_acquire_sys(const unsigned long long* a) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(a) : "memory");
  return v;
}
__device__ __forceinline__ void dist_st_sys_u64(unsigned long long* a, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(a), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long dist_ld_sys_u64(const unsigned long long* a) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(a) : "memory");
  return v;
}
__device__ __forceinline__ void dist_trace(const Comm& c, unsigned tag) {
  if (c.trace && blockIdx.x == 0 && threadIdx.x == 0) {
    const unsigned k = atomicAdd(c.trace_n, 1u);
    if (k < c.trace_cap) { c.trace[2 * k] = tag; c.trace[2 * k + 1] = dist_gtime(); }
  }
}

// the same from thread 0 of the LAST CTA of the grid (an interior CTA of the boundary-first kernels): kernel end
__device__ __forceinline__ void dist_trace_last(const Comm& c, unsigned tag) {
  if (c.trace && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
    const unsigned k = atomicAdd(c.trace_n, 1u);
    if (k < c.trace_cap) { c.trace[2 * k] = tag; c.trace[2 * k + 1] = dist_gtime(); }
  }
}
__device__ __forceinline__ void dist_st_sys_f64(double* a, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(a), "d"(v) : "memory");
}

// current sequence number, one read per CTA
__device__ __forceinline__ unsigned long long dist_seq(const Comm& c) {
  __shared__ unsigned long long s_seq;
  if (threadIdx.x == 0) s_seq = *reinterpret_cast<volatile unsigned long long*>(c.seq);
  __syncthreads();
  return s_seq;
}

// Block until every channel of `w` carries this rank's current sequence number (all threads of the CTA
// call it; returns after a CTA barrier, so every thread may read halo entries afterwards).
__device__ __forceinline__ void halo_wait(const Comm& c, const HaloWait& w) {
  if (w.nch == 0) return;
  const unsigned long long want = dist_seq(c);
  const int t = threadIdx.x;
  if (t < 2 * kMaxRanks) {
    const int k = t / kMaxRanks, j = t % kMaxRanks;
    if (k < w.nch && j < w.n_from[k]) {
      const unsigned long long* f = c.flag_local + (size_t)w.ch[k] * kMaxRanks + w.from[k][j];
      if (dist_ld_acquire_sys(f) < want) {
        const unsigned long long t0 = dist_gtime();
        unsigned spins = 0;
        while (dist_ld_acquire_sys(f) < want) {
          if ((++spins & 255u) == 0) {
            if (*reinterpret_cast<volatile int*>(c.err) != 0) break;
            if (dist_gtime() - t0 > c.timeout_ns) { atomicCAS(c.err, 0, 0x100 | (w.ch[k] << 12) | w.from[k][j]); break; }
          }
        }
      }
    }
  }
  __syncthreads();
}

// v[0..K) (the local sums, identical in every CTA of this rank) -> sums over all ranks, identical everywhere.
// epoch must be the same on all ranks and change parity from one reduction to the next (seq + 1).
template <int K>
__device__ __forceinline__ void rank_allreduce(const Comm& c, double (&v)[K], unsigned long long epoch) {
  static_assert(2 * K <= kRedWords, "mailbox slot too small");
  if (c.world == 1) return;
  __shared__ double s_got[kMaxRanks][K];
  __shared__ double s_sum[K];
  const int t = threadIdx.x;
  const int par = (int)(epoch & 1ull);
  const unsigned long long e32 = (epoch & 0xffffffffull) << 32;
  if (blockIdx.x == 0 && t < c.world) {
    unsigned long long* dst = c.red_peer[t] + ((size_t)par * kMaxRanks + c.rank) * kRedWords;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const unsigned long long bits = (unsigned long long)__double_as_longlong(v[k]);
      dist_st_sys_u64(dst + 2 * k, (bits & 0xffffffffull) | e32);
      dist_st_sys_u64(dst + 2 * k + 1, (bits >> 32) | e32);
    }
  }
  if (t < c.world) {
    const unsigned long long* src = c.red_local + ((size_t)par * kMaxRanks + t) * kRedWords;
    unsigned long long w[2 * K];
    bool ok = false;
    unsigned spins = 0;
    unsigned long long t0 = 0;
    while (!ok) {
      ok = true;
#pragma unroll
      for (int j = 0; j < 2 * K; ++j) { w[j] = dist_ld_sys_u64(src + j); ok = ok && ((w[j] & 0xffffffff00000000ull) == e32); }
      if (!ok && (++spins & 255u) == 0) {
        if (t0 == 0) t0 = dist_gtime();
        if (*reinterpret_cast<volatile int*>(c.err) != 0) break;
        if (dist_gtime() - t0 > c.timeout_ns) { atomicCAS(c.err, 0, 0x200 | t); break; }
      }
    }
#pragma unroll
    for (int k = 0; k < K; ++k)
      s_got[t][k] = __longlong_as_double((long long)((w[2 * k] & 0xffffffffull) | (w[2 * k + 1] << 32)));
  }
  __syncthreads();
  if (t < K) {
    double s = 0.0;
    for (int q = 0; q < c.world; ++q) s += s_got[q][t];
    s_sum[t] = s;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = s_sum[k];
  __syncthreads();
}

// ---- fused push (PushSpec) ----
// the thread that has just produced row `row` of the vector; returns true if anything was stored remotely
__device__ __forceinline__ bool push_row(const PushSpec& ps, int row, double v) {
  if (ps.gather) {
    for (int j = 0; j < ps.n_to; ++j)
      dist_st_sys_f64(reinterpret_cast<double*>(ps.peer_base[ps.to[j]] + ps.vec_off) + (ps.row0 + row), v);
    return ps.n_to > 0;
  }
  const int s = row >> 5, l = row & 31;
  const unsigned m = __ldg(ps.smask + s);
  if (!((m >> l) & 1u)) return false;
  const int idx = __ldg(ps.sbase + s) + __popc(m & ((1u << l) - 1u));
  for (int k = __ldg(ps.uptr + idx); k < __ldg(ps.uptr + idx + 1); ++k)
    dist_st_sys_f64(reinterpret_cast<double*>(ps.peer_base[__ldg(ps.upeer + k)] + ps.vec_off) + __ldg(ps.udst + k), v);
  return true;
}
__device__ __forceinline__ void dist_fence_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ void dist_fence_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// After the send rows have been produced: each of the n_cta CTAs that own send rows calls it (all threads); the one
// that arrives last releases the flags with `stamp`.  No per-thread fences: the CTA barrier orders every thread's
// remote stores before thread 0's release fence (cumulativity), the arrival counter chains the CTAs' fences, and the
// last CTA's fence orders all of them before the flag stores.  (__threadfence_system() is a sequentially-consistent
// fence and was measured at 12-15 us per pushing thread after NVLink stores; acq_rel is what a release needs.)
__device__ __forceinline__ void push_finish(const Comm& c, const PushSpec& ps, bool /*pushed*/, unsigned long long stamp, int n_cta,
                                            unsigned trace_tag = 0) {
  __syncthreads();
  if (threadIdx.x == 0) {
    if (trace_tag) dist_trace(c, trace_tag);
    dist_fence_sys();
    if (trace_tag) dist_trace(c, trace_tag + 1);
    const unsigned prev = atomicAdd(c.done_ctr + 8 + ps.channel, 1u);
    if (prev == (unsigned)n_cta - 1) {
      c.done_ctr[8 + ps.channel] = 0;
      dist_fence_sys();
      for (int j = 0; j < ps.n_to; ++j)
        dist_st_sys_u64(c.flag_peer[ps.to[j]] + (size_t)ps.channel * kMaxRanks + c.rank, stamp);
    }
  }
}

// Called by every CTA at the END of a reducing kernel: true in exactly one thread (thread 0 of the CTA that
// arrives last), after all CTAs of the launch have finished their work.  That thread publishes the launch's
// scalars and must then call dist_seq_bump.
__device__ __forceinline__ bool dist_last_block(const Comm& c, int ctr) {
  __syncthreads();
  bool last = false;
  if (threadIdx.x == 0) {
    dist_fence_gpu();
    const unsigned prev = atomicAdd(c.done_ctr + ctr, 1u);
    if (prev == gridDim.x - 1) { c.done_ctr[ctr] = 0; last = true; dist_fence_gpu(); }
  }
  return last;
}
__device__ __forceinline__ void dist_seq_bump(const Comm& c) {
  dist_fence_gpu();
  atomicAdd(c.seq, 1ull);
}
#endif  // __CUDACC__

}  // namespace fs
