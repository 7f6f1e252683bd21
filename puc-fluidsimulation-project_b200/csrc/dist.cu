// dist.cu -- host side of the partitioned step's communication layer (dist.cuh): the IPC arena, halo
// lists derived from replicated global matrices, row-block extraction, and the push / wait kernels.
#include <cub/cub.cuh>

#include "dist.cuh"

namespace fs {

// ---- block boundaries as a kernel argument ----------------------------------------------------
struct Split9 {
  long long s[kMaxRanks + 1];
  int world;
  __host__ __device__ int rank_of(long long g) const {
    int r = 0;
    while (r + 1 < world && g >= s[r + 1]) ++r;
    return r;
  }
};
static Split9 make_split(const std::vector<int64_t>& v) {
  Split9 s{};
  s.world = (int)v.size() - 1;
  FS_REQUIRE(s.world >= 1 && s.world <= kMaxRanks, "bad split");
  for (int k = 0; k <= s.world; ++k) s.s[k] = v[k];
  return s;
}

// entries whose column (in [c_lo, c_hi)) belongs to another rank than the row: count, then emit (rank << 32 | col - c_lo)
template <bool EMIT>
__global__ void k_cross(CsrView M, Split9 rs, Split9 cs, long long c_lo, long long c_hi, unsigned long long* __restrict__ counter,
                        unsigned long long* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M.n) return;
  const int rr = rs.rank_of(i);
  const long long own_lo = cs.s[rr], own_hi = cs.s[rr + 1];
  int cnt = 0;
  for (int k = M.rowptr[i]; k < M.rowptr[i + 1]; ++k) {
    const long long c = M.colidx[k];
    if (c < c_lo || c >= c_hi) continue;
    const long long g = c - c_lo;
    if (g >= own_lo && g < own_hi) continue;
    if (EMIT) {
      const unsigned long long pos = atomicAdd(counter, 1ull);
      out[pos] = ((unsigned long long)(unsigned)rr << 32) | (unsigned long long)(unsigned)g;
    } else {
      ++cnt;
    }
  }
  if (!EMIT && cnt) atomicAdd(counter, (unsigned long long)cnt);
}

void HaloCollector::add_matrix(const CsrView& M, const std::vector<int64_t>& rsplit, const std::vector<int64_t>& csplit,
                               int64_t c_lo, int64_t c_hi) {
  if (M.n == 0) return;
  FS_REQUIRE(rsplit.back() == M.n, "row split does not cover the matrix");
  cudaStream_t st = stream();
  const Split9 rs = make_split(rsplit), cs = make_split(csplit);
  DBuf<unsigned long long> counter(1);
  counter.zero();
  const int g = div_up(M.n, 256);
  k_cross<false><<<g, 256, 0, st>>>(M, rs, cs, c_lo, c_hi, counter.p, nullptr);
  FS_LAUNCH_CHECK();
  const unsigned long long m = counter.to_host()[0];
  if (!m) return;
  DBuf<unsigned long long> out(m);
  counter.zero();
  k_cross<true><<<g, 256, 0, st>>>(M, rs, cs, c_lo, c_hi, counter.p, out.p);
  FS_LAUNCH_CHECK();
  // duplicates are plentiful (every boundary row repeats its neighbours): compress on the device first
  DBuf<unsigned long long> sorted(m), uniq(m);
  DBuf<unsigned long long> nsel(1);
  size_t b1 = 0, b2 = 0;
  FS_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, b1, out.p, sorted.p, (int)m, 0, 64, st));
  FS_CUDA(cub::DeviceSelect::Unique(nullptr, b2, sorted.p, uniq.p, nsel.p, (int)m, st));
  DBuf<char> tmp(std::max(b1, b2));
  FS_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, b1, out.p, sorted.p, (int)m, 0, 64, st));
  FS_CUDA(cub::DeviceSelect::Unique(tmp.p, b2, sorted.p, uniq.p, nsel.p, (int)m, st));
  count_launch(4);
  const size_t nu = (size_t)nsel.to_host()[0];
  const size_t old = keys.size();
  keys.resize(old + nu);
  FS_CUDA(cudaMemcpyAsync(keys.data() + old, uniq.p, nu * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  FS_CUDA(cudaStreamSynchronize(st));
}

void HaloCollector::finish(Space& sp, int rank, int world) {
  FS_REQUIRE((int)sp.split.size() == world + 1, "space split not set");
  std::sort(keys.begin(), keys.end());
  keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
  sp.halo_all.assign(world, {});
  for (unsigned long long k : keys) {
    const int q = (int)(k >> 32);
    const int id = (int)(k & 0xffffffffu);
    FS_REQUIRE(q >= 0 && q < world, "halo key with a bad rank");
    if (id >= sp.split[q] && id < sp.split[q + 1]) continue;   // own entry: not a halo
    sp.halo_all[q].push_back(id);
  }
  sp.own_lo = sp.split[rank];
  sp.n_own = sp.split[rank + 1] - sp.split[rank];
  sp.n_halo = (int64_t)sp.halo_all[rank].size();
  sp.cap = 0;
  for (int q = 0; q < world; ++q)
    sp.cap = std::max<int64_t>(sp.cap, sp.split[q + 1] - sp.split[q] + (int64_t)sp.halo_all[q].size());
  sp.cap = (sp.cap + 31) / 32 * 32;
  if (sp.n_halo) {
    sp.halo_dev.alloc(sp.n_halo);
    sp.halo_dev.upload(sp.halo_all[rank].data(), sp.n_halo);
  }
  // what the others read from my block, and who I read from
  std::vector<int> rows, peers, dsts;
  bool is_to[kMaxRanks] = {false}, is_from[kMaxRanks] = {false};
  for (int q = 0; q < world; ++q) {
    if (q == rank) continue;
    const std::vector<int>& h = sp.halo_all[q];
    const int64_t nq = sp.split[q + 1] - sp.split[q];
    auto lo = std::lower_bound(h.begin(), h.end(), (int)sp.split[rank]);
    auto hi = std::lower_bound(h.begin(), h.end(), (int)sp.split[rank + 1]);
    for (auto it = lo; it != hi; ++it) {
      rows.push_back(*it - (int)sp.own_lo);
      peers.push_back(q);
      dsts.push_back((int)(nq + (it - h.begin())));
      is_to[q] = true;
    }
  }
  for (int id : sp.halo_all[rank]) is_from[sp.rank_of(id)] = true;
  sp.n_send = (int)rows.size();
  sp.n_to = sp.n_from = 0;
  for (int q = 0; q < world; ++q) {
    if (is_to[q]) sp.to[sp.n_to++] = (signed char)q;
    if (is_from[q]) sp.from[sp.n_from++] = (signed char)q;
  }
  if (sp.n_send) {
    sp.send_row.alloc(sp.n_send); sp.send_row.upload(rows.data(), sp.n_send);
    sp.send_peer.alloc(sp.n_send); sp.send_peer.upload(peers.data(), sp.n_send);
    sp.send_dst.alloc(sp.n_send); sp.send_dst.upload(dsts.data(), sp.n_send);
  }
  {
    // fused-push plan: send entries grouped by own row, rows described per 32-row slice
    const int nsl = (int)((sp.n_own + 31) / 32);
    std::vector<int> order(sp.n_send);
    for (int k = 0; k < sp.n_send; ++k) order[k] = k;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return rows[a] < rows[b]; });
    std::vector<unsigned> smask(std::max(nsl, 1), 0u);
    std::vector<int> sbase(std::max(nsl, 1), 0), uptr(1, 0), upeer(sp.n_send), udst(sp.n_send);
    int prev_row = -1, nu = 0;
    for (int k = 0; k < sp.n_send; ++k) {
      const int e = order[k], r = rows[e];
      if (r != prev_row) {
        if (!smask[r >> 5]) sbase[r >> 5] = nu;
        smask[r >> 5] |= 1u << (r & 31);
        if (prev_row >= 0) uptr.push_back(k);
        ++nu;
        prev_row = r;
      }
      upeer[k] = peers[e];
      udst[k] = dsts[e];
    }
    uptr.push_back(sp.n_send);
    sp.h_urow.clear();
    for (int k = 0, pr = -1; k < sp.n_send; ++k) { const int r = rows[order[k]]; if (r != pr) { sp.h_urow.push_back(r); pr = r; } }
    sp.h_uptr = uptr; sp.h_upeer = upeer; sp.h_udst = udst;
    std::vector<int> slist;
    for (int sl = 0; sl < nsl; ++sl) if (smask[sl]) slist.push_back(sl);
    sp.n_slist = (int)slist.size();
    if (sp.n_slist) { sp.slist.alloc(slist.size()); sp.slist.upload(slist.data(), slist.size()); }
    sp.smask.alloc(smask.size()); sp.smask.upload(smask.data(), smask.size());
    sp.sbase.alloc(sbase.size()); sp.sbase.upload(sbase.data(), sbase.size());
    sp.uptr.alloc(uptr.size()); sp.uptr.upload(uptr.data(), uptr.size());
    if (sp.n_send) {
      sp.upeer.alloc(sp.n_send); sp.upeer.upload(upeer.data(), sp.n_send);
      sp.udst.alloc(sp.n_send); sp.udst.upload(udst.data(), sp.n_send);
    }
    FS_CUDA(cudaStreamSynchronize(stream()));   // the host vectors above go out of scope
  }
  FS_CUDA(cudaStreamSynchronize(stream()));
  keys.clear();
  keys.shrink_to_fit();
}

// ---- row-block extraction ---------------------------------------------------------------------
struct RemapSpace {
  long long own_lo, own_hi;   // global range of the owned block
  int n_own, n_halo;
  const int* halo;            // sorted global ids
  __device__ int local(long long g) const {
    if (g >= own_lo && g < own_hi) return (int)(g - own_lo);
    int lo = 0, hi = n_halo;   // first index with halo[idx] >= g
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (halo[mid] < g) lo = mid + 1; else hi = mid; }
    return (lo < n_halo && halo[lo] == g) ? n_own + lo : -1;
  }
};
static RemapSpace remap_of(const Space& s) {
  return RemapSpace{(long long)s.own_lo, (long long)(s.own_lo + s.n_own), (int)s.n_own, (int)s.n_halo, s.halo_dev.p};
}

__global__ void k_extract_rowptr(const int* __restrict__ grow, int r0, int n, int* __restrict__ rowptr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= n) rowptr[i] = grow[r0 + i] - grow[r0];
}
__global__ void k_extract_fill(CsrView G, int r0, int n, RemapSpace A, long long nsplit_glob, RemapSpace B, int b_identity,
                               int* __restrict__ cols, double* __restrict__ vals, int* __restrict__ bad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int base = G.rowptr[r0];
  const int a_width = A.n_own + A.n_halo;
  for (int k = G.rowptr[r0 + i]; k < G.rowptr[r0 + i + 1]; ++k) {
    const long long c = G.colidx[k];
    int lc;
    if (c < nsplit_glob) lc = A.local(c);
    else if (b_identity) lc = a_width + (int)(c - nsplit_glob);
    else { const int t = B.local(c - nsplit_glob); lc = t < 0 ? -1 : a_width + t; }
    if (lc < 0) { atomicAdd(bad, 1); lc = 0; }
    cols[k - base] = lc;
    vals[k - base] = G.vals[k];
  }
}

void extract_rows(const CsrView& G, int64_t r0, int64_t r1, const Space& A, int64_t nsplit_glob, const Space* B, fs_csr& out) {
  cudaStream_t st = stream();
  const int n = (int)(r1 - r0);
  FS_REQUIRE(n > 0 && r0 >= 0 && r1 <= G.n, "extract_rows: bad row range");
  int ends[2];
  FS_CUDA(cudaMemcpyAsync(&ends[0], G.rowptr + r0, sizeof(int), cudaMemcpyDeviceToHost, st));
  FS_CUDA(cudaMemcpyAsync(&ends[1], G.rowptr + r1, sizeof(int), cudaMemcpyDeviceToHost, st));
  FS_CUDA(cudaStreamSynchronize(st));
  const int64_t nnz = ends[1] - ends[0];
  out.n = n;
  out.nnz = nnz;
  out.rowptr_own.alloc(n + 1);
  out.colidx_own.alloc(std::max<int64_t>(nnz, 1));
  out.vals.alloc(std::max<int64_t>(nnz, 1));
  out.rowptr = out.rowptr_own.p;
  out.colidx = out.colidx_own.p;
  DBuf<int> bad(1);
  bad.zero();
  k_extract_rowptr<<<div_up(n + 1, 256), 256, 0, st>>>(G.rowptr, (int)r0, n, out.rowptr_own.p);
  FS_LAUNCH_CHECK();
  const RemapSpace ra = remap_of(A);
  const RemapSpace rb = (B && !B->gather) ? remap_of(*B) : RemapSpace{0, 0, 0, 0, nullptr};
  k_extract_fill<<<div_up(n, 256), 256, 0, st>>>(G, (int)r0, n, ra, (long long)nsplit_glob, rb, (!B || B->gather) ? 1 : 0,
                                                 out.colidx_own.p, out.vals.p, bad.p);
  FS_LAUNCH_CHECK();
  const int nbad = bad.to_host()[0];
  if (nbad) throw Error(FS_ERR_INTERNAL, "extract_rows: " + std::to_string(nbad) + " columns are neither owned nor in the halo list");
}

// ---- push / wait kernels ------------------------------------------------------------------------
// Several CTAs store; the one that arrives last (all stores fenced) releases the flags.
__global__ void __launch_bounds__(512) k_halo_push(PushArgs a, Comm c) {
  if (c.done && *c.done) return;
  const unsigned long long seq = dist_seq(c);
  const int gt = blockIdx.x * blockDim.x + threadIdx.x, gn = gridDim.x * blockDim.x;
  if (a.gather) {
    const long long total = (long long)a.n_rows * a.stride;
    const double* src = a.src + (size_t)a.row0 * a.stride;
    for (int j = 0; j < a.n_to; ++j) {
      double* dst = reinterpret_cast<double*>(a.peer_base[a.to[j]] + a.vec_off) + (size_t)a.row0 * a.stride;
      for (long long k = gt; k < total; k += gn) dist_st_sys_f64(dst + k, __ldcg(src + k));
    }
  } else {
    for (int k = gt; k < a.n_send; k += gn) {
      const double* src = a.src + (size_t)a.send_row[k] * a.stride;
      double* dst = reinterpret_cast<double*>(a.peer_base[a.send_peer[k]] + a.vec_off) + (size_t)a.send_dst[k] * a.stride;
      for (int s = 0; s < a.stride; ++s) dist_st_sys_f64(dst + s, __ldcg(src + s));
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {       // see push_finish (dist.cuh) for the ordering argument
    dist_fence_sys();
    const unsigned prev = atomicAdd(c.done_ctr + 8 + a.channel, 1u);
    if (prev == gridDim.x - 1) {
      c.done_ctr[8 + a.channel] = 0;
      dist_fence_sys();
      for (int j = 0; j < a.n_to; ++j)
        dist_st_sys_u64(c.flag_peer[a.to[j]] + (size_t)a.channel * kMaxRanks + c.rank, seq);
    }
  }
}

__global__ void k_halo_wait(Comm c, HaloWait w) {
  if (c.done && *c.done) return;
  halo_wait(c, w);
}

// FS_DIST_DEBUG_SKIP (timing experiments only; results are wrong): bit 0 = no halo pushes, bit 1 = no halo waits,
// bit 2 = reductions stay local
static int debug_skip() {
  static const int v = [] { const char* e = std::getenv("FS_DIST_DEBUG_SKIP"); return e ? std::atoi(e) : 0; }();
  return v;
}

void DistCtx::init(int rank_, int world_, size_t vector_bytes) {
  FS_REQUIRE(world_ >= 1 && world_ <= kMaxRanks && rank_ >= 0 && rank_ < world_, "bad rank / world (at most 8 ranks)");
  rank = rank_;
  world = world_;
  arena_bytes = kCtlBytes + vector_bytes + 4096;
  arena.alloc(arena_bytes);
  arena.zero();
  arena_used = kCtlBytes;
  seq.alloc(1); seq.zero();
  err.alloc(1); err.zero();
  done_ctr.alloc(8 + kMaxChannels); done_ctr.zero();
  flags.alloc(4); flags.zero();
  peer_base[rank] = arena.p;
  comm = Comm{};
  comm.rank = rank;
  comm.world = world;
  comm.seq = seq.p;
  comm.err = err.p;
  comm.done_ctr = done_ctr.p;
  comm.done = flags.p;
  comm.red_local = reinterpret_cast<unsigned long long*>(arena.p + kCtlRedOff);
  comm.flag_local = reinterpret_cast<unsigned long long*>(arena.p + kCtlFlagOff);
  comm.red_peer[rank] = comm.red_local;
  comm.flag_peer[rank] = comm.flag_local;
  if (const char* e = std::getenv("FS_DIST_TRACE")) {
    if (std::atoi(e) > 0) {
      comm.trace_cap = 1u << 16;
      trace.alloc(2 * (size_t)comm.trace_cap);
      trace_n.alloc(1);
      trace_n.zero();
      comm.trace = trace.p;
      comm.trace_n = trace_n.p;
    }
  }
  if (const char* e = std::getenv("FS_DIST_TIMEOUT_MS")) comm.timeout_ns = (unsigned long long)std::atof(e) * 1000000ull;
  connected = (world == 1);
  FS_CUDA(cudaStreamSynchronize(stream()));
}

DVec DistCtx::carve(const Space& sp, int stride) {
  const size_t bytes = ((size_t)sp.cap * stride * sizeof(double) + 255) / 256 * 256;
  FS_REQUIRE(arena_used + bytes <= arena_bytes, "arena too small");
  FS_REQUIRE(next_channel < kMaxChannels, "out of halo channels");
  DVec v;
  v.off = arena_used;
  v.p = reinterpret_cast<double*>(arena.p + arena_used);
  v.channel = next_channel++;
  v.stride = stride;
  v.sp = &sp;
  arena_used += bytes;
  return v;
}

void DistCtx::connect(const void* all_handles) {
  FS_REQUIRE(all_handles || world == 1, "handles missing");
  for (int q = 0; q < world; ++q) {
    if (q == rank) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, (const char*)all_handles + 64 * q, 64);
    FS_CUDA(cudaIpcOpenMemHandle(&peer_base[q], h, cudaIpcMemLazyEnablePeerAccess));
    peer_opened[q] = true;
    comm.red_peer[q] = reinterpret_cast<unsigned long long*>((char*)peer_base[q] + kCtlRedOff);
    comm.flag_peer[q] = reinterpret_cast<unsigned long long*>((char*)peer_base[q] + kCtlFlagOff);
  }
  connected = true;
  if (debug_skip() & 4) comm.world = 1;
}

void DistCtx::push(const DVec& v) {
  const Space& sp = *v.sp;
  if (world == 1 || sp.n_to == 0 || (debug_skip() & 1)) return;
  PushArgs a;
  a.src = v.p;
  a.stride = v.stride;
  a.channel = v.channel;
  a.vec_off = v.off;
  for (int q = 0; q < world; ++q) a.peer_base[q] = (char*)peer_base[q];
  a.n_to = sp.n_to;
  for (int k = 0; k < sp.n_to; ++k) a.to[k] = sp.to[k];
  int grid;
  if (sp.gather) {
    a.gather = 1;
    a.row0 = (int)sp.own_lo;
    a.n_rows = (int)sp.n_own;
    grid = std::max(1, std::min(16, div_up((int64_t)sp.n_own * v.stride * sp.n_to, 4096)));
  } else {
    a.n_send = sp.n_send;
    a.send_row = sp.send_row.p;
    a.send_peer = sp.send_peer.p;
    a.send_dst = sp.send_dst.p;
    grid = std::max(1, std::min(8, div_up((int64_t)sp.n_send * v.stride, 2048)));
  }
  k_halo_push<<<grid, 512, 0, stream()>>>(a, comm);
  FS_LAUNCH_CHECK();
}

PushSpec DistCtx::push_spec(const DVec& v) const {
  PushSpec ps;
  const Space& sp = *v.sp;
  if (world == 1 || sp.n_to == 0 || (debug_skip() & 1)) return ps;
  FS_REQUIRE(v.stride == 1, "fused pushes are for scalar vectors");
  ps.enabled = 1;
  ps.channel = v.channel;
  ps.vec_off = v.off;
  for (int q = 0; q < world; ++q) ps.peer_base[q] = (char*)peer_base[q];
  ps.n_to = sp.n_to;
  for (int k = 0; k < sp.n_to; ++k) ps.to[k] = sp.to[k];
  if (sp.gather) {
    ps.gather = 1;
    ps.row0 = (int)sp.own_lo;
  } else {
    ps.smask = sp.smask.p; ps.sbase = sp.sbase.p; ps.uptr = sp.uptr.p; ps.upeer = sp.upeer.p; ps.udst = sp.udst.p;
    ps.slist = sp.slist.p; ps.n_slist = sp.n_slist;
  }
  return ps;
}

HaloWait DistCtx::wait_of(const DVec* a, const DVec* b) const {
  HaloWait w;
  if (world == 1 || (debug_skip() & 2)) return w;
  for (const DVec* v : {a, b}) {
    if (!v || !v->sp || v->sp->n_from == 0) continue;
    const int k = w.nch++;
    w.ch[k] = v->channel;
    w.n_from[k] = v->sp->n_from;
    for (int j = 0; j < v->sp->n_from; ++j) w.from[k][j] = v->sp->from[j];
  }
  return w;
}

void DistCtx::wait(const DVec& v) {
  const HaloWait w = wait_of(&v);
  if (!w.nch) return;
  k_halo_wait<<<1, 32, 0, stream()>>>(comm, w);
  FS_LAUNCH_CHECK();
}

void DistCtx::check(const char* where) {
  int e = 0;
  FS_CUDA(cudaMemcpyAsync(&e, err.p, sizeof(int), cudaMemcpyDeviceToHost, stream()));
  FS_CUDA(cudaStreamSynchronize(stream()));
  if (!e) return;
  char msg[200];
  std::snprintf(msg, sizeof(msg), "%s: rank %d timed out waiting for %s of rank %d (code 0x%x): peer kernel not running or peer memory unreachable",
                where, rank, (e & 0x200) ? "a reduction word" : "a halo flag", e & 0xff, e);
  throw Error(FS_ERR_INTERNAL, msg);
}

}  // namespace fs
