// internal.cuh -- shared plumbing of libfluidsim (device buffers, host/device
// pointer staging, error translation, reductions).  Not part of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <memory>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/fluidsim.h"

namespace fs {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string& s);
cudaStream_t stream();          // current library stream of this thread
void count_launch(int n = 1);   // bumps fs_launch_count()

#define FS_CUDA(call)                                                                       \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess)                                                                 \
      throw fs::Error(FS_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__) +    \
                                       " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); \
  } while (0)

#define FS_REQUIRE(cond, msg)                                  \
  do {                                                         \
    if (!(cond)) throw fs::Error(FS_ERR_ARG, std::string(msg)); \
  } while (0)

#define FS_API_BEGIN try {
#define FS_API_END                                  \
  return FS_OK;                                     \
  }                                                 \
  catch (const fs::Error& e) {                      \
    fs::set_last_error(e.what());                   \
    return e.code;                                  \
  }                                                 \
  catch (const std::exception& e) {                 \
    fs::set_last_error(e.what());                   \
    return FS_ERR_INTERNAL;                         \
  }                                                 \
  catch (...) {                                     \
    fs::set_last_error("unknown C++ exception");    \
    return FS_ERR_INTERNAL;                         \
  }

#define FS_LAUNCH_CHECK()          \
  do {                             \
    fs::count_launch();            \
    FS_CUDA(cudaGetLastError());   \
  } while (0)

// ---------------------------------------------------------------- programmatic dependent launch
// The kernels of one PCG iteration form a chain of ~12 dependent launches, most of them latency bound.  Launched with
// the programmatic-stream-serialization attribute, a kernel may start while its predecessor is still running: its CTAs
// become resident as the predecessor's retire, do whatever does not depend on the predecessor (index arithmetic, the
// slice / row headers, L2 prefetches of the matrix stream they are about to read) and then block in pdl_wait() until the
// predecessor has completed and its writes are visible.  Rules kept by every kernel launched through launch_pdl:
// pdl_launch() first, pdl_wait() before the first read of anything an earlier kernel writes AND before the first global
// write (an earlier kernel may still be reading what this one overwrites).  Both are no-ops in a normal launch.
// Stream capture turns the attribute into programmatic edges of the V-cycle's CUDA graph.  FS_PDL=0: plain launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
bool pdl_enabled();

template <class... KArgs, class... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream();
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  FS_CUDA(cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...));
}

// ---------------------------------------------------------------- device buffer
template <class T>
struct DBuf {
  T* p = nullptr;
  size_t n = 0;
  DBuf() = default;
  explicit DBuf(size_t count) { alloc(count); }
  DBuf(const DBuf&) = delete;
  DBuf& operator=(const DBuf&) = delete;
  DBuf(DBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DBuf& operator=(DBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DBuf() { release(); }
  void alloc(size_t count) {
    release();
    n = count;
    if (count) FS_CUDA(cudaMalloc(&p, count * sizeof(T)));
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  void zero() { if (n) FS_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), stream())); }
  void upload(const T* src, size_t count) {  // src host or device
    FS_REQUIRE(count <= n, "DBuf::upload overflow");
    if (count) FS_CUDA(cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyDefault, stream()));
  }
  std::vector<T> to_host() const {
    std::vector<T> h(n);
    if (n) {
      FS_CUDA(cudaMemcpyAsync(h.data(), p, n * sizeof(T), cudaMemcpyDeviceToHost, stream()));
      FS_CUDA(cudaStreamSynchronize(stream()));
    }
    return h;
  }
};

inline bool is_device_ptr(const void* p) {
  if (!p) return false;
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Read-only argument that may live on the host: staged into a stream-ordered
// temporary when needed.
template <class T>
struct In {
  const T* d = nullptr;
  T* tmp = nullptr;
  In(const T* ptr, size_t count) {
    if (!ptr || count == 0) return;
    if (is_device_ptr(ptr)) { d = ptr; return; }
    FS_CUDA(cudaMallocAsync(&tmp, count * sizeof(T), stream()));
    FS_CUDA(cudaMemcpyAsync(tmp, ptr, count * sizeof(T), cudaMemcpyHostToDevice, stream()));
    d = tmp;
  }
  ~In() { if (tmp) cudaFreeAsync(tmp, stream()); }
  In(const In&) = delete;
  In& operator=(const In&) = delete;
};

// Output (or in/out) argument that may live on the host.  commit() copies back
// and synchronises; call it last in the API function.
template <class T>
struct Out {
  T* d = nullptr;
  T* tmp = nullptr;
  T* host = nullptr;
  size_t count = 0;
  Out(T* ptr, size_t cnt, bool read_in = false) : count(cnt) {
    if (!ptr || cnt == 0) return;
    if (is_device_ptr(ptr)) { d = ptr; return; }
    host = ptr;
    FS_CUDA(cudaMallocAsync(&tmp, cnt * sizeof(T), stream()));
    if (read_in) FS_CUDA(cudaMemcpyAsync(tmp, ptr, cnt * sizeof(T), cudaMemcpyHostToDevice, stream()));
    d = tmp;
  }
  void commit() {
    if (tmp && host)
      FS_CUDA(cudaMemcpyAsync(host, tmp, count * sizeof(T), cudaMemcpyDeviceToHost, stream()));
  }
  ~Out() { if (tmp) cudaFreeAsync(tmp, stream()); }
  Out(const Out&) = delete;
  Out& operator=(const Out&) = delete;
};

inline void sync() { FS_CUDA(cudaStreamSynchronize(stream())); }

inline int div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
int sm_count();

// ---------------------------------------------------------------- structures
struct CsrView {
  int n;
  int64_t nnz;
  const int* rowptr;
  const int* colidx;
  const double* vals;
  int tile_nnz_max = 0;   // max nonzeros in a 256-row tile (0: not computed -> vector SpMV)
  int wtile_nnz_max = 0;  // max nonzeros in a 32-row (one warp) tile
  const float* vals32 = nullptr;   // optional fp32 copy of the values (AMG V-cycle only)
};

}  // namespace fs

namespace fs { struct Amg; void amg_free(Amg*); struct fs_sell; void sell_free(fs_sell*); }

// Opaque handle definitions ----------------------------------------------------
struct fs_csr {
  int64_t n = 0, nnz = 0;
  fs::DBuf<int> rowptr_own, colidx_own;
  const int* rowptr = nullptr;   // either own or borrowed from a mesh pattern
  const int* colidx = nullptr;
  fs::DBuf<double> vals;
  fs::DBuf<double> dinv;         // Jacobi (lazy)
  bool dinv_ready = false;
  // Krylov workspace (lazy, sized n*nrhs)
  fs::DBuf<double> ws;
  size_t ws_n = 0;
  fs::DBuf<double> partials;
  fs::DBuf<double> scal;
  int tile_nnz_max = -1;         // lazily computed by fs::ensure_tiles
  int wtile_nnz_max = 0;
  fs::Amg* amg = nullptr;        // aggregation-AMG hierarchy (lazy, FS_PRECOND_AMG)
  fs_csr() = default;
  fs_csr(const fs_csr&) = delete;
  fs_csr& operator=(const fs_csr&) = delete;
  int pcg_hint = 0, pcg_hint_prev = 0;   // iterations of the last two converged AMG-PCG solves (lagged convergence polling)
  fs::fs_sell* sell64 = nullptr; // SELL-32 fp64 copy for the AMG-preconditioned CG's A*p (lazy)
  ~fs_csr() { if (amg) fs::amg_free(amg); if (sell64) fs::sell_free(sell64); }
  fs::DBuf<float> vals32;        // fp32 copy of the values for the mixed-precision V-cycle (lazy)
  fs::CsrView view() const {
    return fs::CsrView{(int)n, nnz, rowptr, colidx, vals.p, tile_nnz_max > 0 ? tile_nnz_max : 0, wtile_nnz_max, nullptr};
  }
  fs::CsrView view32() const {   // same matrix, values streamed as fp32 where the kernel supports it
    fs::CsrView v = view();
    v.vals32 = vals32.n ? vals32.p : nullptr;
    return v;
  }
};

namespace fs {
// Structural pattern built from (mapped) triangles.
struct Pattern {
  int64_t n = 0, nnz = 0, T = 0;
  DBuf<int> rowptr, colidx;
  DBuf<int> scatter;     // (T,9) element entry -> nnz id
  DBuf<int> seg_start;   // (nnz+1) start of each nonzero's contribution list
  DBuf<unsigned> contrib; // (9T) element-entry ids (e*9+k), sorted by nnz, ascending e inside
};
void build_pattern(const int* d_tris, int64_t T, int64_t n, const int* d_dof /*nullable*/, Pattern& out);
void assemble_on_pattern(const Pattern& pat, const double* d_ke /* (T,9) */, double* d_vals);
}  // namespace fs

namespace fs { struct Locator; struct Amg; void amg_free(Amg*); }

struct fs_mesh {
  int64_t N = 0, T = 0;
  // Sub-mesh of a partitioned run: only nodes [0, n_active) are owned; the node passes of divergence /
  // gradient evaluate (and write) those alone -- the remaining nodes are halo copies whose incident
  // elements are incomplete and whose vector entries belong to other ranks.  -1: all nodes.
  int64_t n_active = -1;
  int64_t n_eval() const { return n_active >= 0 ? n_active : N; }
  fs::DBuf<double> coords;    // (N,2)
  fs::DBuf<int> tris;         // (T,3)
  fs::DBuf<int> markers;      // (N)
  fs::Pattern pat;            // pattern of K on the N nodes
  // node -> incident element-corner list (ascending element id)
  fs::DBuf<int> inc_ptr;      // (N+1)
  fs::DBuf<unsigned> inc;     // (3T) entries e*3+i
  // per-element scratch
  fs::DBuf<double> ke;        // (T,9) lazily allocated
  fs::DBuf<double> elem_a, elem_b, elem_c;  // (T) scratch for lumps
  fs::DBuf<double> area_sum;  // (N) sum of area/3 over non-degenerate incident elements
  fs::DBuf<double> mass;      // (N) lumped mass (no skip)
  bool geom_ready = false;
  // boundary sets
  fs::DBuf<int> wall, inner, pairs, interior;
  fs::DBuf<double> inner_sin, inner_cos, inner_sin2;
  int64_t n_wall = 0, n_inner = 0, n_pairs = 0, n_interior = 0;
  std::vector<int> pairs_host;
  bool bc_ready = false;
  // locator (lazy)
  fs::Locator* loc = nullptr;
  ~fs_mesh();
};

namespace fs {
void ensure_geom(fs_mesh* m);
void divergence_dev(fs_mesh* m, const double* d_u, double* d_div, double* d_div_sum /*nullable*/);
void gradient_dev(fs_mesh* m, const double* d_p, double* d_gx, double* d_gy);
void per_bcu_dev(fs_mesh* m, double* d_u);
void dir_bcu_dev(fs_mesh* m, double* d_u, double B1, double B2);
void rot_bcu_dev(fs_mesh* m, double* d_u, double omega, double cx, double cy);
void grad_update_dev(fs_mesh* m, const double* d_p, const double* d_ui, double* d_uo, double DT,
                     const unsigned char* d_interior_flag /* null: all nodes */);
void divergence_batch_dev(fs_mesh* m, int B, const double* d_u, double* d_div, double* d_lump);
void grad_update_batch_dev(fs_mesh* m, int B, const double* d_p, const double* d_ui, double* d_uo, double DT,
                           const unsigned char* d_interior_flag, double* d_lx, double* d_ly);
void bcu_batch_dev(fs_mesh* m, int B, double* d_u, const double* d_b12);   // makePerBCU + makeDirBCU with (B1, B2) per configuration
void jacobi_prepare(fs_csr* a);
void ensure_tiles(fs_csr* a);
// warp-granular SpMV with fused epilogues (spmv_warp.cu); returns the grid used or 0 if unsupported
enum { EPI_AX = 0, EPI_RESID = 1, EPI_JACOBI = 2, EPI_PRESM = 3, EPI_ADD = 4, EPI_AX2 = 5, EPI_AXS = 6 };
int spmv_warp(const CsrView& A, int epi, const double* x, double* y, const double* b, const double* dinv, double w,
              double* xout, double* dot_partials, const double* x2 = nullptr, int nsplit = 0);
// SELL-32 copy of a CSR matrix (spmv_sell.cu): slices of 32 rows, column-major, padded per slice
struct fs_sell {
  int n = 0, nslices = 0;
  long long padded = 0, nnz = 0;
  int nsplit = -1;        // >= 0: split form (see spmv_sell.cu)
  DBuf<long long> sptr;   // nslices + 1 element offsets (multiples of 32)
  DBuf<int> wg;           // split form: per-slice width of the first part
  DBuf<int> cols;
  DBuf<float> v32;        // exactly one of v32 / v64 is filled
  DBuf<double> v64;
  // packed form of an fp32 operator (FS_SELL_PACK, default on): one 32-bit word per entry = fp16 value (scaled by the power
  // of two 1 / pk_inv) << 16 | 16-bit column offset from the slice's base column cbase[s] (.x first part, .y second part of a
  // split slice).  Same element offsets as cols / v32.  A slice whose columns span more than 65535 keeps cbase[s].x = INT_MIN
  // and is read from cols / v32, which then hold the SAME fp16-rounded values: both encodings give the same bits.
  DBuf<unsigned> pk;
  DBuf<int2> cbase;
  double pk_inv = 1.0;
  long long pk_unpacked = 0;   // slices left in the 32-bit encoding
  DBuf<int> perm;         // SELL-C-sigma: slot (32 s + lane) holds row perm[slot] (rows sorted by length inside windows
                          // of sigma rows, so a slice pads to its own longest row only); empty: identity
  // partitioned step: slices that read halo entries of their input vectors (bit per slice + list); the DIST kernel
  // does all other slices first and waits for the neighbours' halo flags only before these
  DBuf<unsigned> bmask;
  DBuf<int> blist;
  int n_blist = 0;
  DBuf<int2> btab;        // per boundary slice and lane: up to two destinations (rank << 28 | slot; -1 none; .y = -2: look up)
};
// pk_maxabs: largest |value| that fixes the fp16 scale of the packed form (0: take it from A; the partitioned set-up
// passes the global operator's so that every rank rounds exactly like the single-GPU hierarchy)
void sell_build(const fs_csr& A, bool f32, fs_sell& out, int nsplit = -1, int sigma = 0, double pk_maxabs = 0.0);
double csr_maxabs(const fs_csr& A);
bool sell_pack_enabled();
// fp32 mirrors for the packed kernel with fp32 gathers: xf / x2f mirror x / x2 (both needed to select that kernel),
// yf (optional, any kernel) receives an fp32 copy of y
struct SellF32 { const float* xf = nullptr; const float* x2f = nullptr; float* yf = nullptr; };
int spmv_sell(const fs_sell& S, const double* x, double* y, const double* x2, double* dot_partials, const SellF32* f32v = nullptr);
int spmv_sell_grid(const fs_sell& S);   // grid (= dot partials per column) of spmv_sell2
void spmv_sell2(const fs_sell& S, const double* x, double* y, double* dot_partials, const int* done);
// any-matrix fallback: y = A [x; x2] (columns >= nsplit gather from x2; x2 null: y = A x)
void spmv_sub(const CsrView& A, const double* x, double* y, const double* x2, int nsplit, float* yf = nullptr);   // yf: fp32 mirror of y
struct AmgPartSpec;
Amg* amg_setup(fs_csr* fine, const AmgPartSpec* part = nullptr);   // part: row-block partitioned cycle (dist.cuh)
// x0_ready: the caller has written w D^-1 r into the buffer given by amg_presmooth_target (unfolded
// cycle only).  rz_part (optional): room for per-CTA partial sums of r.z; the return value is how
// many were written by the cycle's last kernel (0: the caller computes r.z itself).
// top_ev (optional, two events): run the cycle eagerly and bracket its largest kernel (the finest
// level's up-sweep) with them -- the sampled roofline timing of bench.py.
// r32 (optional): fp32 mirror of r, kept by the caller next to r -- the two finest-level kernels of the folded cycle then
// gather in fp32 (spmv_sell.cu, FMT 4).
int amg_apply(Amg* amg, const double* r, double* z, bool x0_ready = false, double* rz_part = nullptr,
              cudaEvent_t* top_ev = nullptr, const float* r32 = nullptr);
double amg_top_bytes(const Amg* amg);   // algorithmic bytes of one launch of that kernel (0: unfolded cycle)
double amg_cycle_bytes(const Amg* amg, bool mirrors);   // ... of one application of the whole folded cycle
void amg_presmooth_target(Amg* amg, double** x0, const double** dinv, double* omega);
int amg_levels(const Amg* amg, int* sizes, int cap);
bool cg_persistent_supported(const CsrView& A, size_t* smem_out);
void cg_persistent_launch(const CsrView& A, double* x, double* r, double* p, double* Ap, const double* dinv,
                          double* partA, double* partB, double* scal, int* flags, int maxit, double tol2);
// CG on device pointers; returns iterations, writes relres.
int cg_dev(fs_csr* a, const double* d_b, double* d_x, int nrhs, double rtol, int maxit, int precond,
           int project_mean, double* relres);
void cg_small_batch_dev(fs_csr* a, const double* d_b, double* d_x, int nrhs, int B, double rtol, int maxit, int precond,
                        int project_mean, double* ws, double* out, int* flags);
int cg_small_limit();
void spmv_dev(const CsrView& A, const double* d_x, double* d_y);
double resid_norm2_dev(fs_csr* a, const double* b, const double* x, double* work);   // |b - A x|^2
void spmv_best_dev(fs_csr* a, const double* x, double* y);                          // y = A x, fastest layout at hand
void cand_norms3_dev(fs_csr* a, const double* b, const double* y0, const double* y1, const double* y2 /* nullable */,
                     double* out3);   // |b - y0|^2, |b - (2 y0 - y1)|^2, |b - (3 y0 - 3 y1 + y2)|^2
void lin3_dev(int64_t n, double a, const double* x, double b, const double* y, double* out);   // out = a x + b y
double max_abs_dev(const double* d_x, int64_t n);
void free_locator(Locator* l);
}  // namespace fs
