// topology.cu -- device-side mesh ingest: structural CSR pattern of the P1
// stiffness matrix, element->nonzero scatter map, per-nonzero contribution
// lists and node->element incidence lists, all built with radix sorts + scans
// on the GPU.  Replaces the dense np.zeros((N,N)) scatter target of
// buildStiffnessMatrix (code/StokesColor.py:98-126): entry (a,b) exists iff a
// triangle holds both nodes.
#include <cub/cub.cuh>

#include "internal.cuh"

namespace fs {

__global__ void k_make_keys(const int* __restrict__ tris, const int* __restrict__ dof, int64_t T,
                            unsigned long long* __restrict__ keys, unsigned* __restrict__ payload) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= T) return;
  int v[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    int nd = tris[3 * e + i];
    v[i] = dof ? dof[nd] : nd;
  }
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      int64_t k = 9 * e + 3 * i + j;
      keys[k] = ((unsigned long long)(unsigned)v[i] << 32) | (unsigned)v[j];
      payload[k] = (unsigned)k;
    }
}

__global__ void k_head_flags(const unsigned long long* __restrict__ keys, int64_t m, int* __restrict__ flag) {
  int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= m) return;
  flag[k] = (k == 0 || keys[k] != keys[k - 1]) ? 1 : 0;
}

// pos = inclusive scan of flags; nnz id of sorted entry k is pos[k]-1
__global__ void k_fill_pattern(const unsigned long long* __restrict__ keys, const unsigned* __restrict__ payload,
                               const int* __restrict__ flag, const int* __restrict__ pos, int64_t m,
                               int* __restrict__ colidx, int* __restrict__ rowof, int* __restrict__ seg_start,
                               int* __restrict__ scatter) {
  int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= m) return;
  int z = pos[k] - 1;
  scatter[payload[k]] = z;
  if (flag[k]) {
    colidx[z] = (int)(keys[k] & 0xffffffffu);
    rowof[z] = (int)(keys[k] >> 32);
    seg_start[z] = (int)k;
  }
}

// rowptr from the (ascending) row id of every nonzero; handles empty rows.
__global__ void k_rowptr(const int* __restrict__ rowof, int64_t nnz, int64_t n, int* __restrict__ rowptr) {
  int64_t z = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (z > nnz) return;
  int64_t lo = (z == 0) ? -1 : rowof[z - 1];
  int64_t hi = (z == nnz) ? n : rowof[z];
  for (int64_t r = lo + 1; r <= hi; ++r) rowptr[r] = (int)z;
}

static int bits_for(uint64_t v) {
  int b = 1;
  while (b < 64 && (v >> b)) ++b;
  return b;
}

void build_pattern(const int* d_tris, int64_t T, int64_t n, const int* d_dof, Pattern& out) {
  cudaStream_t st = stream();
  const int64_t m = 9 * T;
  FS_REQUIRE(m < (int64_t)1 << 31, "mesh too large for 32-bit contribution ids");
  out.n = n;
  out.T = T;
  DBuf<unsigned long long> keys(m), keys_alt(m);
  DBuf<unsigned> pay(m);
  out.contrib.alloc(m);
  const int B = 256;
  k_make_keys<<<div_up(T, B), B, 0, st>>>(d_tris, d_dof, T, keys.p, pay.p);
  FS_LAUNCH_CHECK();
  // stable LSD radix sort: equal keys keep ascending element-entry order
  cub::DoubleBuffer<unsigned long long> kb(keys.p, keys_alt.p);
  DBuf<unsigned> pay_alt(m);
  cub::DoubleBuffer<unsigned> vb(pay.p, pay_alt.p);
  int nbits = bits_for((uint64_t)(n > 1 ? n - 1 : 1));
  size_t tmp_bytes = 0;
  FS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kb, vb, (int)m, 0, 32 + nbits, st));
  DBuf<char> tmp(tmp_bytes);
  // low word: bits [0,nbits); high word: bits [32, 32+nbits).  Sorting the full
  // [0, 32+nbits) range is simplest and still skips the top passes.
  FS_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, kb, vb, (int)m, 0, 32 + nbits, st));
  count_launch(8);
  const unsigned long long* skeys = kb.Current();
  const unsigned* spay = vb.Current();
  DBuf<int> flag(m), pos(m);
  k_head_flags<<<div_up(m, B), B, 0, st>>>(skeys, m, flag.p);
  FS_LAUNCH_CHECK();
  size_t scan_bytes = 0;
  FS_CUDA(cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, flag.p, pos.p, (int)m, st));
  DBuf<char> tmp2(scan_bytes);
  FS_CUDA(cub::DeviceScan::InclusiveSum(tmp2.p, scan_bytes, flag.p, pos.p, (int)m, st));
  count_launch(2);
  int nnz = 0;
  if (m > 0) {
    FS_CUDA(cudaMemcpyAsync(&nnz, pos.p + (m - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
    FS_CUDA(cudaStreamSynchronize(st));
  }
  out.nnz = nnz;
  out.colidx.alloc(nnz);
  out.seg_start.alloc((size_t)nnz + 1);
  out.scatter.alloc(m);
  out.rowptr.alloc(n + 1);
  DBuf<int> rowof(nnz);
  k_fill_pattern<<<div_up(m, B), B, 0, st>>>(skeys, spay, flag.p, pos.p, m, out.colidx.p, rowof.p,
                                             out.seg_start.p, out.scatter.p);
  FS_LAUNCH_CHECK();
  int mi = (int)m;
  FS_CUDA(cudaMemcpyAsync(out.seg_start.p + nnz, &mi, sizeof(int), cudaMemcpyHostToDevice, st));
  k_rowptr<<<div_up((int64_t)nnz + 1, B), B, 0, st>>>(rowof.p, nnz, n, out.rowptr.p);
  FS_LAUNCH_CHECK();
  FS_CUDA(cudaMemcpyAsync(out.contrib.p, spay, m * sizeof(unsigned), cudaMemcpyDeviceToDevice, st));
  FS_CUDA(cudaStreamSynchronize(st));
}

// vals[z] = sum of the element entries in z's contribution list, in list order
// (ascending element id) -> same floating-point sum as the reference's
// sequential "AMatrix[tri[i], tri[j]] += integral_val".
__global__ void k_assemble(const int* __restrict__ seg_start, const unsigned* __restrict__ contrib,
                           const double* __restrict__ ke, int64_t nnz, double* __restrict__ vals) {
  int64_t z = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (z >= nnz) return;
  int a = seg_start[z], b = seg_start[z + 1];
  double s = 0.0;
  for (int k = a; k < b; ++k) s += ke[contrib[k]];
  vals[z] = s;
}

void assemble_on_pattern(const Pattern& pat, const double* d_ke, double* d_vals) {
  if (pat.nnz == 0) return;
  k_assemble<<<div_up(pat.nnz, 256), 256, 0, stream()>>>(pat.seg_start.p, pat.contrib.p, d_ke, pat.nnz, d_vals);
  FS_LAUNCH_CHECK();
}

// ---- node -> incident element corners --------------------------------------
__global__ void k_inc_keys(const int* __restrict__ tris, int64_t m, unsigned* __restrict__ keys,
                           unsigned* __restrict__ payload) {
  int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= m) return;
  keys[k] = (unsigned)tris[k];
  payload[k] = (unsigned)k;
}

__global__ void k_inc_ptr(const unsigned* __restrict__ skeys, int64_t m, int64_t n, int* __restrict__ ptr) {
  int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k > m) return;
  int64_t lo = (k == 0) ? -1 : (int64_t)skeys[k - 1];
  int64_t hi = (k == m) ? n : (int64_t)skeys[k];
  for (int64_t r = lo + 1; r <= hi; ++r) ptr[r] = (int)k;
}

static void build_incidence(fs_mesh* mesh) {
  cudaStream_t st = stream();
  const int64_t m = 3 * mesh->T;
  DBuf<unsigned> keys(m), keys_alt(m), pay(m), pay_alt(m);
  const int B = 256;
  k_inc_keys<<<div_up(m, B), B, 0, st>>>(mesh->tris.p, m, keys.p, pay.p);
  FS_LAUNCH_CHECK();
  cub::DoubleBuffer<unsigned> kb(keys.p, keys_alt.p), vb(pay.p, pay_alt.p);
  int nbits = bits_for((uint64_t)(mesh->N > 1 ? mesh->N - 1 : 1));
  size_t tmp_bytes = 0;
  FS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kb, vb, (int)m, 0, nbits, st));
  DBuf<char> tmp(tmp_bytes);
  FS_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, kb, vb, (int)m, 0, nbits, st));
  count_launch(4);
  mesh->inc_ptr.alloc(mesh->N + 1);
  mesh->inc.alloc(m);
  k_inc_ptr<<<div_up(m + 1, B), B, 0, st>>>(kb.Current(), m, mesh->N, mesh->inc_ptr.p);
  FS_LAUNCH_CHECK();
  FS_CUDA(cudaMemcpyAsync(mesh->inc.p, vb.Current(), m * sizeof(unsigned), cudaMemcpyDeviceToDevice, st));
  FS_CUDA(cudaStreamSynchronize(st));
}

__global__ void k_check_tris(const int* __restrict__ tris, int64_t m, int n, int* __restrict__ bad) {
  int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= m) return;
  int v = tris[k];
  if (v < 0 || v >= n) atomicExch(bad, 1);
}

}  // namespace fs

using namespace fs;

fs_mesh::~fs_mesh() {
  if (loc) fs::free_locator(loc);
}

extern "C" {

int fs_mesh_create(const double* coords, int64_t N, const int32_t* tris, int64_t T, const int32_t* markers,
                   fs_mesh** out) {
  FS_API_BEGIN
  FS_REQUIRE(out, "out is NULL");
  *out = nullptr;
  FS_REQUIRE(coords && tris, "coords/tris are NULL");
  FS_REQUIRE(N > 0 && T > 0, "empty mesh");
  FS_REQUIRE(N < ((int64_t)1 << 31) - 2, "too many nodes for 32-bit indices");
  std::unique_ptr<fs_mesh> m(new fs_mesh());
  m->N = N;
  m->T = T;
  m->coords.alloc(2 * N);
  m->coords.upload(coords, 2 * N);
  m->tris.alloc(3 * T);
  m->tris.upload(tris, 3 * T);
  m->markers.alloc(N);
  if (markers) m->markers.upload(markers, N); else m->markers.zero();
  DBuf<int> bad(1);
  bad.zero();
  k_check_tris<<<div_up(3 * T, 256), 256, 0, stream()>>>(m->tris.p, 3 * T, (int)N, bad.p);
  FS_LAUNCH_CHECK();
  if (bad.to_host()[0]) throw Error(FS_ERR_ARG, "triangle refers to a node id outside [0, N)");
  build_pattern(m->tris.p, T, N, nullptr, m->pat);
  build_incidence(m.get());
  *out = m.release();
  FS_API_END
}

int fs_mesh_destroy(fs_mesh* m) {
  FS_API_BEGIN
  if (m) { cudaStreamSynchronize(stream()); delete m; }
  FS_API_END
}

int fs_mesh_sizes(const fs_mesh* m, int64_t* n, int64_t* t, int64_t* nnz) {
  FS_API_BEGIN
  FS_REQUIRE(m, "mesh is NULL");
  if (n) *n = m->N;
  if (t) *t = m->T;
  if (nnz) *nnz = m->pat.nnz;
  FS_API_END
}

int fs_csr_pattern(const fs_mesh* m, int32_t* rowptr, int32_t* colidx) {
  FS_API_BEGIN
  FS_REQUIRE(m, "mesh is NULL");
  if (rowptr) FS_CUDA(cudaMemcpyAsync(rowptr, m->pat.rowptr.p, (m->N + 1) * sizeof(int), cudaMemcpyDefault, stream()));
  if (colidx && m->pat.nnz)
    FS_CUDA(cudaMemcpyAsync(colidx, m->pat.colidx.p, m->pat.nnz * sizeof(int), cudaMemcpyDefault, stream()));
  fs::sync();
  FS_API_END
}

int fs_scatter_map(const fs_mesh* m, int32_t* scatter) {
  FS_API_BEGIN
  FS_REQUIRE(m && scatter, "NULL argument");
  FS_CUDA(cudaMemcpyAsync(scatter, m->pat.scatter.p, 9 * m->T * sizeof(int), cudaMemcpyDefault, stream()));
  fs::sync();
  FS_API_END
}

}  // extern "C"
