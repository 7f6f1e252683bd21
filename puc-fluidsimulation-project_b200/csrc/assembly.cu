// assembly.cu -- P1 element kernels: stiffness, lumped mass, FEM load vector,
// lumped nodal divergence / gradient, boundary-condition imposition.
// Compiled with -fmad=false: every formula keeps the reference's operation
// order so K, M and div are bit-identical to the reference's Python loops.
#include "internal.cuh"

namespace fs {

constexpr int kWarpsPerBlock = 4;
constexpr int kEB = 32;  // elements per warp batch

// Stage one warp batch (32 elements): triangle corner ids with coalesced loads,
// then the 96 vertex coordinates gathered cooperatively into shared memory.
template <class S>
struct ElemStage {
  int tri[kEB * 3];
  S x[kEB * 3];
  S y[kEB * 3];
  double out[kEB * 9];
};

template <class S>
__device__ __forceinline__ void stage_batch(ElemStage<S>& s, const double2* __restrict__ coords,
                                            const int* __restrict__ tris, int64_t e0, int64_t T, int lane) {
  const int64_t base = 3 * e0;
  const int64_t lim = 3 * T;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    int64_t k = base + r * 32 + lane;
    int v = (k < lim) ? tris[k] : 0;
    s.tri[r * 32 + lane] = v;
    double2 c = (k < lim) ? __ldg(&coords[v]) : make_double2(0.0, 0.0);
    s.x[r * 32 + lane] = (S)c.x;   // (float) cast == np.float32 storage, code/poisson.py:40
    s.y[r * 32 + lane] = (S)c.y;
  }
  __syncwarp();
}

// MODE 0: buildStiffnessMatrix (code/StokesColor.py:111-124): 2|ADet|, skip |ADet|<1e-14
// MODE 1: buildFemSystem      (code/poisson.py:111-125): ADet expanded, signed 2*ADet, skip ==0
template <class S, int MODE>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
k_element_stiffness(const double2* __restrict__ coords, const int* __restrict__ tris, int64_t T,
                    double* __restrict__ ke, const double* __restrict__ g_centroid, double g_const,
                    double* __restrict__ src /* MODE 1: per-element load contribution */) {
  __shared__ ElemStage<S> stage[kWarpsPerBlock];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ElemStage<S>& s = stage[warp];
  const int64_t nb = (T + kEB - 1) / kEB;
  for (int64_t b = blockIdx.x * (int64_t)kWarpsPerBlock + warp; b < nb; b += (int64_t)gridDim.x * kWarpsPerBlock) {
    const int64_t e0 = b * kEB;
    stage_batch<S>(s, coords, tris, e0, T, lane);
    const int64_t e = e0 + lane;
    const S x1 = s.x[3 * lane], x2 = s.x[3 * lane + 1], x3 = s.x[3 * lane + 2];
    const S y1 = s.y[3 * lane], y2 = s.y[3 * lane + 1], y3 = s.y[3 * lane + 2];
    S adet;
    bool skip;
    if (MODE == 0) {
      adet = x1 * (y2 - y3) + x2 * (y3 - y1) + x3 * (y1 - y2);
      skip = fabs((double)adet) < 1e-14;
    } else {
      adet = x1 * y2 - x1 * y3 - x2 * y1 + x2 * y3 + x3 * y1 - x3 * y2;
      skip = (adet == (S)0);
    }
    const S yd[3] = {y2 - y3, y3 - y1, y1 - y2};
    const S xd[3] = {x3 - x2, x1 - x3, x2 - x1};
    const S den = (MODE == 0) ? (S)2 * (S)fabs((double)adet) : (S)2.0 * adet;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        S num = yd[i] * yd[j] + xd[i] * xd[j];
        s.out[lane * 9 + i * 3 + j] = skip ? 0.0 : (double)(num / den);
      }
    if (MODE == 1 && src && e < T) {
      S area = (S)0.5 * adet;
      S g = g_centroid ? (S)g_centroid[e] : (S)g_const;
      src[e] = skip ? 0.0 : (double)(g * (area / (S)3));
    }
    __syncwarp();
    // coalesced write of the batch's 288 entries
    const int64_t obase = 9 * e0, olim = 9 * T;
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      int64_t k = obase + r * 32 + lane;
      if (k < olim) ke[k] = s.out[r * 32 + lane];
    }
    __syncwarp();
  }
}

// centroid in the coordinate dtype: ((x1+x2+x3)/3), code/poisson.py:133
template <class S>
__global__ void k_centroids(const double2* __restrict__ coords, const int* __restrict__ tris, int64_t T,
                            double* __restrict__ cx, double* __restrict__ cy) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= T) return;
  double2 a = __ldg(&coords[tris[3 * e]]), b = __ldg(&coords[tris[3 * e + 1]]), c = __ldg(&coords[tris[3 * e + 2]]);
  cx[e] = (double)(((S)a.x + (S)b.x + (S)c.x) / (S)3);
  cy[e] = (double)(((S)a.y + (S)b.y + (S)c.y) / (S)3);
}

struct Geo {
  double det, yd[3], xd[3];
};

__device__ __forceinline__ Geo elem_geo(const double2* __restrict__ coords, int a, int b, int c) {
  double2 p1 = __ldg(&coords[a]), p2 = __ldg(&coords[b]), p3 = __ldg(&coords[c]);
  Geo g;
  g.det = p1.x * (p2.y - p3.y) + p2.x * (p3.y - p1.y) + p3.x * (p1.y - p2.y);
  g.yd[0] = p2.y - p3.y; g.yd[1] = p3.y - p1.y; g.yd[2] = p1.y - p2.y;
  g.xd[0] = p3.x - p2.x; g.xd[1] = p1.x - p3.x; g.xd[2] = p2.x - p1.x;
  return g;
}

// area/3 per element: with skip (div/grad rule) and without (lumped-mass rule)
__global__ void k_elem_thirds(const double2* __restrict__ coords, const int* __restrict__ tris, int64_t T,
                              double* __restrict__ third_all, double* __restrict__ third_skip) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= T) return;
  Geo g = elem_geo(coords, tris[3 * e], tris[3 * e + 1], tris[3 * e + 2]);
  double area = 0.5 * fabs(g.det);
  double third = area / 3.0;
  third_all[e] = third;
  third_skip[e] = (fabs(g.det) < 1e-14) ? 0.0 : third;
}

// out[n] = sum over incident element corners (ascending element id) of elem[e]
__global__ void k_node_sum(const int* __restrict__ inc_ptr, const unsigned* __restrict__ inc,
                           const double* __restrict__ elem, int64_t N, double* __restrict__ out, double scale) {
  int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (n >= N) return;
  double s = 0.0;
  for (int k = inc_ptr[n]; k < inc_ptr[n + 1]; ++k) s += elem[inc[k] / 3u];
  out[n] = scale * s;
}

// calculate_divergence element part, code/StokesColor.py:145-160
// (blockIdx.y = configuration of a batched sweep: own fields on the shared mesh, strides bs_* between configurations)
__global__ void k_div_elem(const double2* __restrict__ coords, const int* __restrict__ tris,
                           const double2* __restrict__ u, int64_t T, double* __restrict__ lump, int64_t bs_n = 0) {
  u += blockIdx.y * bs_n; lump += blockIdx.y * T;
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= T) return;
  int a = tris[3 * e], b = tris[3 * e + 1], c = tris[3 * e + 2];
  Geo g = elem_geo(coords, a, b, c);
  if (fabs(g.det) < 1e-14) { lump[e] = 0.0; return; }
  double inv2A = 1.0 / g.det;
  double area = 0.5 * fabs(g.det);
  double2 u1 = __ldg(&u[a]), u2 = __ldg(&u[b]), u3 = __ldg(&u[c]);
  double dux = (u1.x * g.yd[0] + u2.x * g.yd[1] + u3.x * g.yd[2]) * inv2A;
  double duy = (u1.y * g.xd[0] + u2.y * g.xd[1] + u3.y * g.xd[2]) * inv2A;
  double div_tri = dux + duy;
  lump[e] = div_tri * (area / 3.0);
}

__global__ void k_div_node(const int* __restrict__ inc_ptr, const unsigned* __restrict__ inc,
                           const double* __restrict__ lump, const double* __restrict__ area_sum, int64_t N,
                           double* __restrict__ div, double* __restrict__ div_sum, int64_t bs_n = 0, int64_t bs_t = 0) {
  lump += blockIdx.y * bs_t;
  if (div) div += blockIdx.y * bs_n;
  int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (n >= N) return;
  double s = 0.0;
  for (int k = inc_ptr[n]; k < inc_ptr[n + 1]; ++k) s += lump[inc[k] / 3u];
  if (div_sum) div_sum[n] = s;
  if (div) div[n] = s / (area_sum[n] + 1e-12);
}

// calculate_gradiant element part, code/StokesColor.py:235-253
__global__ void k_grad_elem(const double2* __restrict__ coords, const int* __restrict__ tris,
                            const double* __restrict__ p, int64_t T, double* __restrict__ lx,
                            double* __restrict__ ly, int64_t bs_n = 0) {
  p += blockIdx.y * bs_n; lx += blockIdx.y * T; ly += blockIdx.y * T;
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= T) return;
  int a = tris[3 * e], b = tris[3 * e + 1], c = tris[3 * e + 2];
  Geo g = elem_geo(coords, a, b, c);
  if (fabs(g.det) < 1e-14) { lx[e] = 0.0; ly[e] = 0.0; return; }
  double inv2A = 1.0 / g.det;
  double area = 0.5 * fabs(g.det);
  double p1 = __ldg(&p[a]), p2 = __ldg(&p[b]), p3 = __ldg(&p[c]);
  double gx = (g.yd[0] * inv2A) * p1 + (g.yd[1] * inv2A) * p2 + (g.yd[2] * inv2A) * p3;
  double gy = (g.xd[0] * inv2A) * p1 + (g.xd[1] * inv2A) * p2 + (g.xd[2] * inv2A) * p3;
  lx[e] = gx * (area / 3.0);
  ly[e] = gy * (area / 3.0);
}

// MODE 0: write gx, gy.   MODE 1: uo = ui - DT*grad (code/StokesColor.py:561-562)
// MODE 2: uo[interior] -= DT*grad[interior] (:572-573), uo == ui
template <int MODE>
__global__ void k_grad_node(const int* __restrict__ inc_ptr, const unsigned* __restrict__ inc,
                            const double* __restrict__ lx, const double* __restrict__ ly,
                            const double* __restrict__ area_sum, int64_t N, double* __restrict__ gx,
                            double* __restrict__ gy, const double2* __restrict__ ui, double2* __restrict__ uo,
                            double DT, const unsigned char* __restrict__ is_interior, int64_t bs_n = 0, int64_t bs_t = 0) {
  lx += blockIdx.y * bs_t; ly += blockIdx.y * bs_t;
  if (MODE != 0) { ui += blockIdx.y * bs_n; uo += blockIdx.y * bs_n; }
  int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (n >= N) return;
  if (MODE == 2 && !is_interior[n]) return;
  double sx = 0.0, sy = 0.0;
  for (int k = inc_ptr[n]; k < inc_ptr[n + 1]; ++k) {
    unsigned e = inc[k] / 3u;
    sx += lx[e];
    sy += ly[e];
  }
  double den = area_sum[n] + 1e-12;
  double vx = sx / den, vy = sy / den;
  if (MODE == 0) { gx[n] = vx; gy[n] = vy; }
  else {
    double2 v = ui[n];
    v.x = v.x - DT * vx;
    v.y = v.y - DT * vy;
    uo[n] = v;
  }
}

void ensure_geom(fs_mesh* m) {
  if (m->geom_ready) return;
  cudaStream_t st = stream();
  m->elem_a.alloc(m->T); m->elem_b.alloc(m->T); m->elem_c.alloc(m->T);
  m->area_sum.alloc(m->N); m->mass.alloc(m->N);
  const int B = 256;
  k_elem_thirds<<<div_up(m->T, B), B, 0, st>>>((const double2*)m->coords.p, m->tris.p, m->T, m->elem_a.p, m->elem_b.p);
  FS_LAUNCH_CHECK();
  k_node_sum<<<div_up(m->N, B), B, 0, st>>>(m->inc_ptr.p, m->inc.p, m->elem_a.p, m->N, m->mass.p, 1.0);
  FS_LAUNCH_CHECK();
  k_node_sum<<<div_up(m->N, B), B, 0, st>>>(m->inc_ptr.p, m->inc.p, m->elem_b.p, m->N, m->area_sum.p, 1.0);
  FS_LAUNCH_CHECK();
  m->geom_ready = true;
}

void divergence_dev(fs_mesh* m, const double* d_u, double* d_div, double* d_div_sum) {
  ensure_geom(m);
  cudaStream_t st = stream();
  const int B = 256;
  k_div_elem<<<div_up(m->T, B), B, 0, st>>>((const double2*)m->coords.p, m->tris.p, (const double2*)d_u, m->T, m->elem_a.p);
  FS_LAUNCH_CHECK();
  k_div_node<<<div_up(m->n_eval(), B), B, 0, st>>>(m->inc_ptr.p, m->inc.p, m->elem_a.p, m->area_sum.p, m->n_eval(), d_div, d_div_sum);
  FS_LAUNCH_CHECK();
}

static void grad_elem(fs_mesh* m, const double* d_p) {
  ensure_geom(m);
  k_grad_elem<<<div_up(m->T, 256), 256, 0, stream()>>>((const double2*)m->coords.p, m->tris.p, d_p, m->T, m->elem_a.p, m->elem_c.p);
  FS_LAUNCH_CHECK();
}

void gradient_dev(fs_mesh* m, const double* d_p, double* d_gx, double* d_gy) {
  grad_elem(m, d_p);
  k_grad_node<0><<<div_up(m->n_eval(), 256), 256, 0, stream()>>>(m->inc_ptr.p, m->inc.p, m->elem_a.p, m->elem_c.p, m->area_sum.p,
                                                           m->n_eval(), d_gx, d_gy, nullptr, nullptr, 0.0, nullptr);
  FS_LAUNCH_CHECK();
}

void grad_update_dev(fs_mesh* m, const double* d_p, const double* d_ui, double* d_uo, double DT,
                     const unsigned char* d_interior_flag) {
  grad_elem(m, d_p);
  if (d_interior_flag)
    k_grad_node<2><<<div_up(m->n_eval(), 256), 256, 0, stream()>>>(m->inc_ptr.p, m->inc.p, m->elem_a.p, m->elem_c.p, m->area_sum.p,
                                                             m->n_eval(), nullptr, nullptr, (const double2*)d_ui, (double2*)d_uo, DT,
                                                             d_interior_flag);
  else
    k_grad_node<1><<<div_up(m->n_eval(), 256), 256, 0, stream()>>>(m->inc_ptr.p, m->inc.p, m->elem_a.p, m->elem_c.p, m->area_sum.p,
                                                             m->n_eval(), nullptr, nullptr, (const double2*)d_ui, (double2*)d_uo, DT,
                                                             nullptr);
  FS_LAUNCH_CHECK();
}

// ---- boundary conditions -----------------------------------------------------
__global__ void k_inner_trig(const double2* __restrict__ coords, const int* __restrict__ inner, int64_t ni,
                             double* __restrict__ s, double* __restrict__ c, double* __restrict__ s2) {
  int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= ni) return;
  double2 p = coords[inner[k]];
  double th = atan2(p.y - 0.5, p.x - 0.5);   // code/StokesColor.py:413-415
  s[k] = sin(th);
  c[k] = cos(th);
  s2[k] = sin(2 * th);
}

__global__ void k_per_bcu_par(const int* __restrict__ pairs, int64_t np, double2* __restrict__ u, int64_t bs_n = 0) {
  u += blockIdx.y * bs_n;
  int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= np) return;
  u[pairs[2 * k + 1]] = u[pairs[2 * k]];
}

__global__ void k_per_bcu_seq(const int* __restrict__ pairs, int64_t np, double2* __restrict__ u, int64_t bs_n = 0) {
  u += blockIdx.y * bs_n;
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int64_t k = 0; k < np; ++k) u[pairs[2 * k + 1]] = u[pairs[2 * k]];
}

__global__ void k_scalar_per_seq(const int* __restrict__ pairs, int64_t np, double* __restrict__ u) {
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int64_t k = 0; k < np; ++k) u[pairs[2 * k + 1]] = u[pairs[2 * k]];
}

// makeDirBCU, code/StokesColor.py:405-427
__global__ void k_dir_bcu(const int* __restrict__ wall, int64_t nw, const int* __restrict__ inner, int64_t ni,
                          const double* __restrict__ s, const double* __restrict__ c, const double* __restrict__ s2,
                          double B1, double B2, double2* __restrict__ u, const double* __restrict__ b12 = nullptr, int64_t bs_n = 0) {
  u += blockIdx.y * bs_n;
  if (b12) { B1 = b12[2 * blockIdx.y]; B2 = b12[2 * blockIdx.y + 1]; }     // batched sweep: (B1, B2) per configuration
  int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k < nw) u[wall[k]] = make_double2(0.0, 0.0);
  else if (k < nw + ni) {
    int64_t q = k - nw;
    double vt = B1 * s[q] + B2 * s2[q];
    u[inner[q]] = make_double2(vt * (-s[q]), vt * c[q]);
  }
}

__global__ void k_scalar_dir(const int* __restrict__ wall, int64_t nw, const int* __restrict__ inner, int64_t ni,
                             double wv, double iv, double* __restrict__ u) {
  int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k < nw) u[wall[k]] = wv;
  else if (k < nw + ni) u[inner[k - nw]] = iv;
}

static bool pairs_independent(const std::vector<int>& pr) {
  // parallel-safe iff no node is written twice and no written node is also read
  std::vector<int> slaves, masters;
  for (size_t k = 0; k + 1 < pr.size(); k += 2) { masters.push_back(pr[k]); slaves.push_back(pr[k + 1]); }
  std::sort(slaves.begin(), slaves.end());
  if (std::adjacent_find(slaves.begin(), slaves.end()) != slaves.end()) return false;
  for (int mm : masters)
    if (std::binary_search(slaves.begin(), slaves.end(), mm)) return false;
  return true;
}

void per_bcu_dev(fs_mesh* m, double* d_u) {
  FS_REQUIRE(m->bc_ready, "fs_bc_set has not been called");
  if (m->n_pairs == 0) return;
  if (pairs_independent(m->pairs_host))
    k_per_bcu_par<<<div_up(m->n_pairs, 128), 128, 0, stream()>>>(m->pairs.p, m->n_pairs, (double2*)d_u);
  else
    k_per_bcu_seq<<<1, 32, 0, stream()>>>(m->pairs.p, m->n_pairs, (double2*)d_u);
  FS_LAUNCH_CHECK();
}

void dir_bcu_dev(fs_mesh* m, double* d_u, double B1, double B2) {
  FS_REQUIRE(m->bc_ready, "fs_bc_set has not been called");
  int64_t tot = m->n_wall + m->n_inner;
  if (tot == 0) return;
  k_dir_bcu<<<div_up(tot, 128), 128, 0, stream()>>>(m->wall.p, m->n_wall, m->inner.p, m->n_inner, m->inner_sin.p,
                                                     m->inner_cos.p, m->inner_sin2.p, B1, B2, (double2*)d_u);
  FS_LAUNCH_CHECK();
}

// ---- physics variants of the reference's draft scripts ----------------------------------
// rotating inner cylinder: u = omega x r on the inner boundary, 0 on the walls (scripts/stokes_report.py:1155-1171)
__global__ void k_rot_bcu(const double2* __restrict__ coords, const int* __restrict__ wall, int64_t nw, const int* __restrict__ inner,
                          int64_t ni, double omega, double cx, double cy, double2* __restrict__ u) {
  int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k < nw) u[wall[k]] = make_double2(0.0, 0.0);
  else if (k < nw + ni) {
    const int idx = inner[k - nw];
    const double2 p = coords[idx];
    const double rx = p.x - cx, ry = p.y - cy;
    u[idx] = make_double2(-ry * omega, rx * omega);
  }
}
void rot_bcu_dev(fs_mesh* m, double* d_u, double omega, double cx, double cy) {
  FS_REQUIRE(m->bc_ready, "fs_bc_set has not been called");
  const int64_t tot = m->n_wall + m->n_inner;
  if (tot == 0) return;
  k_rot_bcu<<<div_up(tot, 128), 128, 0, stream()>>>((const double2*)m->coords.p, m->wall.p, m->n_wall, m->inner.p, m->n_inner, omega, cx, cy,
                                                     (double2*)d_u);
  FS_LAUNCH_CHECK();
}

// consistent mass and convection element matrices of build_mass_and_convection, code/StokesColor.py:286-312
__global__ void k_elem_mass_conv(const double2* __restrict__ coords, const int* __restrict__ tris, const double2* __restrict__ u,
                                 int64_t T, double* __restrict__ keM, double* __restrict__ keC) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= T) return;
  const int a = tris[3 * e], b = tris[3 * e + 1], c = tris[3 * e + 2];
  const double2 p1 = coords[a], p2 = coords[b], p3 = coords[c];
  const double det = p1.x * (p2.y - p3.y) + p2.x * (p3.y - p1.y) + p3.x * (p1.y - p2.y);
  const bool skip = fabs(det) < 1e-14;
  const double area = 0.5 * fabs(det);
  const double2 ua = u[a], ub = u[b], uc = u[c];
  const double ucx = ((ua.x + ub.x) + uc.x) / 3.0, ucy = ((ua.y + ub.y) + uc.y) / 3.0;      // u[idx].mean(axis=0)
  const double den = 2 * fabs(det);
  const double gx[3] = {(p2.y - p3.y) / den, (p3.y - p1.y) / den, (p1.y - p2.y) / den};
  const double gy[3] = {(p3.x - p2.x) / den, (p1.x - p3.x) / den, (p2.x - p1.x) / den};
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      keM[9 * e + 3 * i + j] = skip ? 0.0 : (area / 12.0) * (i != j ? 1.0 : 2.0);
      keC[9 * e + 3 * i + j] = skip ? 0.0 : (area / 3) * (ucx * gx[j] + ucy * gy[j]);
    }
}

// c <- clip(c + DT D (K c), 0, 1): the explicit dye "diffusion" of scripts/good_visualization2.py:704-715 (sign as in the script)
__global__ void k_dye_diffuse(CsrView K, const double* __restrict__ c, double s, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K.n) return;
  double lap = 0.0;
  for (int k = K.rowptr[i]; k < K.rowptr[i + 1]; ++k) lap += K.vals[k] * c[K.colidx[k]];
  const double v = c[i] + s * lap;
  out[i] = fmin(fmax(v, 0.0), 1.0);
}

// ---- the same operators for B configurations at once (shared mesh, fields strided by N / T) ----
void divergence_batch_dev(fs_mesh* m, int B, const double* d_u, double* d_div, double* d_lump) {
  ensure_geom(m);
  cudaStream_t st = stream();
  k_div_elem<<<dim3(div_up(m->T, 256), B), 256, 0, st>>>((const double2*)m->coords.p, m->tris.p, (const double2*)d_u, m->T, d_lump, m->N);
  FS_LAUNCH_CHECK();
  k_div_node<<<dim3(div_up(m->n_eval(), 256), B), 256, 0, st>>>(m->inc_ptr.p, m->inc.p, d_lump, m->area_sum.p, m->n_eval(), d_div, nullptr, m->N, m->T);
  FS_LAUNCH_CHECK();
}
void grad_update_batch_dev(fs_mesh* m, int B, const double* d_p, const double* d_ui, double* d_uo, double DT,
                           const unsigned char* d_interior_flag, double* d_lx, double* d_ly) {
  ensure_geom(m);
  cudaStream_t st = stream();
  k_grad_elem<<<dim3(div_up(m->T, 256), B), 256, 0, st>>>((const double2*)m->coords.p, m->tris.p, d_p, m->T, d_lx, d_ly, m->N);
  FS_LAUNCH_CHECK();
  const dim3 g(div_up(m->n_eval(), 256), B);
  if (d_interior_flag)
    k_grad_node<2><<<g, 256, 0, st>>>(m->inc_ptr.p, m->inc.p, d_lx, d_ly, m->area_sum.p, m->n_eval(), nullptr, nullptr, (const double2*)d_ui,
                                      (double2*)d_uo, DT, d_interior_flag, m->N, m->T);
  else
    k_grad_node<1><<<g, 256, 0, st>>>(m->inc_ptr.p, m->inc.p, d_lx, d_ly, m->area_sum.p, m->n_eval(), nullptr, nullptr, (const double2*)d_ui,
                                      (double2*)d_uo, DT, nullptr, m->N, m->T);
  FS_LAUNCH_CHECK();
}
void bcu_batch_dev(fs_mesh* m, int B, double* d_u, const double* d_b12) {
  FS_REQUIRE(m->bc_ready, "fs_bc_set has not been called");
  cudaStream_t st = stream();
  if (m->n_pairs) {
    if (pairs_independent(m->pairs_host)) k_per_bcu_par<<<dim3(div_up(m->n_pairs, 128), B), 128, 0, st>>>(m->pairs.p, m->n_pairs, (double2*)d_u, m->N);
    else k_per_bcu_seq<<<dim3(1, B), 32, 0, st>>>(m->pairs.p, m->n_pairs, (double2*)d_u, m->N);
    FS_LAUNCH_CHECK();
  }
  const int64_t tot = m->n_wall + m->n_inner;
  if (tot) {
    k_dir_bcu<<<dim3(div_up(tot, 128), B), 128, 0, st>>>(m->wall.p, m->n_wall, m->inner.p, m->n_inner, m->inner_sin.p, m->inner_cos.p,
                                                        m->inner_sin2.p, 0.0, 0.0, (double2*)d_u, d_b12, m->N);
    FS_LAUNCH_CHECK();
  }
}

}  // namespace fs

using namespace fs;

extern "C" {

int fs_assemble_stiffness(fs_mesh* m, double* vals) {
  FS_API_BEGIN
  FS_REQUIRE(m && vals, "NULL argument");
  if (!m->ke.n) m->ke.alloc(9 * m->T);
  int64_t nb = (m->T + kEB - 1) / kEB;
  int grid = (int)std::min<int64_t>(div_up(nb, kWarpsPerBlock), (int64_t)sm_count() * 16);
  k_element_stiffness<double, 0><<<grid, kWarpsPerBlock * 32, 0, stream()>>>(
      (const double2*)m->coords.p, m->tris.p, m->T, m->ke.p, nullptr, 0.0, nullptr);
  FS_LAUNCH_CHECK();
  Out<double> o(vals, m->pat.nnz);
  assemble_on_pattern(m->pat, m->ke.p, o.d);
  o.commit();
  fs::sync();
  FS_API_END
}

int fs_assemble_mass_convection(fs_mesh* m, const double* u, double* m_vals, double* c_vals) {
  FS_API_BEGIN
  FS_REQUIRE(m && u && m_vals && c_vals, "NULL argument");
  In<double> iu(u, 2 * m->N);
  DBuf<double> keM((size_t)9 * m->T), keC((size_t)9 * m->T);
  k_elem_mass_conv<<<div_up(m->T, 256), 256, 0, stream()>>>((const double2*)m->coords.p, m->tris.p, (const double2*)iu.d, m->T, keM.p, keC.p);
  FS_LAUNCH_CHECK();
  Out<double> om(m_vals, m->pat.nnz), oc(c_vals, m->pat.nnz);
  assemble_on_pattern(m->pat, keM.p, om.d);
  assemble_on_pattern(m->pat, keC.p, oc.d);
  om.commit(); oc.commit();
  fs::sync();
  FS_API_END
}

int fs_make_rot_bcu(fs_mesh* m, double* u, double omega, double cx, double cy) {
  FS_API_BEGIN
  FS_REQUIRE(m && u, "NULL argument");
  Out<double> o(u, 2 * m->N, true);
  rot_bcu_dev(m, o.d, omega, cx, cy);
  o.commit();
  fs::sync();
  FS_API_END
}

int fs_dye_diffuse(fs_csr* K, double* c, double DT, double D) {
  FS_API_BEGIN
  FS_REQUIRE(K && c, "NULL argument");
  Out<double> oc(c, K->n, true);
  DBuf<double> tmp(K->n);
  k_dye_diffuse<<<div_up(K->n, 256), 256, 0, stream()>>>(K->view(), oc.d, DT * D, tmp.p);
  FS_LAUNCH_CHECK();
  FS_CUDA(cudaMemcpyAsync(oc.d, tmp.p, K->n * sizeof(double), cudaMemcpyDeviceToDevice, stream()));
  oc.commit();
  fs::sync();
  FS_API_END
}

int fs_lumped_mass(fs_mesh* m, double* mass) {
  FS_API_BEGIN
  FS_REQUIRE(m && mass, "NULL argument");
  ensure_geom(m);
  FS_CUDA(cudaMemcpyAsync(mass, m->mass.p, m->N * sizeof(double), cudaMemcpyDefault, stream()));
  fs::sync();
  FS_API_END
}

int fs_centroids(fs_mesh* m, int f32, double* cx, double* cy) {
  FS_API_BEGIN
  FS_REQUIRE(m && cx && cy, "NULL argument");
  Out<double> ox(cx, m->T), oy(cy, m->T);
  if (f32) k_centroids<float><<<div_up(m->T, 256), 256, 0, stream()>>>((const double2*)m->coords.p, m->tris.p, m->T, ox.d, oy.d);
  else k_centroids<double><<<div_up(m->T, 256), 256, 0, stream()>>>((const double2*)m->coords.p, m->tris.p, m->T, ox.d, oy.d);
  FS_LAUNCH_CHECK();
  ox.commit(); oy.commit();
  fs::sync();
  FS_API_END
}

int fs_assemble_fem(fs_mesh* m, int f32, const double* g_centroid, double g_const, double* vals, double* b) {
  FS_API_BEGIN
  FS_REQUIRE(m && vals && b, "NULL argument");
  ensure_geom(m);
  if (!m->ke.n) m->ke.alloc(9 * m->T);
  In<double> g(g_centroid, m->T);
  int64_t nb = (m->T + kEB - 1) / kEB;
  int grid = (int)std::min<int64_t>(div_up(nb, kWarpsPerBlock), (int64_t)sm_count() * 16);
  if (f32)
    k_element_stiffness<float, 1><<<grid, kWarpsPerBlock * 32, 0, stream()>>>(
        (const double2*)m->coords.p, m->tris.p, m->T, m->ke.p, g.d, g_const, m->elem_a.p);
  else
    k_element_stiffness<double, 1><<<grid, kWarpsPerBlock * 32, 0, stream()>>>(
        (const double2*)m->coords.p, m->tris.p, m->T, m->ke.p, g.d, g_const, m->elem_a.p);
  FS_LAUNCH_CHECK();
  Out<double> ov(vals, m->pat.nnz), ob(b, m->N);
  assemble_on_pattern(m->pat, m->ke.p, ov.d);
  // returns -BVector (code/poisson.py:146)
  k_node_sum<<<div_up(m->N, 256), 256, 0, stream()>>>(m->inc_ptr.p, m->inc.p, m->elem_a.p, m->N, ob.d, -1.0);
  FS_LAUNCH_CHECK();
  ov.commit(); ob.commit();
  fs::sync();
  FS_API_END
}

int fs_divergence(fs_mesh* m, const double* u, double* div) {
  FS_API_BEGIN
  FS_REQUIRE(m && u && div, "NULL argument");
  In<double> iu(u, 2 * m->N);
  Out<double> od(div, m->N);
  divergence_dev(m, iu.d, od.d, nullptr);
  od.commit();
  fs::sync();
  FS_API_END
}

int fs_gradient(fs_mesh* m, const double* p, double* gx, double* gy) {
  FS_API_BEGIN
  FS_REQUIRE(m && p && gx && gy, "NULL argument");
  In<double> ip(p, m->N);
  Out<double> ox(gx, m->N), oy(gy, m->N);
  gradient_dev(m, ip.d, ox.d, oy.d);
  ox.commit(); oy.commit();
  fs::sync();
  FS_API_END
}

int fs_bc_set(fs_mesh* m, const int32_t* wall, int64_t nw, const int32_t* inner, int64_t ni, const int32_t* pairs,
              int64_t np, const int32_t* interior, int64_t nint) {
  FS_API_BEGIN
  FS_REQUIRE(m, "mesh is NULL");
  FS_REQUIRE(nw >= 0 && ni >= 0 && np >= 0 && nint >= 0, "negative count");
  auto up = [&](DBuf<int>& d, const int32_t* src, int64_t cnt) {
    d.alloc(cnt);
    if (cnt) { FS_REQUIRE(src, "index array is NULL"); d.upload(src, cnt); }
  };
  up(m->inner, inner, ni); up(m->pairs, pairs, 2 * np); up(m->interior, interior, nint);
  {
    // the reference writes the wall values first and the inner-boundary values
    // second (code/StokesColor.py:406-427), so a node in both sets keeps the
    // inner value: drop such nodes from the wall list and one kernel is race free.
    DBuf<int> tmpw, tmpi;
    up(tmpw, wall, nw); up(tmpi, inner, ni);
    std::vector<int> hw = tmpw.to_host(), hi = tmpi.to_host();
    std::sort(hi.begin(), hi.end());
    std::vector<int> eff;
    for (int v : hw) {
      FS_REQUIRE(v >= 0 && v < m->N, "wall index out of range");
      if (!std::binary_search(hi.begin(), hi.end(), v)) eff.push_back(v);
    }
    for (int v : hi) FS_REQUIRE(v >= 0 && v < m->N, "inner index out of range");
    m->wall.alloc(eff.size());
    if (!eff.empty()) m->wall.upload(eff.data(), eff.size());
    fs::sync();
    nw = (int64_t)eff.size();
  }
  m->n_wall = nw; m->n_inner = ni; m->n_pairs = np; m->n_interior = nint;
  m->pairs_host = m->pairs.to_host();
  for (int v : m->pairs_host) FS_REQUIRE(v >= 0 && v < m->N, "pair index out of range");
  m->inner_sin.alloc(ni); m->inner_cos.alloc(ni); m->inner_sin2.alloc(ni);
  if (ni) {
    k_inner_trig<<<div_up(ni, 128), 128, 0, stream()>>>((const double2*)m->coords.p, m->inner.p, ni, m->inner_sin.p,
                                                        m->inner_cos.p, m->inner_sin2.p);
    FS_LAUNCH_CHECK();
  }
  m->bc_ready = true;
  fs::sync();
  FS_API_END
}

int fs_make_per_bcu(fs_mesh* m, double* u) {
  FS_API_BEGIN
  FS_REQUIRE(m && u, "NULL argument");
  Out<double> o(u, 2 * m->N, true);
  per_bcu_dev(m, o.d);
  o.commit();
  fs::sync();
  FS_API_END
}

int fs_make_dir_bcu(fs_mesh* m, double* u, double B1, double B2) {
  FS_API_BEGIN
  FS_REQUIRE(m && u, "NULL argument");
  Out<double> o(u, 2 * m->N, true);
  dir_bcu_dev(m, o.d, B1, B2);
  o.commit();
  fs::sync();
  FS_API_END
}

int fs_reapply_scalar_bc(fs_mesh* m, double* u, const int32_t* pairs_all, int64_t npa, double wv, double iv) {
  FS_API_BEGIN
  FS_REQUIRE(m && u, "NULL argument");
  FS_REQUIRE(m->bc_ready, "fs_bc_set has not been called");
  Out<double> o(u, m->N, true);
  In<int> pr(pairs_all, 2 * npa);
  if (npa) {
    k_scalar_per_seq<<<1, 32, 0, stream()>>>(pr.d, npa, o.d);
    FS_LAUNCH_CHECK();
  }
  int64_t tot = m->n_wall + m->n_inner;
  if (tot) {
    // code/heatEq.py:282-295: inner value wins over wall value (wall list excludes inner nodes)
    k_scalar_dir<<<div_up(tot, 128), 128, 0, stream()>>>(m->wall.p, m->n_wall, m->inner.p, m->n_inner, wv, iv, o.d);
    FS_LAUNCH_CHECK();
  }
  o.commit();
  fs::sync();
  FS_API_END
}

}  // extern "C"
