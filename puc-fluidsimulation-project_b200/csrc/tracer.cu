// tracer.cu -- point location, semi-Lagrangian dye advection, food tracers and
// the mixing index.
//   PointLocator.find       code/StokesColor.py:314-345  -> k_locate_knn
//   advect_semilagrange     code/StokesColor.py:347-389  -> k_advect_dye
//   mixing_index            code/StokesColor.py:391-403  -> k_mix_pass1/2
//   food tracer step        code/StokesFood.py:482-503   -> k_tracer_step
// The reference's KDTree over triangle centroids becomes a uniform cell grid of
// centroids (cells sorted by a radix sort); the k=10 nearest are found by
// expanding rings of cells, which yields the same ascending-distance candidate
// list as the KDTree query.  The exact containing-triangle search used for the
// food tracers walks through the triangles over an edge-neighbour table,
// starting from the tracer's previous host triangle.
#include <cub/cub.cuh>

#include "internal.cuh"

namespace fs {

struct Locator {
  int G = 0;                  // grid is G x G cells
  double x0 = 0, y0 = 0, h = 1;   // origin and cell size
  int R_max = 1;              // rings that cover the longest edge
  DBuf<double> cen;           // (T,2) centroids
  DBuf<int> cell_start;       // (G*G+1)
  DBuf<int> cell_tri;         // (T) triangle ids sorted by cell (ascending id inside a cell)
  DBuf<double> cen_sorted;    // (T,2) centroids in cell_tri order: a row of cells is one contiguous run
  DBuf<int> nbr;              // (T,3) neighbour across the edge opposite corner i, -1 on the boundary
};

void free_locator(Locator* l) { delete l; }

struct LocView {
  int G;
  double x0, y0, h, inv_h;
  int R_max;
  const double2* cen;
  const int* cell_start;
  const int* cell_tri;
  const double2* cen_s;
  const int* nbr;
  const double2* coords;
  const int* tris;
  int T;
};

// centroid = np.mean(nodes[triangles], axis=1) = ((a+b)+c)/3, code/StokesColor.py:321
__global__ void k_centroid64(const double2* __restrict__ coords, const int* __restrict__ tris, int64_t T,
                             double2* __restrict__ cen, double* __restrict__ emax2) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= T) return;
  double2 a = coords[tris[3 * e]], b = coords[tris[3 * e + 1]], c = coords[tris[3 * e + 2]];
  cen[e] = make_double2(((a.x + b.x) + c.x) / 3.0, ((a.y + b.y) + c.y) / 3.0);
  double l0 = (a.x - b.x) * (a.x - b.x) + (a.y - b.y) * (a.y - b.y);
  double l1 = (b.x - c.x) * (b.x - c.x) + (b.y - c.y) * (b.y - c.y);
  double l2 = (c.x - a.x) * (c.x - a.x) + (c.y - a.y) * (c.y - a.y);
  emax2[e] = fmax(l0, fmax(l1, l2));
}

__device__ __forceinline__ int cell_coord(double v, double o, double inv_h, int G) {
  int c = (int)floor((v - o) * inv_h);
  return min(max(c, 0), G - 1);
}

__global__ void k_cell_keys(const double2* __restrict__ cen, int64_t T, double x0, double y0, double inv_h, int G,
                            unsigned* __restrict__ keys, int* __restrict__ ids) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= T) return;
  int cx = cell_coord(cen[e].x, x0, inv_h, G), cy = cell_coord(cen[e].y, y0, inv_h, G);
  keys[e] = (unsigned)(cy * G + cx);
  ids[e] = (int)e;
}

__global__ void k_gather_cen(const double2* __restrict__ cen, const int* __restrict__ order, int64_t T,
                             double2* __restrict__ out) {
  int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k < T) out[k] = cen[order[k]];
}

__global__ void k_cell_start(const unsigned* __restrict__ skeys, int64_t T, int64_t ncell, int* __restrict__ start) {
  int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k > T) return;
  int64_t lo = (k == 0) ? -1 : (int64_t)skeys[k - 1];
  int64_t hi = (k == T) ? ncell : (int64_t)skeys[k];
  for (int64_t c = lo + 1; c <= hi; ++c) start[c] = (int)k;
}

// ---- edge neighbours --------------------------------------------------------------
__global__ void k_edge_keys(const int* __restrict__ tris, int64_t T, unsigned long long* __restrict__ keys,
                            unsigned* __restrict__ pay) {
  int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= 3 * T) return;
  int64_t e = k / 3;
  int i = (int)(k % 3);                     // edge opposite corner i: corners (i+1, i+2)
  unsigned a = (unsigned)tris[3 * e + (i + 1) % 3], b = (unsigned)tris[3 * e + (i + 2) % 3];
  unsigned lo = min(a, b), hi = max(a, b);
  keys[k] = ((unsigned long long)lo << 32) | hi;
  pay[k] = (unsigned)k;
}

__global__ void k_edge_nbr(const unsigned long long* __restrict__ skeys, const unsigned* __restrict__ spay, int64_t m,
                           int* __restrict__ nbr) {
  int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= m) return;
  bool prev = (k > 0) && skeys[k - 1] == skeys[k];
  bool next = (k + 1 < m) && skeys[k + 1] == skeys[k];
  int other = -1;
  if (prev) other = (int)(spay[k - 1] / 3u);
  else if (next) other = (int)(spay[k + 1] / 3u);
  nbr[spay[k]] = other;
}

static int bits_for(uint64_t v) {
  int b = 1;
  while (b < 64 && (v >> b)) ++b;
  return b;
}

static Locator* build_locator(fs_mesh* m) {
  std::unique_ptr<Locator> L(new Locator());
  cudaStream_t st = stream();
  const int64_t T = m->T;
  L->cen.alloc(2 * T);
  DBuf<double> emax2(T);
  k_centroid64<<<div_up(T, 256), 256, 0, st>>>((const double2*)m->coords.p, m->tris.p, T, (double2*)L->cen.p, emax2.p);
  FS_LAUNCH_CHECK();
  // bounding box and longest edge on the host (setup only)
  std::vector<double> hc = m->coords.to_host();
  double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
  for (int64_t i = 0; i < m->N; ++i) {
    xmin = std::min(xmin, hc[2 * i]); xmax = std::max(xmax, hc[2 * i]);
    ymin = std::min(ymin, hc[2 * i + 1]); ymax = std::max(ymax, hc[2 * i + 1]);
  }
  double emax = std::sqrt(max_abs_dev(emax2.p, T));
  int G = (int)std::floor(std::sqrt((double)T / 2.0));
  G = std::max(1, std::min(G, 8192));
  double ext = std::max(xmax - xmin, ymax - ymin);
  if (!(ext > 0)) ext = 1.0;
  L->G = G; L->x0 = xmin; L->y0 = ymin; L->h = ext / G * (1.0 + 1e-12);
  L->R_max = (int)std::ceil(emax / L->h) + 1;
  const int64_t ncell = (int64_t)G * G;
  DBuf<unsigned> keys(T), keys_alt(T);
  DBuf<int> ids(T), ids_alt(T);
  k_cell_keys<<<div_up(T, 256), 256, 0, st>>>((const double2*)L->cen.p, T, L->x0, L->y0, 1.0 / L->h, G, keys.p, ids.p);
  FS_LAUNCH_CHECK();
  {
    cub::DoubleBuffer<unsigned> kb(keys.p, keys_alt.p);
    cub::DoubleBuffer<int> vb(ids.p, ids_alt.p);
    size_t bytes = 0;
    int nb = bits_for((uint64_t)ncell);
    FS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kb, vb, (int)T, 0, nb, st));
    DBuf<char> tmp(bytes);
    FS_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, kb, vb, (int)T, 0, nb, st));
    count_launch(4);
    L->cell_start.alloc(ncell + 1);
    L->cell_tri.alloc(T);
    k_cell_start<<<div_up(T + 1, 256), 256, 0, st>>>(kb.Current(), T, ncell, L->cell_start.p);
    FS_LAUNCH_CHECK();
    FS_CUDA(cudaMemcpyAsync(L->cell_tri.p, vb.Current(), T * sizeof(int), cudaMemcpyDeviceToDevice, st));
    L->cen_sorted.alloc(2 * T);
    k_gather_cen<<<div_up(T, 256), 256, 0, st>>>((const double2*)L->cen.p, L->cell_tri.p, T, (double2*)L->cen_sorted.p);
    FS_LAUNCH_CHECK();
    FS_CUDA(cudaStreamSynchronize(st));
  }
  {
    const int64_t me = 3 * T;
    DBuf<unsigned long long> ek(me), ek_alt(me);
    DBuf<unsigned> ep(me), ep_alt(me);
    k_edge_keys<<<div_up(me, 256), 256, 0, st>>>(m->tris.p, T, ek.p, ep.p);
    FS_LAUNCH_CHECK();
    cub::DoubleBuffer<unsigned long long> kb(ek.p, ek_alt.p);
    cub::DoubleBuffer<unsigned> vb(ep.p, ep_alt.p);
    size_t bytes = 0;
    FS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kb, vb, (int)me, 0, 64, st));
    DBuf<char> tmp(bytes);
    FS_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, kb, vb, (int)me, 0, 64, st));
    count_launch(8);
    L->nbr.alloc(me);
    k_edge_nbr<<<div_up(me, 256), 256, 0, st>>>(kb.Current(), vb.Current(), me, L->nbr.p);
    FS_LAUNCH_CHECK();
    FS_CUDA(cudaStreamSynchronize(st));
  }
  return L.release();
}

static LocView loc_view(fs_mesh* m) {
  if (!m->loc) m->loc = build_locator(m);
  Locator* L = m->loc;
  return LocView{L->G, L->x0, L->y0, L->h, 1.0 / L->h, L->R_max, (const double2*)L->cen.p, L->cell_start.p,
                 L->cell_tri.p, (const double2*)L->cen_sorted.p, L->nbr.p, (const double2*)m->coords.p, m->tris.p, (int)m->T};
}

// barycentric weights of code/StokesColor.py:334-340; returns false for |det|<1e-14
__device__ __forceinline__ bool bary(const LocView& V, int t, double x, double y, double& w1, double& w2, double& w3) {
  const double2 p1 = __ldg(&V.coords[__ldg(&V.tris[3 * t])]);
  const double2 p2 = __ldg(&V.coords[__ldg(&V.tris[3 * t + 1])]);
  const double2 p3 = __ldg(&V.coords[__ldg(&V.tris[3 * t + 2])]);
  const double det = (p2.x - p1.x) * (p3.y - p1.y) - (p3.x - p1.x) * (p2.y - p1.y);
  if (fabs(det) < 1e-14) return false;
  w1 = ((p2.x - x) * (p3.y - y) - (p3.x - x) * (p2.y - y)) / det;
  w2 = ((p3.x - x) * (p1.y - y) - (p1.x - x) * (p3.y - y)) / det;
  w3 = 1.0 - w1 - w2;
  return true;
}

constexpr int KNN = 10;

// k nearest centroids (ascending distance, ties by id) via expanding cell rings,
// then the first candidate that contains the point.
__device__ int locate_knn(const LocView& V, double x, double y) {
  double bd[KNN];
  int bi[KNN];
#pragma unroll
  for (int k = 0; k < KNN; ++k) { bd[k] = INFINITY; bi[k] = -1; }
  const int cx = cell_coord(x, V.x0, V.inv_h, V.G), cy = cell_coord(y, V.y0, V.inv_h, V.G);
  for (int r = 0; r < V.G; ++r) {
    const int xl = cx - r, xh = cx + r, yl = cy - r, yh = cy + r;
    for (int yy = max(yl, 0); yy <= min(yh, V.G - 1); ++yy) {
      // the new cells of this ring in row yy: the whole span on the two edge rows, else the two
      // end cells.  Cells of a row are consecutive in the sorted arrays, so a span is ONE run.
      const bool edge_row = (yy == yl || yy == yh);
      const int nspan = (edge_row || r == 0) ? 1 : 2;
      for (int sp = 0; sp < nspan; ++sp) {
        int xa, xb;
        if (nspan == 1) { xa = max(xl, 0); xb = min(xh, V.G - 1); }
        else { xa = xb = (sp == 0) ? xl : xh; if (xa < 0 || xa >= V.G) continue; }
        if (xa > xb) continue;
        const int k0 = __ldg(&V.cell_start[yy * V.G + xa]);
        const int k1 = __ldg(&V.cell_start[yy * V.G + xb + 1]);
        for (int k = k0; k < k1; ++k) {
          const double2 cc = __ldg(&V.cen_s[k]);
          const double dx = cc.x - x, dy = cc.y - y;
          double d = dx * dx + dy * dy;
          if (d > bd[KNN - 1]) continue;
          int id = __ldg(&V.cell_tri[k]);
          if (d < bd[KNN - 1] || id < bi[KNN - 1]) {
            // insertion keeping (d, id) ascending; fully unrolled so the arrays stay in registers
#pragma unroll
            for (int q = 0; q < KNN; ++q) {
              const bool less = d < bd[q] || (d == bd[q] && id < bi[q]);
              if (less) {
                const double td = bd[q]; const int ti = bi[q];
                bd[q] = d; bi[q] = id; d = td; id = ti;
              }
            }
          }
        }
      }
    }
    // everything outside the (2r+1)^2 block is farther than dmin
    const double bx0 = V.x0 + xl * V.h, bx1 = V.x0 + (xh + 1) * V.h;
    const double by0 = V.y0 + yl * V.h, by1 = V.y0 + (yh + 1) * V.h;
    double dmin = INFINITY;
    if (xl > 0) dmin = fmin(dmin, x - bx0);
    if (xh < V.G - 1) dmin = fmin(dmin, bx1 - x);
    if (yl > 0) dmin = fmin(dmin, y - by0);
    if (yh < V.G - 1) dmin = fmin(dmin, by1 - y);
    if (dmin == INFINITY) break;                       // block covers the whole grid
    if (dmin > 0 && bd[KNN - 1] < dmin * dmin) break;
  }
#pragma unroll
  for (int k = 0; k < KNN; ++k) {
    const int t = bi[k];
    if (t < 0) break;
    double w1, w2, w3;
    if (!bary(V, t, x, y, w1, w2, w3)) continue;
    if (w1 >= 0.0 && w2 >= 0.0 && w3 >= 0.0) return t;
  }
  return -1;
}

__global__ void __launch_bounds__(128)
k_locate_knn(LocView V, const double2* __restrict__ pts, int64_t P, int* __restrict__ ids) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= P) return;
  double2 p = pts[i];
  ids[i] = locate_knn(V, p.x, p.y);
}

// python-style a % 1.0 (numpy float64 remainder), code/StokesColor.py:361
__device__ __forceinline__ double pymod(double a, double b) {
  double r = fmod(a, b);
  if (r != 0.0) { if ((b < 0.0) != (r < 0.0)) r += b; }
  else r = copysign(0.0, b);
  return r;
}

__device__ __forceinline__ double pdx(double a, double b) {   // :353-357
  double d = a - b;
  if (d > 0.5) d -= 1.0;
  if (d < -0.5) d += 1.0;
  return d;
}

__global__ void __launch_bounds__(128)
k_advect_dye(LocView V, const double2* __restrict__ u, const double* __restrict__ c, double DT, int64_t N,
             double* __restrict__ c_new, int* __restrict__ ids_out) {
  int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (n >= N) return;
  const double2 xn = V.coords[n], un = u[n];
  double xb = pymod(xn.x - DT * un.x * 1.0, 1.0);
  double yb = xn.y - DT * un.y * 1.0;
  if (yb < 0.0) yb = 1e-12;
  if (yb > 1.0) yb = 1.0 - 1e-12;
  const int t = locate_knn(V, xb, yb);
  if (ids_out) ids_out[n] = t;
  if (t < 0) { c_new[n] = c[n]; return; }
  const int i = V.tris[3 * t], j = V.tris[3 * t + 1], k = V.tris[3 * t + 2];
  const double2 p1 = V.coords[i], p2 = V.coords[j], p3 = V.coords[k];
  const double det = pdx(p2.x, p1.x) * (p3.y - p1.y) - pdx(p3.x, p1.x) * (p2.y - p1.y);
  const double w1 = (pdx(p2.x, xb) * (p3.y - yb) - pdx(p3.x, xb) * (p2.y - yb)) / det;
  const double w2 = (pdx(p3.x, xb) * (p1.y - yb) - pdx(p1.x, xb) * (p3.y - yb)) / det;
  const double w3 = 1.0 - w1 - w2;
  c_new[n] = w1 * c[i] + w2 * c[j] + w3 * c[k];
}

// ---- exact containing triangle: walk, then ring search -------------------------------
__device__ int walk_from(const LocView& V, int t, double x, double y, int max_steps) {
  for (int s = 0; s < max_steps && t >= 0; ++s) {
    double w1, w2, w3;
    if (!bary(V, t, x, y, w1, w2, w3)) return -2;
    if (w1 >= 0.0 && w2 >= 0.0 && w3 >= 0.0) return t;
    int i = 0;
    double wm = w1;
    if (w2 < wm) { wm = w2; i = 1; }
    if (w3 < wm) { wm = w3; i = 2; }
    t = __ldg(&V.nbr[3 * t + i]);
  }
  return -2;   // left the mesh or ran out of steps
}

__device__ int locate_exact(const LocView& V, double x, double y, int hint) {
  if (!(x == x) || !(y == y)) return -1;
  if (hint >= 0 && hint < V.T) {
    int t = walk_from(V, hint, x, y, 48);
    if (t >= 0) return t;
  }
  // ring search over the centroid cells: a containing triangle has its centroid
  // within the longest edge length of the point
  const int cx = cell_coord(x, V.x0, V.inv_h, V.G), cy = cell_coord(y, V.y0, V.inv_h, V.G);
  int best = -1;
  for (int r = 0; r <= V.R_max; ++r) {
    const int xl = cx - r, xh = cx + r, yl = cy - r, yh = cy + r;
    for (int yy = max(yl, 0); yy <= min(yh, V.G - 1); ++yy) {
      const bool edge_row = (yy == yl || yy == yh);
      const int step = edge_row ? 1 : max(2 * r, 1);
      for (int xx = xl; xx <= xh; xx += step) {
        if (xx < 0 || xx >= V.G) continue;
        const int c = yy * V.G + xx;
        for (int k = __ldg(&V.cell_start[c]); k < __ldg(&V.cell_start[c + 1]); ++k) {
          const int t = __ldg(&V.cell_tri[k]);
          double w1, w2, w3;
          if (!bary(V, t, x, y, w1, w2, w3)) continue;
          if (w1 >= 0.0 && w2 >= 0.0 && w3 >= 0.0 && (best < 0 || t < best)) best = t;
        }
      }
    }
    if (best >= 0) return best;
  }
  return -1;
}

__global__ void __launch_bounds__(128)
k_locate_exact(LocView V, const double2* __restrict__ pts, int64_t P, int* __restrict__ hint) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= P) return;
  double2 p = pts[i];
  hint[i] = locate_exact(V, p.x, p.y, hint[i]);
}

// code/StokesFood.py:482-503
__global__ void __launch_bounds__(128)
k_tracer_step(LocView V, const double2* __restrict__ u, double2* __restrict__ pts, int* __restrict__ status,
              int* __restrict__ hint, int64_t P, double DT, double L, double cx, double cy, double rcap,
              unsigned long long* __restrict__ eaten, int64_t u_stride) {
  // blockIdx.y = configuration of a batched sweep (own velocity field and tracer set on the shared mesh)
  u += blockIdx.y * u_stride; pts += blockIdx.y * P; status += blockIdx.y * P; hint += blockIdx.y * P; eaten += blockIdx.y;
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int st = 0;
  if (i < P) {
    double2 p = pts[i];
    st = status[i];
    const int t = locate_exact(V, p.x, p.y, hint[i]);
    hint[i] = t;
    if (t < 0) {
      // outside the mesh: LinearTriInterpolator returns a masked NaN, the tracer is lost
      p.x = NAN; p.y = NAN;
    } else {
      double w1, w2, w3;
      bary(V, t, p.x, p.y, w1, w2, w3);
      const double2 ua = __ldg(&u[V.tris[3 * t]]), ub = __ldg(&u[V.tris[3 * t + 1]]), uc = __ldg(&u[V.tris[3 * t + 2]]);
      const double vx = w1 * ua.x + w2 * ub.x + w3 * uc.x;
      const double vy = w1 * ua.y + w2 * ub.y + w3 * uc.y;
      p.x = p.x + vx * DT;
      p.y = p.y + vy * DT;
      p.x = pymod(p.x, L);
      const double dx = p.x - cx, dy = p.y - cy;
      const double dist = sqrt(dx * dx + dy * dy);
      if (dist <= rcap) st = 1;
    }
    pts[i] = p;
    status[i] = st;
  }
  // block count of eaten tracers -> one integer atomic per block (deterministic)
  unsigned ball = __ballot_sync(0xffffffffu, st != 0);
  __shared__ int wsum[4];
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = __popc(ball);
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += wsum[w];
    if (s) atomicAdd(eaten, (unsigned long long)s);
  }
}

// ---- mixing index ------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_mix_pass(const double* __restrict__ c, const double* __restrict__ mass, const int* __restrict__ idx, int64_t n,
           double mu, int pass, double* __restrict__ part) {
  __shared__ double red[2][8];
  double a = 0.0, b = 0.0;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx ? idx[k] : k;
    const double m = mass[i], cv = c[i];
    if (pass == 0) { a += m; b += m * cv; }
    else { const double d = cv - mu; a += m * (d * d); }
  }
  for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double sa = 0.0, sb = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { sa += red[0][w]; sb += red[1][w]; }
    part[2 * blockIdx.x] = sa;
    part[2 * blockIdx.x + 1] = sb;
  }
}

// ---- output sink: what ax.tripcolor(triang, c, shading="gouraud") +
// ax.scatter(tracers) + plt.pause draw (code/StokesColor.py:508-511,593-598, code/StokesFood.py:511-526),
// as a device raster: the nodal field is sampled at the pixel centres by barycentric interpolation in
// the containing triangle (NaN outside the mesh, i.e. in the hole), colour-mapped through a 256-entry
// table, and tracers are drawn as discs on top.
constexpr int kRasterRun = 8;   // consecutive pixels of a row per thread: the previous triangle seeds the walk

__global__ void __launch_bounds__(128)
k_raster_field(LocView V, const double* __restrict__ field, int W, int H, double x0, double dx, double ytop, double dy,
               float* __restrict__ img) {
  const int runs_per_row = (W + kRasterRun - 1) / kRasterRun;
  const long long id = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (id >= (long long)runs_per_row * H) return;
  const int iy = (int)(id / runs_per_row), ix0 = (int)(id % runs_per_row) * kRasterRun;
  const double y = ytop - (iy + 0.5) * dy;
  int t = -1;
  for (int ix = ix0; ix < min(ix0 + kRasterRun, W); ++ix) {
    const double x = x0 + (ix + 0.5) * dx;
    t = locate_exact(V, x, y, t);
    float v = __int_as_float(0x7fc00000);
    if (t >= 0) {
      double w1, w2, w3;
      bary(V, t, x, y, w1, w2, w3);
      v = (float)(w1 * __ldg(field + V.tris[3 * t]) + w2 * __ldg(field + V.tris[3 * t + 1]) + w3 * __ldg(field + V.tris[3 * t + 2]));
    }
    img[(size_t)iy * W + ix] = v;
  }
}

__global__ void k_raster_colormap(const float* __restrict__ img, long long n, float vmin, float vmax,
                                  const unsigned char* __restrict__ lut, uchar4 bg, uchar4* __restrict__ rgba) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = img[i];
  if (!(v == v)) { rgba[i] = bg; return; }
  float s = (vmax > vmin) ? (v - vmin) / (vmax - vmin) : 0.f;
  s = fminf(fmaxf(s, 0.f), 1.f);
  const int k = (int)(s * 255.f + 0.5f);
  rgba[i] = make_uchar4(lut[3 * k], lut[3 * k + 1], lut[3 * k + 2], 255);
}

// discs of radius r pixels; where discs overlap the point with the largest index wins (two passes:
// owner by atomicMax, then colour), so the picture does not depend on the thread schedule
__global__ void k_raster_owner(const double2* __restrict__ pts, long long P, int W, int H, double x0, double inv_dx, double ytop,
                               double inv_dy, float r, int* __restrict__ owner) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= P) return;
  const double2 p = pts[i];
  if (!(p.x == p.x) || !(p.y == p.y)) return;
  const float cx = (float)((p.x - x0) * inv_dx), cy = (float)((ytop - p.y) * inv_dy);   // pixel coordinates (centres at +0.5)
  const int xl = max((int)floorf(cx - r), 0), xh = min((int)ceilf(cx + r), W - 1);
  const int yl = max((int)floorf(cy - r), 0), yh = min((int)ceilf(cy + r), H - 1);
  for (int yy = yl; yy <= yh; ++yy)
    for (int xx = xl; xx <= xh; ++xx) {
      const float ddx = xx + 0.5f - cx, ddy = yy + 0.5f - cy;
      if (ddx * ddx + ddy * ddy <= r * r) atomicMax(&owner[(size_t)yy * W + xx], (int)i);
    }
}

__global__ void k_raster_paint(const int* __restrict__ owner, long long n, const int* __restrict__ status,
                               const unsigned char* __restrict__ colors, int n_colors, uchar4* __restrict__ rgba) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int o = owner[i];
  if (o < 0) return;
  int k = status ? status[o] : 0;
  k = min(max(k, 0), n_colors - 1);
  rgba[i] = make_uchar4(colors[3 * k], colors[3 * k + 1], colors[3 * k + 2], 255);
}

// Velocity arrows (ax.quiver(x, y, u, v, angles='xy', scale_units='xy', scale=s), code/StokesColor.py:514-527,
// code/StokesFood.py:517-519): shaft from (x, y) to (x + u/s, y + v/s) and two head strokes of head_frac x the shaft length
// at +-25 degrees; a pixel is painted when its centre lies within half_width_px of one of the three segments (float32
// pixel arithmetic, same rule in the oracle).  One thread per arrow; all arrows share one colour, so overlaps need no order.
__device__ __forceinline__ float seg_dist2(float px, float py, float ax, float ay, float bx, float by) {
  const float dx = bx - ax, dy = by - ay;
  const float l2 = dx * dx + dy * dy;
  float t = l2 > 0.f ? ((px - ax) * dx + (py - ay) * dy) / l2 : 0.f;
  t = fminf(fmaxf(t, 0.f), 1.f);
  const float qx = ax + t * dx - px, qy = ay + t * dy - py;
  return qx * qx + qy * qy;
}
__global__ void k_raster_quiver(const double2* __restrict__ pts, const double2* __restrict__ vec, long long P, int W, int H, double x0,
                                double inv_dx, double ytop, double inv_dy, double inv_scale, float hw, float head_frac, uchar4 color,
                                uchar4* __restrict__ rgba) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= P) return;
  const double2 p = pts[i], v = vec[i];
  if (!(p.x == p.x) || !(p.y == p.y) || !(v.x == v.x) || !(v.y == v.y)) return;
  const float ax = (float)((p.x - x0) * inv_dx), ay = (float)((ytop - p.y) * inv_dy);
  const float bx = (float)((p.x + v.x * inv_scale - x0) * inv_dx), by = (float)((ytop - (p.y + v.y * inv_scale)) * inv_dy);
  const float sx = ax - bx, sy = ay - by;                       // from the tip back along the shaft
  const float c = 0.90630779f, s = 0.42261826f;                 // cos / sin of 25 degrees
  const float h1x = bx + head_frac * (c * sx - s * sy), h1y = by + head_frac * (s * sx + c * sy);
  const float h2x = bx + head_frac * (c * sx + s * sy), h2y = by + head_frac * (-s * sx + c * sy);
  const float xl = fminf(fminf(ax, bx), fminf(h1x, h2x)) - hw, xh = fmaxf(fmaxf(ax, bx), fmaxf(h1x, h2x)) + hw;
  const float yl = fminf(fminf(ay, by), fminf(h1y, h2y)) - hw, yh = fmaxf(fmaxf(ay, by), fmaxf(h1y, h2y)) + hw;
  const int ix0 = max((int)floorf(xl), 0), ix1 = min((int)ceilf(xh), W - 1);
  const int iy0 = max((int)floorf(yl), 0), iy1 = min((int)ceilf(yh), H - 1);
  const float r2 = hw * hw;
  for (int yy = iy0; yy <= iy1; ++yy)
    for (int xx = ix0; xx <= ix1; ++xx) {
      const float px = xx + 0.5f, py = yy + 0.5f;
      if (seg_dist2(px, py, ax, ay, bx, by) <= r2 || seg_dist2(px, py, bx, by, h1x, h1y) <= r2 || seg_dist2(px, py, bx, by, h2x, h2y) <= r2)
        rgba[(size_t)yy * W + xx] = color;
    }
}

}  // namespace fs

using namespace fs;

extern "C" {

int fs_raster_quiver(uint8_t* rgba, int32_t W, int32_t H, double x0, double x1, double y0, double y1, const double* pts,
                     const double* vec, int64_t P, double scale, double half_width_px, double head_frac, const uint8_t* color_rgb) {
  FS_API_BEGIN
  FS_REQUIRE(rgba && W > 0 && H > 0 && x1 > x0 && y1 > y0 && scale > 0 && color_rgb && (P == 0 || (pts && vec)), "bad arguments");
  if (P == 0) return FS_OK;
  In<double> ip(pts, 2 * P), iv(vec, 2 * P);
  Out<uint8_t> oo(rgba, (size_t)4 * W * H, true);
  const uchar4 col = make_uchar4(color_rgb[0], color_rgb[1], color_rgb[2], 255);
  k_raster_quiver<<<(unsigned)div_up(P, 128), 128, 0, stream()>>>((const double2*)ip.d, (const double2*)iv.d, P, W, H, x0, W / (x1 - x0), y1,
                                                                  H / (y1 - y0), 1.0 / scale, (float)half_width_px, (float)head_frac, col,
                                                                  (uchar4*)oo.d);
  FS_LAUNCH_CHECK();
  oo.commit();
  fs::sync();
  FS_API_END
}

int fs_locate(fs_mesh* m, const double* pts, int64_t P, int32_t* ids) {
  FS_API_BEGIN
  FS_REQUIRE(m && (P == 0 || (pts && ids)), "NULL argument");
  if (P == 0) return FS_OK;
  LocView V = loc_view(m);
  In<double> ip(pts, 2 * P);
  Out<int> oi(ids, P);
  k_locate_knn<<<div_up(P, 128), 128, 0, stream()>>>(V, (const double2*)ip.d, P, oi.d);
  FS_LAUNCH_CHECK();
  oi.commit();
  fs::sync();
  FS_API_END
}

int fs_advect_dye(fs_mesh* m, double* c, const double* u, double DT, int32_t* ids_out) {
  FS_API_BEGIN
  FS_REQUIRE(m && c && u, "NULL argument");
  LocView V = loc_view(m);
  const int64_t N = m->N;
  In<double> iu(u, 2 * N);
  Out<double> oc(c, N, true);
  Out<int> oi(ids_out, N);
  double* cnew = nullptr;
  FS_CUDA(cudaMallocAsync(&cnew, N * sizeof(double), stream()));
  k_advect_dye<<<div_up(N, 128), 128, 0, stream()>>>(V, (const double2*)iu.d, oc.d, DT, N, cnew, oi.d);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(oc.d, cnew, N * sizeof(double), cudaMemcpyDeviceToDevice, stream());
  cudaFreeAsync(cnew, stream());
  FS_CUDA(e);
  oc.commit(); oi.commit();
  fs::sync();
  FS_API_END
}

int fs_mixing_index(fs_mesh* m, const double* c, const double* mass, const int32_t* mask_idx, int64_t n_mask, double* out3) {
  FS_API_BEGIN
  FS_REQUIRE(m && c && mass && out3, "NULL argument");
  const int64_t N = m->N;
  const int64_t n = mask_idx ? n_mask : N;
  In<double> ic(c, N), im(mass, N);
  In<int> ii(mask_idx, mask_idx ? n_mask : 0);
  const int g = std::max(1, std::min(div_up(n, 256), 256));
  DBuf<double> part(2 * g);
  auto run = [&](double mu, int pass, double& a, double& b) {
    k_mix_pass<<<g, 256, 0, stream()>>>(ic.d, im.d, ii.d, n, mu, pass, part.p);
    FS_LAUNCH_CHECK();
    std::vector<double> h = part.to_host();
    a = b = 0.0;
    for (int k = 0; k < g; ++k) { a += h[2 * k]; b += h[2 * k + 1]; }
  };
  double W, mc, sv, dummy;
  run(0.0, 0, W, mc);
  const double mu = mc / W;
  run(mu, 1, sv, dummy);
  const double var = sv / W;
  out3[0] = var / (mu * (1 - mu) + 1e-16);
  out3[1] = mu;
  out3[2] = var;
  FS_API_END
}

int fs_locate_exact(fs_mesh* m, const double* pts, int64_t P, int32_t* hint_ids) {
  FS_API_BEGIN
  FS_REQUIRE(m && (P == 0 || (pts && hint_ids)), "NULL argument");
  if (P == 0) return FS_OK;
  LocView V = loc_view(m);
  In<double> ip(pts, 2 * P);
  Out<int> oh(hint_ids, P, true);
  k_locate_exact<<<div_up(P, 128), 128, 0, stream()>>>(V, (const double2*)ip.d, P, oh.d);
  FS_LAUNCH_CHECK();
  oh.commit();
  fs::sync();
  FS_API_END
}

int fs_tracer_step(fs_mesh* m, double* pts, int32_t* status, int32_t* hint_ids, int64_t P, const double* u, double DT,
                   double L, double cx, double cy, double rcap, int64_t* eaten) {
  FS_API_BEGIN
  FS_REQUIRE(m && u && (P == 0 || (pts && status && hint_ids)), "NULL argument");
  if (P == 0) { if (eaten) *eaten = 0; return FS_OK; }
  LocView V = loc_view(m);
  In<double> iu(u, 2 * m->N);
  Out<double> op(pts, 2 * P, true);
  Out<int> os(status, P, true), oh(hint_ids, P, true);
  DBuf<unsigned long long> cnt(1);
  cnt.zero();
  k_tracer_step<<<div_up(P, 128), 128, 0, stream()>>>(V, (const double2*)iu.d, (double2*)op.d, os.d, oh.d, P, DT, L, cx,
                                                       cy, rcap, cnt.p, 0);
  FS_LAUNCH_CHECK();
  op.commit(); os.commit(); oh.commit();
  unsigned long long h = cnt.to_host()[0];
  if (eaten) *eaten = (int64_t)h;
  FS_API_END
}

// B configurations of a sweep in one launch: pts (B,P,2), status / hint (B,P), u (B,N,2), eaten (B) host
int fs_tracer_step_batch(fs_mesh* m, int32_t B, double* pts, int32_t* status, int32_t* hint_ids, int64_t P, const double* u,
                         double DT, double L, double cx, double cy, double rcap, int64_t* eaten) {
  FS_API_BEGIN
  FS_REQUIRE(m && u && B >= 1 && (P == 0 || (pts && status && hint_ids)), "bad arguments");
  if (P == 0) { if (eaten) for (int c = 0; c < B; ++c) eaten[c] = 0; return FS_OK; }
  FS_REQUIRE(B <= 65535, "at most 65535 configurations per call");
  LocView V = loc_view(m);
  In<double> iu(u, (size_t)B * 2 * m->N);
  Out<double> op(pts, (size_t)B * 2 * P, true);
  Out<int> os(status, (size_t)B * P, true), oh(hint_ids, (size_t)B * P, true);
  DBuf<unsigned long long> cnt(B);
  cnt.zero();
  k_tracer_step<<<dim3(div_up(P, 128), B), 128, 0, stream()>>>(V, (const double2*)iu.d, (double2*)op.d, os.d, oh.d, P, DT, L, cx,
                                                                cy, rcap, cnt.p, m->N);
  FS_LAUNCH_CHECK();
  op.commit(); os.commit(); oh.commit();
  std::vector<unsigned long long> h = cnt.to_host();
  if (eaten) for (int c = 0; c < B; ++c) eaten[c] = (int64_t)h[c];
  FS_API_END
}

int fs_raster_field(fs_mesh* m, const double* field, int32_t W, int32_t H, double x0, double x1, double y0, double y1,
                    float* img) {
  FS_API_BEGIN
  FS_REQUIRE(m && field && img, "NULL argument");
  FS_REQUIRE(W > 0 && H > 0 && x1 > x0 && y1 > y0, "empty raster");
  LocView V = loc_view(m);
  In<double> ifld(field, m->N);
  Out<float> oi(img, (size_t)W * H, false);
  const long long runs = (long long)((W + kRasterRun - 1) / kRasterRun) * H;
  k_raster_field<<<(unsigned)div_up(runs, 128), 128, 0, stream()>>>(V, ifld.d, W, H, x0, (x1 - x0) / W, y1, (y1 - y0) / H, oi.d);
  FS_LAUNCH_CHECK();
  oi.commit();
  fs::sync();
  FS_API_END
}

int fs_raster_colormap(const float* img, int32_t W, int32_t H, double vmin, double vmax, const uint8_t* lut256x3,
                       const uint8_t* background_rgba, uint8_t* rgba) {
  FS_API_BEGIN
  FS_REQUIRE(img && lut256x3 && background_rgba && rgba, "NULL argument");
  FS_REQUIRE(W > 0 && H > 0, "empty raster");
  const long long n = (long long)W * H;
  In<float> ii(img, n);
  In<unsigned char> il(lut256x3, 768);
  Out<unsigned char> oo(rgba, 4 * n, false);
  const uchar4 bg = make_uchar4(background_rgba[0], background_rgba[1], background_rgba[2], background_rgba[3]);
  k_raster_colormap<<<(unsigned)div_up(n, 256), 256, 0, stream()>>>(ii.d, n, (float)vmin, (float)vmax, il.d, bg, (uchar4*)oo.d);
  FS_LAUNCH_CHECK();
  oo.commit();
  fs::sync();
  FS_API_END
}

int fs_raster_points(uint8_t* rgba, int32_t W, int32_t H, double x0, double x1, double y0, double y1, const double* pts,
                     const int32_t* status, int64_t P, const uint8_t* colors_kx3, int32_t n_colors, double radius_px) {
  FS_API_BEGIN
  FS_REQUIRE(rgba && colors_kx3 && (P == 0 || pts), "NULL argument");
  FS_REQUIRE(W > 0 && H > 0 && x1 > x0 && y1 > y0 && n_colors > 0 && radius_px >= 0, "bad raster arguments");
  FS_REQUIRE(P < ((int64_t)1 << 31), "too many points");
  if (P == 0) return FS_OK;
  const long long n = (long long)W * H;
  In<double> ip(pts, 2 * P);
  In<int> is(status, status ? P : 0);
  In<unsigned char> ic(colors_kx3, 3 * (size_t)n_colors);
  Out<unsigned char> oo(rgba, 4 * n, true);
  DBuf<int> owner((size_t)n);
  FS_CUDA(cudaMemsetAsync(owner.p, 0xff, n * sizeof(int), stream()));
  k_raster_owner<<<(unsigned)div_up(P, 128), 128, 0, stream()>>>((const double2*)ip.d, P, W, H, x0, W / (x1 - x0), y1, H / (y1 - y0),
                                                                 (float)radius_px, owner.p);
  FS_LAUNCH_CHECK();
  k_raster_paint<<<(unsigned)div_up(n, 256), 256, 0, stream()>>>(owner.p, n, status ? is.d : nullptr, ic.d, n_colors, (uchar4*)oo.d);
  FS_LAUNCH_CHECK();
  oo.commit();
  fs::sync();
  FS_API_END
}

}  // extern "C"
