// recycle.cu -- solution-subspace projection for successive right-hand sides (see recycle.cuh).
// Vector kernels: one pass over (k + 1) vectors per multi-dot / combination, deterministic sums (per-CTA partials in
// chunks of 6, reduced in a fixed order by the caller's hook), coefficients passed by value.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "recycle.cuh"
#include "reduce.cuh"

namespace fs {

struct RecPtrs { const double* v[kRecCap]; };
struct RecCoef { double c[kRecCap]; };

// part[(chunk * gridDim.x + blockIdx.x) * 6 + j] = sum_i a[i] * X_{6 chunk + j}[i]   (vectors >= k: 0)
template <int NC>
__global__ void __launch_bounds__(kBlock) k_rec_dots(int64_t n, const double* __restrict__ a, RecPtrs X, int k, double* __restrict__ part) {
  __shared__ double red[6 * 32];
  double acc[6 * NC];
#pragma unroll
  for (int j = 0; j < 6 * NC; ++j) acc[j] = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double av = a[i];
#pragma unroll
    for (int j = 0; j < 6 * NC; ++j)
      if (j < k) acc[j] += av * X.v[j][i];
  }
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    double v[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) v[j] = acc[6 * c + j];
    block_reduce<6>(v, red);
    if (threadIdx.x == 0)
#pragma unroll
      for (int j = 0; j < 6; ++j) part[((size_t)c * gridDim.x + blockIdx.x) * 6 + j] = v[j];
  }
}

// out = beta * in + sum_{j < k} c_j X_j   (in may be null when beta == 0; out may alias in)
template <int NC>
__global__ void __launch_bounds__(kBlock) k_rec_comb(int64_t n, double beta, const double* in, RecCoef c, RecPtrs X, int k, double* out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double v = in ? beta * in[i] : 0.0;
#pragma unroll
    for (int j = 0; j < 6 * NC; ++j)
      if (j < k) v += c.c[j] * X.v[j][i];
    out[i] = v;
  }
}

static int rec_grid(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>(div_up(n, kBlock), 592)); }

static void dots(const Recycler& R, const double* a, int k, const Recycler::Reduce& reduce, double* host /* k */) {
  RecPtrs P{};
  for (int j = 0; j < k; ++j) P.v[j] = R.X[j].p;
  const int nc = div_up(k, 6), g = rec_grid(R.n);
  switch (nc) {
    case 1: k_rec_dots<1><<<g, kBlock, 0, stream()>>>(R.n, a, P, k, R.part.p); break;
    case 2: k_rec_dots<2><<<g, kBlock, 0, stream()>>>(R.n, a, P, k, R.part.p); break;
    default: k_rec_dots<3><<<g, kBlock, 0, stream()>>>(R.n, a, P, k, R.part.p); break;
  }
  FS_LAUNCH_CHECK();
  double tot[kRecCap] = {0};
  reduce(R.part.p, g, nc, tot);
  for (int j = 0; j < k; ++j) host[j] = tot[j];
}

static void comb(const Recycler& R, double beta, const double* in, const double* coef, int k, double* out) {
  RecPtrs P{};
  RecCoef Cf{};
  for (int j = 0; j < k; ++j) { P.v[j] = R.X[j].p; Cf.c[j] = coef[j]; }
  const int nc = std::max(1, div_up(k, 6)), g = rec_grid(R.n);
  switch (nc) {
    case 1: k_rec_comb<1><<<g, kBlock, 0, stream()>>>(R.n, beta, in, Cf, P, k, out); break;
    case 2: k_rec_comb<2><<<g, kBlock, 0, stream()>>>(R.n, beta, in, Cf, P, k, out); break;
    default: k_rec_comb<3><<<g, kBlock, 0, stream()>>>(R.n, beta, in, Cf, P, k, out); break;
  }
  FS_LAUNCH_CHECK();
}

bool recycle_enabled() {     // read at every solve: tests compare the two warm starts in one process
  const char* e = std::getenv("FS_STOKES_RECYCLE");
  return !e || std::atoi(e) != 0;
}
int recycle_kmax() {
  static const int v = [] { const char* e = std::getenv("FS_RECYCLE_K"); return std::max(2, std::min(kRecCap - 1, e ? std::atoi(e) : 12)); }();
  return v;
}
int recycle_keep() {
  static const int v = [] { const char* e = std::getenv("FS_RECYCLE_KEEP"); return e ? std::atoi(e) : 6; }();
  return std::max(1, std::min(v, recycle_kmax() - 1));
}

void Recycler::init(int64_t n_, int kmax_, int keep_) {
  n = n_;
  kmax = std::max(2, std::min(kmax_, kRecCap - 1));
  keep = std::max(1, std::min(keep_, kmax - 1));
  X.clear();
  X.resize((size_t)kmax + keep);
  for (auto& b : X) b.alloc((size_t)n);
  x0.alloc((size_t)n);
  d.alloc((size_t)n);
  Ad.alloc((size_t)n);
  part.alloc((size_t)3 * 6 * 1024);
  reset();
}

bool Recycler::guess(const double* b, double* x, const Reduce& reduce) {
  have_x0 = false;
  if (k == 0) return false;
  alpha.assign((size_t)k, 0.0);
  dots(*this, b, k, reduce, alpha.data());
  comb(*this, 0.0, nullptr, alpha.data(), k, x0.p);
  FS_CUDA(cudaMemcpyAsync(x, x0.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, stream()));
  have_x0 = true;
  return true;
}

void Recycler::update(const double* x, const MatVec& matvec, const Reduce& reduce) {
  const bool proj = have_x0 && k > 0;
  have_x0 = false;
  if (!proj && k > 0) reset();      // a solve that did not start from the projection: its coordinates are unknown
  // d = x - x0, A-orthogonalised against the basis (in exact arithmetic it already is: the projection is Galerkin)
  std::vector<double> coords((size_t)k + 1, 0.0);
  if (proj) {
    const double m1 = -1.0;
    RecPtrs P{};
    RecCoef Cf{};
    P.v[0] = x0.p; Cf.c[0] = m1;
    k_rec_comb<1><<<rec_grid(n), kBlock, 0, stream()>>>(n, 1.0, x, Cf, P, 1, d.p);
    FS_LAUNCH_CHECK();
    matvec(d.p, Ad.p);
    std::vector<double> c((size_t)k, 0.0);
    dots(*this, Ad.p, k, reduce, c.data());
    std::vector<double> mc(c);
    for (auto& v : mc) v = -v;
    comb(*this, 1.0, d.p, mc.data(), k, d.p);
    for (int j = 0; j < k; ++j) coords[(size_t)j] = alpha[(size_t)j] + c[(size_t)j];
  } else {
    FS_CUDA(cudaMemcpyAsync(d.p, x, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, stream()));
  }
  matvec(d.p, Ad.p);
  double nrm2 = 0.0;
  {
    // d . A d through the same multi-dot (one vector): temporarily the candidate sits in slot k
    std::swap(X[(size_t)k], d);
    RecPtrs P{};
    P.v[0] = X[(size_t)k].p;
    k_rec_dots<1><<<rec_grid(n), kBlock, 0, stream()>>>(n, Ad.p, P, 1, part.p);
    FS_LAUNCH_CHECK();
    double tot[kRecCap] = {0};
    reduce(part.p, rec_grid(n), 1, tot);
    nrm2 = tot[0];
  }
  double have2 = 0.0;
  for (int j = 0; j < k; ++j) have2 += coords[(size_t)j] * coords[(size_t)j];
  const bool finite = std::isfinite(nrm2) && std::isfinite(have2);
  if (!finite) { reset(); return; }
  // a correction below 1e-13 of the solution (A-norm) is rounding noise: the solution already lies in the span
  const bool add = nrm2 > 0.0 && nrm2 > 1e-26 * have2;
  if (add) {
    const double nrm = std::sqrt(nrm2);
    RecPtrs P{};
    RecCoef Cf{};
    k_rec_comb<1><<<rec_grid(n), kBlock, 0, stream()>>>(n, 1.0 / nrm, X[(size_t)k].p, Cf, P, 0, X[(size_t)k].p);
    FS_LAUNCH_CHECK();
    coords[(size_t)k] = nrm;
    for (auto& c : C) c.push_back(0.0);
    ++k;
  } else {
    coords.pop_back();
    if (k == 0) return;
  }
  C.push_back(coords);
  if ((int)C.size() > keep) C.erase(C.begin());
  if (k < kmax) return;
  // ---- compress: orthonormal basis (Euclidean in coordinates = A-orthonormal in vectors) of the span of the kept
  // solutions, newest first; modified Gram-Schmidt, twice
  std::vector<std::vector<double>> Q;
  double first = 0.0;
  for (int s = (int)C.size() - 1; s >= 0; --s) {
    std::vector<double> v = C[(size_t)s];
    for (int pass = 0; pass < 2; ++pass)
      for (const auto& q : Q) {
        double h = 0.0;
        for (int i = 0; i < k; ++i) h += q[(size_t)i] * v[(size_t)i];
        for (int i = 0; i < k; ++i) v[(size_t)i] -= h * q[(size_t)i];
      }
    double nv = 0.0;
    for (int i = 0; i < k; ++i) nv += v[(size_t)i] * v[(size_t)i];
    nv = std::sqrt(nv);
    if (Q.empty()) first = nv;
    if (!(nv > 1e-10 * first) || nv == 0.0) continue;
    for (int i = 0; i < k; ++i) v[(size_t)i] /= nv;
    Q.push_back(std::move(v));
  }
  const int m = (int)Q.size();
  if (m == 0) { reset(); return; }
  for (int j = 0; j < m; ++j) comb(*this, 0.0, nullptr, Q[(size_t)j].data(), k, X[(size_t)kmax + j].p);
  for (auto& c : C) {
    std::vector<double> t((size_t)m, 0.0);
    for (int j = 0; j < m; ++j)
      for (int i = 0; i < k; ++i) t[(size_t)j] += Q[(size_t)j][(size_t)i] * c[(size_t)i];
    c = std::move(t);
  }
  for (int j = 0; j < m; ++j) std::swap(X[(size_t)j], X[(size_t)kmax + j]);
  k = m;
  ++compressions;
}

void Recycler::get_state(double* host) const {
  host[0] = (double)k; host[1] = (double)C.size(); host[2] = (double)kmax; host[3] = (double)keep;
  double* p = host + 4;
  for (const auto& c : C) { std::copy(c.begin(), c.end(), p); p += k; }
  for (int j = 0; j < k; ++j, p += n)
    FS_CUDA(cudaMemcpyAsync(p, X[(size_t)j].p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, stream()));
  FS_CUDA(cudaStreamSynchronize(stream()));
}

void Recycler::set_state(const double* host, int64_t count) {
  reset();
  FS_REQUIRE(count >= 4, "recycle state: truncated");
  const int k_ = (int)host[0], nc = (int)host[1];
  FS_REQUIRE(k_ >= 0 && k_ < kmax + 1 && nc >= 0 && nc <= kRecCap, "recycle state: bad header");
  FS_REQUIRE(count == 4 + (int64_t)nc * k_ + (int64_t)k_ * n, "recycle state: size does not match its header");
  const double* p = host + 4;
  for (int s = 0; s < nc; ++s, p += k_) C.emplace_back(p, p + k_);
  for (int j = 0; j < k_; ++j, p += n)
    FS_CUDA(cudaMemcpyAsync(X[(size_t)j].p, p, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, stream()));
  FS_CUDA(cudaStreamSynchronize(stream()));
  k = k_;
}

}  // namespace fs
