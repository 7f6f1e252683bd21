// recycle.cuh -- initial guesses for a sequence of solves A x_n = b_n with one SPD operator and slowly varying
// right-hand sides: the two pressure solves of every Stokes step (code/StokesColor.py:554-555, 568-569 -- the reference
// solves each from scratch with a dense LU; an iterative solver can start from what the earlier steps already know).
//
// Fischer's projection (P. F. Fischer, "Projection techniques for iterative solution of Ax = b with successive
// right-hand sides", CMAME 163, 1998): keep an A-orthonormal basis X of the span of earlier solutions; the best
// approximation of the new solution in that span (in the A-norm) is x0 = X (X^T b) -- k dot products and one
// combination, no matrix product.  After the solve the correction x - x0 is A-orthonormalised against X and appended.
// Where Fischer restarts from the last solution alone when the basis is full (which costs a run of expensive solves
// while the basis grows again), the basis here is COMPRESSED: the last `keep` solutions are known through their
// coordinates in the basis, a QR factorisation of those coordinate vectors (on the host, k x keep numbers) gives an
// orthonormal basis of their span as combinations of the old vectors, and the solver continues from there without a
// jump in the iteration counts.  Measured on the bench trajectory (scripts/proto_projected_guess.py, CPU restatement, 262k
// triangles, steps 6-25): 921 PCG iterations with the time-extrapolated warm start, 630 with Fischer's restart
// (k = 12), 542 with compression (k = 12, keep = 6).  Every solve still runs to its tolerance: only the starting point
// changes.
#pragma once
#include <functional>
#include <vector>

#include "internal.cuh"

namespace fs {

constexpr int kRecCap = 18;   // basis vectors one kernel launch can carry

struct Recycler {
  // y = A x on this rank's rows (x, y: n entries; the partitioned step exchanges the halo inside)
  using MatVec = std::function<void(const double* x, double* y)>;
  // part: device array [nchunk][nblk][6] of per-CTA partial sums -> host[6 * nchunk] totals, summed in a fixed order
  // (and over the ranks, identically on every rank, in the partitioned step)
  using Reduce = std::function<void(const double* part, int nblk, int nchunk, double* host)>;

  int64_t n = 0;
  int kmax = 12, keep = 6;
  int k = 0;                               // live basis vectors: X[0..k)
  std::vector<DBuf<double>> X;             // kmax + keep buffers (the spare ones receive a compressed basis)
  std::vector<std::vector<double>> C;      // coordinates in the basis of the last <= keep solutions, oldest first
  std::vector<double> alpha;               // X^T b of the current solve
  DBuf<double> x0, d, Ad, part;
  bool have_x0 = false;
  long long compressions = 0;

  void init(int64_t n_, int kmax_, int keep_);
  bool ready() const { return n > 0; }
  void reset() { k = 0; C.clear(); have_x0 = false; }
  // x <- projection of the solution of A x = b onto the basis; false (x untouched) while the basis is empty
  bool guess(const double* b, double* x, const Reduce& reduce);
  // x: the converged solution of the system guess() was called for (or of the first system)
  void update(const double* x, const MatVec& matvec, const Reduce& reduce);
  // checkpoint: {k, |C|, kmax, keep, C (|C| x k), X (k x n)} as doubles
  int64_t state_size() const { return 4 + (int64_t)C.size() * k + (int64_t)k * n; }
  void get_state(double* host) const;
  void set_state(const double* host, int64_t count);
};

bool recycle_enabled();   // FS_STOKES_RECYCLE=0: the time-extrapolated warm start of round 1 instead
int recycle_kmax();
int recycle_keep();

}  // namespace fs
