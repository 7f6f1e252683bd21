/* fluidsim.h -- C ABI of libfluidsim.so, the B200 (sm_100a) replacement for the
 * per-timestep Stokes hot path of TobiasHoffmannP/PUC-Fluidsimulation-Project.
 *
 * The reference has no FFI: its boundary is a set of module-level Python
 * functions.  Each entry point below names the reference
 * function (file:line under /root/reference) it replaces; the Python mirror in
 * puc-fluidsimulation-project_b200/ binds these with ctypes and keeps the
 * reference's names and signatures.  INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - every function returns int: 0 = ok, <0 = error (fs_last_error() has text);
 *     no C++ exception crosses the boundary; nothing falls back to the CPU.
 *   - array arguments are plain pointers and may be HOST or DEVICE pointers
 *     (detected with cudaPointerGetAttributes).  Host arrays are staged through
 *     device memory inside the call, device arrays are used in place.
 *   - layouts are the reference's: coords (N,2) f64 C-order, triangles (T,3)
 *     i32 0-based, velocity (N,2) f64 C-order, scalars (N,) f64, CSR with i32
 *     rowptr/colidx and f64 values, columns sorted ascending inside a row.
 *   - calls are synchronous with respect to the host unless stated; a handle is
 *     bound to the CUDA device current at creation and is not thread-safe.
 */
#ifndef FLUIDSIM_H
#define FLUIDSIM_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fs_mesh fs_mesh;       /* device mesh + topology              */
typedef struct fs_csr fs_csr;         /* device CSR matrix                   */
typedef struct fs_stokes fs_stokes;   /* operator-split Stokes step state    */

#define FS_OK 0
#define FS_ERR_ARG -1
#define FS_ERR_CUDA -2
#define FS_ERR_IO -3
#define FS_ERR_NOCONV -4   /* Krylov solver hit maxit (result still written) */
#define FS_ERR_INTERNAL -5

#define FS_PRECOND_NONE 0
#define FS_PRECOND_JACOBI 1
#define FS_PRECOND_AMG 2      /* one V(1,1) cycle of smoothed-aggregation AMG (single RHS) */
#define FS_PRECOND_AUTO 3     /* AMG for single-RHS systems with more than 20 000 rows, else Jacobi */

int fs_version(void);
const char* fs_last_error(void);
int fs_device_count(int* count);
int fs_set_device(int device);
/* Bind subsequent work of this host thread to a caller-owned CUDA stream
 * (cudaStream_t cast to void*; NULL = the library's own stream). */
int fs_set_stream(void* cuda_stream);
int fs_sync(void);
/* Device-pointer arguments are read on the library's stream.  A caller that has just written such
 * a buffer on ANOTHER stream (e.g. torch's current stream) must either synchronise, bind the library
 * to that stream with fs_set_stream, or call fs_stream_wait(producer_stream) before the entry: it
 * makes the library stream wait (on the device, no host block) for all work queued so far on the
 * producer.  The Python mirror does this for every torch CUDA tensor it is handed.  Outputs need
 * nothing: every entry returns after its work has completed. */
int fs_stream_wait(void* producer_cuda_stream);
/* number of kernels this library has launched since load (bench.py gpu_launches) */
int64_t fs_launch_count(void);
/* Serialise all ranks' kernels for profiling-free timing: record / read a CUDA
 * event pair on the library stream.  ms = elapsed between the two marks. */
int fs_timer_start(void);
int fs_timer_stop(float* ms);

/* Sampled kernel timing for the roofline report: with every>0 the three passes of
 * every `every`-th single-RHS CG iteration (A: SpMV+dot, B: x/r update+dots,
 * C: p update) are bracketed by CUDA events on the library stream.  fs_profile_read
 * returns the summed milliseconds per pass, the number of sampled iterations and
 * the CG iterations launched since fs_profile(every).  every=0 switches it off. */
int fs_profile(int every);
int fs_profile_read(double* ms3, int64_t* samples, int64_t* iters);
/* AMG-preconditioned CG only: the V-cycle's largest kernel (the finest level's up-sweep SpMV) is
 * timed with its own event pair on every 8th iteration (that cycle runs outside its CUDA graph).
 * Summed milliseconds, number of timed launches, algorithmic bytes of one launch. */
int fs_profile_read_top(double* ms, int64_t* samples, double* bytes_per_launch);

/* ---- ingest: readNode / readEle, code/StokesColor.py:54-95 (fp32 variant
 * code/poisson.py:27-74 = same parse, caller casts).  Two-step: count, then fill. */
int fs_node_file_count(const char* path, int64_t* n_nodes);
int fs_read_node(const char* path, double* coords /* (n,2) host */, int32_t* markers /* (n,) host */,
                 int64_t n_nodes);
int fs_ele_file_count(const char* path, int64_t* n_tris, int32_t* nodes_per_tri);
int fs_read_ele(const char* path, int32_t* tris /* (t,3) host */, int64_t n_tris);

/* ---- mesh handle: uploads the mesh and builds, on the device, the structural
 * CSR pattern of the P1 stiffness matrix, the element->nonzero scatter map, the
 * per-nonzero contribution lists (ascending element order) and the
 * node->element incidence lists.  Replaces the dense np.zeros((N,N)) of
 * buildStiffnessMatrix, code/StokesColor.py:98-101. */
int fs_mesh_create(const double* coords, int64_t n_nodes, const int32_t* tris, int64_t n_tris,
                   const int32_t* markers /* may be NULL */, fs_mesh** out);
int fs_mesh_destroy(fs_mesh* m);
int fs_mesh_sizes(const fs_mesh* m, int64_t* n_nodes, int64_t* n_tris, int64_t* nnz);
int fs_csr_pattern(const fs_mesh* m, int32_t* rowptr /* n+1 */, int32_t* colidx /* nnz */);
int fs_scatter_map(const fs_mesh* m, int32_t* scatter /* (t,9): entry (i,j) at i*3+j */);

/* ---- assembly.
 * fs_assemble_stiffness: buildStiffnessMatrix, code/StokesColor.py:98-128
 *   (2|det| denominator, skip |det|<1e-14), values on the structural pattern,
 *   contributions summed in ascending element order => bit-identical to the
 *   reference's dense matrix.
 * fs_lumped_mass: buildLumpedMassMatrix, code/StokesColor.py:266-284.
 * fs_assemble_fem: buildFemSystem, code/poisson.py:100-146 (signed 2*ADet, skip
 *   only ADet==0, load vector b_j += g(centroid)*area/3, returns -b).
 *   f32_arith != 0 reproduces the reference's float32 coordinate arithmetic.
 *   g_centroid: g evaluated at the element centroids ((t,) f64), or NULL for the
 *   constant g_const.  fs_centroids gives the centroids in the same arithmetic. */
int fs_assemble_stiffness(fs_mesh* m, double* vals /* nnz */);
int fs_lumped_mass(fs_mesh* m, double* mass /* n */);
/* build_mass_and_convection, code/StokesColor.py:286-312 (defined by the reference, used by its convection drafts):
 * consistent mass (area/12 (1 + delta_ij)) and convection (area/3 u_c . grad phi_j) on the structural pattern */
int fs_assemble_mass_convection(fs_mesh* m, const double* u /* (n,2) */, double* m_vals /* nnz */, double* c_vals /* nnz */);
int fs_centroids(fs_mesh* m, int f32_arith, double* cx /* t */, double* cy /* t */);
int fs_assemble_fem(fs_mesh* m, int f32_arith, const double* g_centroid, double g_const,
                    double* vals /* nnz */, double* b /* n */);

/* ---- nodal operators.
 * fs_divergence: calculate_divergence, code/StokesColor.py:130-165.
 * fs_gradient:   calculate_gradiant,  code/StokesColor.py:224-263. */
int fs_divergence(fs_mesh* m, const double* u /* (n,2) */, double* div /* n */);
int fs_gradient(fs_mesh* m, const double* p /* n */, double* gx /* n */, double* gy /* n */);

/* ---- boundary conditions.
 * fs_bc_set: the index sets of code/StokesColor.py:442-464 (computed by the
 *   Python mirror with the reference's rules) are stored on the device.
 * fs_make_per_bcu: makePerBCU, code/StokesColor.py:429-431 (u[slave]=u[master],
 *   sequential pair order).   fs_make_dir_bcu: makeDirBCU, :405-427.
 * fs_reapply_scalar_bc: reapply_periodic_u + reapply_dirchlect_u,
 *   code/heatEq.py:282-301 (pairs = the unfiltered list given here). */
int fs_bc_set(fs_mesh* m, const int32_t* wall, int64_t n_wall, const int32_t* inner, int64_t n_inner,
              const int32_t* pairs /* (np,2) master,slave */, int64_t n_pairs,
              const int32_t* interior, int64_t n_interior);
int fs_make_per_bcu(fs_mesh* m, double* u /* (n,2) */);
int fs_make_dir_bcu(fs_mesh* m, double* u /* (n,2) */, double B1, double B2);
/* rotating inner cylinder (scripts/stokes_report.py:1155-1171): u = omega x (r - centre) on the inner boundary, 0 on the walls */
int fs_make_rot_bcu(fs_mesh* m, double* u /* (n,2) */, double omega, double cx, double cy);
int fs_reapply_scalar_bc(fs_mesh* m, double* u /* n */, const int32_t* pairs_all, int64_t n_pairs_all,
                         double wall_value, double inner_value);

/* ---- sparse matrices and Krylov solvers: replace np.linalg.solve(A, b),
 * code/StokesColor.py:544-545,555,569, code/heatEq.py:323, code/poisson.py:285. */
int fs_csr_create(int64_t n, int64_t nnz, const int32_t* rowptr, const int32_t* colidx,
                  const double* vals, fs_csr** out);
int fs_csr_from_mesh(fs_mesh* m, const double* vals /* nnz */, fs_csr** out); /* shares the pattern */
int fs_csr_destroy(fs_csr* a);
int fs_csr_sizes(const fs_csr* a, int64_t* n, int64_t* nnz);
int fs_csr_get(const fs_csr* a, int32_t* rowptr, int32_t* colidx, double* vals);
int fs_spmv(fs_csr* a, const double* x, double* y);
/* Conjugate gradients, stop when ||r||_2 <= rtol*||b||_2.  x holds the initial
 * guess on entry.  nrhs = 1 or 2 interleaved right-hand sides ((n,nrhs) C-order).
 * project_mean != 0: A is singular with the constants as null space (pressure):
 * b is made mean-free first and the mean of x is removed at the end. */
int fs_cg(fs_csr* a, const double* b, double* x, int nrhs, double rtol, int maxit, int precond,
          int project_mean, int* iters, double* relres);
int fs_bicgstab(fs_csr* a, const double* b, double* x, double rtol, int maxit, int precond,
                int* iters, double* relres);
/* One application z = M^-1 r of the smoothed-aggregation V(1,1) cycle that FS_PRECOND_AMG uses (the hierarchy is
 * built on first use and kept in the handle).  A fixed symmetric positive definite map: what a caller needs to run
 * its own Krylov method around the library's preconditioner.  No reference analogue (every reference solve is a
 * dense LU, code/StokesColor.py:555). */
int fs_precond_apply(fs_csr* a, const double* r, double* z);
/* algorithmic bytes one application of that cycle moves (operator entries once, gather sources and results once):
 * the denominator of bench.py's whole-iteration roofline figure */
int fs_precond_bytes(fs_csr* a, double* bytes_per_apply);

/* ---- the operator-split Stokes step, code/StokesColor.py:537-575 ==
 * code/StokesFood.py:441-479.  fs_stokes_create assembles K, the lumped mass,
 * A_visc = I + DT*nu*K with Dirichlet rows+columns eliminated (:471-475) and the
 * periodic-merged SPD pressure operator Z^T K Z (restatement of :478-479, see
 * DESIGN.md); needs fs_bc_set first.  fs_stokes_step advances u in place: two
 * viscous CG solves (one 2-RHS solve), BCs, divergence, pressure CG, gradient,
 * velocity update, BCs, second projection on the interior nodes -- all on the
 * device with no host round trip except the convergence polls. */
typedef struct fs_stokes_opts {
  double rtol_visc;      /* CG tolerance of the viscous solves    (default 1e-12) */
  double rtol_pressure;  /* CG tolerance of the pressure solves   (default 1e-10) */
  int maxit;             /* per solve                              (default 200000) */
  int precond;           /* FS_PRECOND_* of the pressure solves    (default AUTO) */
  int warm_start;        /* start pressure CG from the previous step's p / p2 (default 1) */
  int final_div;         /* also evaluate final_div (:575) and its max-norm (default 0) */
  int bc_mode;           /* Dirichlet data of the inner boundary: 0 = squirmer (B1,B2), makeDirBCU :405-427 (default);
                            1 = rotating cylinder u = omega x r about (0.5,0.5), scripts/stokes_report.py:1155-1171 */
  double omega;          /* bc_mode 1: angular velocity of this step (the caller ramps it, :1159-1162) */
} fs_stokes_opts;
typedef struct fs_stokes_stats {
  int iters_visc, iters_p1, iters_p2;
  double relres_visc, relres_p1, relres_p2;
  double max_div_ustar, max_final_div;
} fs_stokes_stats;
int fs_stokes_default_opts(fs_stokes_opts* o);
int fs_stokes_create(fs_mesh* m, double DT, double nu, fs_stokes** out);
int fs_stokes_destroy(fs_stokes* s);
int fs_stokes_step(fs_stokes* s, double* u /* (n,2) in/out */, double B1, double B2,
                   const fs_stokes_opts* opts, fs_stokes_stats* stats);
/* p and p2 of the last step, mean-free, expanded to the N nodes */
int fs_stokes_pressure(fs_stokes* s, double* p /* n or NULL */, double* p2 /* n or NULL */);
/* restore those two fields (checkpoint resume; either may be NULL) */
int fs_stokes_set_pressure(fs_stokes* s, const double* p /* n */, const double* p2 /* n */);
/* the two operators, for inspection / parity (borrowed handles, do not destroy) */
int fs_stokes_matrices(fs_stokes* s, fs_csr** a_visc, fs_csr** k_pressure, int32_t* dof /* n or NULL */);
/* the CG warm-start state: the merged-dof pressures of the two solves of the last three steps (the
 * initial guess is extrapolated in time when that lowers the residual), then how many of the older
 * pairs are valid (2 values): 6*n_dof + 2 doubles, n_dof = fs_csr_sizes(k_pressure).
 * Read (set=0) or restore (set=1).  With u this is the complete state of the time loop
 * (checkpoint / resume, repeatable benchmarks). */
int fs_stokes_warm_state(fs_stokes* s, double* q /* 6*n_dof + 2 */, int set);
/* Large pressure systems (>= 20000 merged dofs) start each solve from the projection of the new solution onto the span
 * of the previous ones (an A-orthonormal basis of up to 12 vectors per solve, Fischer 1998; csrc/recycle.cuh) instead
 * of the time extrapolation.  That basis is part of the loop state: read it (set=0; buf may be NULL or too small,
 * *needed always receives the size in doubles) or restore it (set=1, cap = the size read; cap = 0 empties it). */
int fs_stokes_recycle_state(fs_stokes* s, double* buf /* host */, int64_t cap, int set, int64_t* needed);

/* ---- partitioned pressure CG, one rank (process) per GPU, peer memory over NVLink.
 * Replaces the same np.linalg.solve(A_pressure, .) for meshes that are split across GPUs
 * (SURVEY 8e / BASELINE config 5).  Rank r owns a contiguous block of rows of the SPD
 * operator; the local CSR slice has its columns renumbered to [0,n_own) (owned) and
 * [n_own, n_own+n_halo) (external).  The search direction p lives as [own | halo] in a
 * CUDA-IPC shared allocation: inside the persistent CG kernel every rank stores its
 * boundary values of p directly into its neighbours' halo slots and the per-iteration dot
 * products are summed through the same mailboxes in rank order (deterministic); no host
 * or NCCL call happens inside a solve.  Rendezvous (exchange of the 64-byte IPC handles
 * and of the halo index lists, and the two scalars of the init) is the caller's job
 * (torch.distributed in the Python mirror).
 *   fs_dist_create   local matrix                       fs_dist_ipc_handle  this rank's handle
 *   fs_dist_connect  all ranks' handles + send list: own row send_row[k] goes to slot
 *                    send_dst[k] of rank send_peer[k]'s p (sorted by send_row)
 *   fs_dist_cg_begin x0=0: r=b, p=Dinv b; returns the local (b.b, b.Dinv b)
 *   fs_dist_cg_run   globally summed (b.b, r.z) in, x_own out; all ranks call it together */
typedef struct fs_dist fs_dist;
int fs_dist_create(int rank, int world, int64_t n_own, int64_t n_halo, int64_t nnz, const int32_t* rowptr,
                   const int32_t* colidx, const double* vals, fs_dist** out);
int fs_dist_destroy(fs_dist* d);
int fs_dist_ipc_handle(fs_dist* d, void* handle64 /* 64 bytes, host */);
int fs_dist_connect(fs_dist* d, const void* all_handles /* world*64 bytes, host */, const int32_t* send_row,
                    const int32_t* send_peer, const int32_t* send_dst, int64_t n_send /* host arrays */);
int fs_dist_cg_begin(fs_dist* d, const double* b_own, int precond, double* local_sums2 /* host */);
int fs_dist_cg_run(fs_dist* d, double bb_global, double rz_global, double* x_own, double rtol, int maxit,
                   int precond, int* iters, double* relres, double* ns_pass3 /* host, may be NULL */);

/* ---- B configurations of a squirmer (B1,B2) sweep on one GPU in one call (BASELINE config 4).  The configurations
 * share the mesh and therefore A_visc and Z^T K Z; B1, B2 enter only through makeDirBCU (code/StokesColor.py:419).
 * fs_stokes_step_batch runs every kernel of the step once for all of them (grid.y = configuration, the single-CTA
 * Krylov solvers as one CTA per configuration).  Small meshes (N <= 8192, the shipped ones); warm start = each
 * configuration's previous pressures.  u: (B,N,2) in/out, b1b2: (B,2), iters: (B,3) {viscous, p, p2} or NULL.
 * fs_tracer_step_batch: fs_tracer_step for all B tracer sets (pts (B,P,2), status / hint (B,P), u (B,N,2), eaten (B)). */
typedef struct fs_stokes_batch fs_stokes_batch;
int fs_stokes_batch_create(fs_stokes* s, int32_t n_configs, fs_stokes_batch** out);
int fs_stokes_batch_destroy(fs_stokes_batch* b);
int fs_stokes_step_batch(fs_stokes_batch* b, double* u, const double* b1b2, const fs_stokes_opts* opts, int32_t* iters);
int fs_tracer_step_batch(fs_mesh* m, int32_t n_configs, double* pts, int32_t* status, int32_t* hint_ids, int64_t n_pts,
                         const double* u, double DT, double L, double cx, double cy, double rcap, int64_t* eaten /* (B) host */);

/* ---- the whole Stokes step on a mesh cut into contiguous node blocks, one rank (process, GPU) per block
 * (BASELINE config 5: the 32M-triangle annulus on 1/2/4/8 GPUs).  Same sequence as fs_stokes_step
 * (code/StokesColor.py:537-575): 2-RHS viscous CG, BCs, divergence, AMG-preconditioned pressure CG, gradient
 * update, BCs, second projection -- every kernel works on this rank's rows; halo values of the vectors that are
 * read across block boundaries and the partial sums of every dot product travel as peer stores over NVLink
 * issued from inside the kernels (CUDA-IPC arena, monotone sequence flags, deterministic rank-ordered sums);
 * no host, NCCL or copy-engine call happens inside a step.
 *   fs_pstokes_create  global = the replicated global state (fs_stokes_create on the whole mesh; may be destroyed
 *                      afterwards); local_mesh = this rank's sub-mesh: all elements touching its nodes
 *                      [node_split[rank], node_split[rank+1]), local numbering = owned nodes first (global order),
 *                      then the other corners in ascending global id (l2g: local -> global, host), fs_bc_set done
 *                      with the owned part of the index sets.  A periodic pair must lie inside one block.
 *                      AMG levels with more than gather_rows rows (<=0: 100000) are partitioned like the mesh, the
 *                      smaller ones are replicated.
 *   fs_pstokes_ipc_handle / fs_pstokes_connect   exchange of the 64-byte IPC handles is the caller's job
 *                      (torch.distributed.all_gather_object in the Python mirror); all_handles = world x 64 bytes.
 *   fs_pstokes_step    u_own: this rank's rows of the velocity, (n_own,2), host or device, in/out.  Collective.
 *   fs_pstokes_state   warm-start history of the two pressure solves (10*n_own_dofs + 4 doubles), read or restore. */
typedef struct fs_pstokes fs_pstokes;
int fs_pstokes_create(fs_stokes* global, fs_mesh* local_mesh, int rank, int world, const int64_t* node_split /* world+1, host */,
                      const int32_t* l2g /* local nodes, host */, int gather_rows, fs_pstokes** out);
int fs_pstokes_destroy(fs_pstokes* s);
int fs_pstokes_sizes(const fs_pstokes* s, int64_t* n_own_nodes, int64_t* n_halo_nodes, int64_t* n_own_dofs, int64_t* n_halo_dofs,
                     int32_t* levels_partitioned);
int fs_pstokes_ipc_handle(fs_pstokes* s, void* handle64 /* 64 bytes, host */);
int fs_pstokes_connect(fs_pstokes* s, const void* all_handles /* world*64 bytes, host */);
int fs_pstokes_step(fs_pstokes* s, double* u_own /* (n_own,2) in/out */, double B1, double B2, const fs_stokes_opts* opts,
                    fs_stokes_stats* stats);
int fs_pstokes_pressure(fs_pstokes* s, double* p_own /* n_own or NULL */, double* p2_own);
/* profiling aid: `iters` PCG iterations (no convergence test) on the last pressure right-hand side, timed with
 * CUDA events; collective */
int fs_pstokes_profile_pcg(fs_pstokes* s, int iters, double* us_per_iter);
/* FS_DIST_TRACE=1 (environment, read at creation): CTA 0 of the partitioned kernels logs {tag, %globaltimer ns} events;
 * this copies them out (pairs of uint64) and resets the log */
int fs_pstokes_trace(fs_pstokes* s, uint64_t* out /* 2*cap */, int64_t cap, int64_t* n);
int fs_pstokes_state(fs_pstokes* s, double* buf, int set);
/* the projection bases of the two pressure solves, this rank's rows (see fs_stokes_recycle_state) */
int fs_pstokes_recycle_state(fs_pstokes* s, double* buf /* host */, int64_t cap, int set, int64_t* needed);

/* ---- tracers and dye.
 * fs_locate: PointLocator.find, code/StokesColor.py:314-345 -- the 10 nearest
 *   centroids in ascending distance, first triangle with w1,w2,w3 >= 0, else -1.
 * fs_advect_dye: advect_semilagrange, :347-389 (in place on c; ids_out optional).
 * fs_mixing_index: mixing_index, :391-403; out = {I, mu, var}.
 * fs_locate_exact: containing triangle by walking from hint[i] (or a grid seed),
 *   -1 outside the mesh; the tri-finder under code/StokesFood.py:482-486.
 * fs_tracer_step: code/StokesFood.py:482-503 -- interpolate u at the tracers,
 *   forward Euler, wrap x, sticky capture; eaten = sum(status). */
int fs_locate(fs_mesh* m, const double* pts /* (P,2) */, int64_t n_pts, int32_t* ids /* P */);
int fs_advect_dye(fs_mesh* m, double* c /* n */, const double* u /* (n,2) */, double DT,
                  int32_t* ids_out /* n or NULL */);
/* explicit dye diffusion of scripts/good_visualization2.py:704-715: c <- clip(c + DT*D*(K c), 0, 1), K = stiffness handle */
int fs_dye_diffuse(fs_csr* K, double* c /* n */, double DT, double D);
int fs_mixing_index(fs_mesh* m, const double* c, const double* mass, const int32_t* mask_idx,
                    int64_t n_mask, double* out3);
int fs_locate_exact(fs_mesh* m, const double* pts, int64_t n_pts, int32_t* hint_ids /* P in/out */);
int fs_tracer_step(fs_mesh* m, double* pts /* (P,2) in/out */, int32_t* status /* P in/out */,
                   int32_t* hint_ids /* P in/out */, int64_t n_pts, const double* u /* (n,2) */,
                   double DT, double L, double cx, double cy, double rcap, int64_t* eaten);

/* ---- output sink: the picture the reference draws every step with
 * ax.tripcolor(triang, c, shading="gouraud", cmap=..., vmin, vmax) + ax.scatter(tracers) + plt.pause
 * (code/StokesColor.py:508-511,593-598; code/StokesFood.py:511-526), rendered on the device.
 * fs_raster_field: sample the nodal (P1) field at the centres of a W x H pixel grid over
 *   [x0,x1] x [y0,y1] (row 0 is the top row) by barycentric interpolation in the containing
 *   triangle; pixels outside the mesh (the hole) are NaN.  img: H*W floats, row-major.
 * fs_raster_colormap: NaN -> background_rgba (4 bytes, host), else lut[round(255*clamp((v-vmin)/(vmax-vmin)))]
 *   (lut: 256x3 bytes), alpha 255.  rgba: H*W*4 bytes.
 * fs_raster_points: discs of radius_px pixels at the tracer positions, colour colors[status[i]]
 *   (status NULL: colors[0]); where discs overlap, the tracer with the largest index is on top. */
int fs_raster_field(fs_mesh* m, const double* field /* n */, int32_t W, int32_t H, double x0, double x1, double y0,
                    double y1, float* img /* H*W */);
int fs_raster_colormap(const float* img, int32_t W, int32_t H, double vmin, double vmax, const uint8_t* lut256x3,
                       const uint8_t* background_rgba /* host */, uint8_t* rgba /* H*W*4 */);
int fs_raster_points(uint8_t* rgba /* H*W*4 in/out */, int32_t W, int32_t H, double x0, double x1, double y0, double y1,
                     const double* pts /* (P,2) */, const int32_t* status /* P or NULL */, int64_t P,
                     const uint8_t* colors_kx3, int32_t n_colors, double radius_px);

/* fs_raster_quiver: velocity arrows of ax.quiver(x, y, u, v, angles='xy', scale_units='xy', scale=s)
 *   (code/StokesColor.py:514-527, code/StokesFood.py:517-519): shaft (x,y) -> (x+u/s, y+v/s) plus two head strokes of
 *   head_frac x the shaft length at +-25 degrees, painted where a pixel centre is within half_width_px of a stroke.
 *   color_rgb: 3 bytes, host. */
int fs_raster_quiver(uint8_t* rgba /* H*W*4 in/out */, int32_t W, int32_t H, double x0, double x1, double y0, double y1,
                     const double* pts /* (P,2) */, const double* vec /* (P,2) */, int64_t P, double scale,
                     double half_width_px, double head_frac, const uint8_t* color_rgb);

#ifdef __cplusplus
}
#endif
#endif /* FLUIDSIM_H */
