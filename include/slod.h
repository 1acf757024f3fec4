/*
 * slod.h -- C ABI of the B200-native SLOD offline phase (libslod_b200.so).
 *
 * The reference (camillabelponer/dealii-slod) has no FFI layer: the seam this library replaces is
 * the pair of protected member functions
 *     LOD::compute_basis_function_candidates()   source/LOD.cc:296-768   (declared include/LOD.h:176)
 *     LOD::assemble_global_matrix()              source/LOD.cc:860-973   (declared include/LOD.h:178)
 * called from LOD::run() (source/LOD.cc:1433-1434), together with the integer set-up they depend on
 * (create_patches source/LOD.cc:122-244, create_mesh_for_patch :770-858, fill_dofs_indices_vector
 * include/LODtools.h:334-375) and the PDE hook assemble_stiffness (include/Diffusion.h:111-207,
 * include/Elasticity.h:163-299).  On top of those two, the stages that consume their results are exported as well
 * (SURVEY section 8f rows 1-2): LOD::solve() and the prolongation (source/LOD.cc:975-1001, :1251), the fine FEM reference
 * solve and the error norms (source/LOD.cc:1004-1094, :1252).  INTEGRATION.md shows the reference-side binding.
 *
 * Conventions
 *  - plain C, opaque handle, no exceptions cross the boundary; every call returns an int status
 *    (SLOD_OK == 0) and slod_last_error() gives the message of the last failure on that handle.
 *  - all floating point data is fp64; all buffers are caller-owned.
 *  - host-buffer calls return when their results are complete, with one exception: slod_assemble_coarse only enqueues
 *    (see there).  The device-buffer entry points (slod_*_device) only enqueue on the caller's stream;
 *    slod_synchronize() is their synchronisation point and reports numerical failures of the enqueued patches.
 *  - multi-GPU, two ways, both with NCCL inside the library (libnccl.so.2 is loaded on first use, a single-GPU handle
 *    never needs it): (a) slod_params.n_gpus = N -- one handle, one process, N devices; slod_compute_basis and
 *    slod_assemble_coarse split the patches over the devices and combine the results over NVLink, every other call
 *    behaves as on one device; (b) one process and one handle per GPU (MPI / torchrun style): slod_comm_init joins the
 *    handles into one NCCL communicator and slod_offline_distributed runs the phase with its two all-gathers.
 *    The slod_*_device entry points remain for callers that do their own exchange.
 *  - there is NO CPU fallback: if no CUDA device is usable slod_create fails with SLOD_ERR_CUDA.
 *  - a handle is thread-compatible, not thread-safe (one call at a time per handle).  DIFFERENT handles may be used
 *    concurrently from different host threads and streams, on the same device too: the kernels read their parameters
 *    from one __constant__ block per device, and every call that launches kernels binds its handle's block for its
 *    scope (the calls of different handles on one device enqueue one after the other; the block is re-uploaded,
 *    stream-ordered, only when the resident bytes differ).
 *  - patch size: n_subdivisions >= 2 (with 1 the matrix P^T A^-1 P of some patch is singular: refused with
 *    SLOD_ERR_UNSUPPORTED) and any oversampling whose patches have at most 128 coarse dofs; patches
 *    too large for the shared-memory solvers (3-D, 4 subdivisions, oversampling 2) run through a direct banded solver
 *    with its windows in global memory (slower), like the reference's direct solver (include/LODtools.h:575-580).
 *  - patch id == active-cell index of the centre cell after refine_global (Morton / Z-order, x low
 *    bit), exactly as in source/LOD.cc:184-192.
 *  - patch-local fine nodes are numbered lexicographically (x fastest) on the patch's own node box
 *    (p_a = m_a * n_subdivisions + 1 nodes along axis a); a fine DoF is spacedim*node + comp.
 *    slod_get_patch_fine_dofs() maps them to deal.II's global fine DoF numbers,
 *    slod_get_patch_local_dofs() to deal.II's patch-local numbers (Patch::basis_function layout,
 *    include/LOD.h:79).
 */
#ifndef SLOD_B200_H
#define SLOD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct slod_ctx slod_ctx;

enum {
  SLOD_OK = 0,
  SLOD_ERR_INVALID = 1,      /* bad argument / inconsistent parameters                  */
  SLOD_ERR_UNSUPPORTED = 2,  /* configuration outside the implemented envelope           */
  SLOD_ERR_CUDA = 3,         /* CUDA runtime error or no usable device                   */
  SLOD_ERR_STATE = 4,        /* call order violated (e.g. basis requested before compute) */
  SLOD_ERR_NUMERIC = 5       /* a patch problem was not SPD / an eigen-solver or CG did not converge */
};

enum { SLOD_PROBLEM_DIFFUSION = 0, SLOD_PROBLEM_ELASTICITY = 1 };
/* slod_params.device: -1 = current CUDA device; SLOD_DEVICE_NONE = no device, integer maps only
 * (patch lists, DoF maps, CSR pattern) -- every compute call on such a handle returns SLOD_ERR_CUDA. */
enum { SLOD_DEVICE_CURRENT = -1, SLOD_DEVICE_NONE = -2 };

/* Mirrors LODParameters (include/LOD.h:85-157) + the template arguments <dim, spacedim>. */
typedef struct {
  int dim;                  /* 2 or 3 (3 only with spacedim 1)                                   */
  int spacedim;             /* 1 diffusion, 2 = dim elasticity                                    */
  int n_global_refinements; /* "Number of global refinements"                 include/LOD.h:95   */
  int n_subdivisions;       /* "Number of subdivisions" (power of two)        include/LOD.h:94   */
  int oversampling;         /* "Oversampling"                                 include/LOD.h:93   */
  int stabilize;            /* "Stabilize phi_LOD candidates" (SLOD)          include/LOD.h:97   */
  int problem;              /* SLOD_PROBLEM_*                                                     */
  int quirk_presaved;       /* reproduce source/LOD.cc:354-362 ("Constant problem coefficients") */
  int device;               /* CUDA device ordinal, SLOD_DEVICE_CURRENT or SLOD_DEVICE_NONE        */
  int n_gpus;               /* 0 or 1: the handle drives `device` alone.  N > 1: ONE handle drives the N devices
                               device .. device + N - 1 of this process (SLOD_DEVICE_CURRENT counts as 0): the patches
                               are split into contiguous even ranges (the reference's locally_owned_patches,
                               source/LOD.cc:116-118), one host thread and one NCCL rank per device inside the library */
} slod_params;

/* life cycle ------------------------------------------------------------------------------------*/
int slod_create(const slod_params *par, slod_ctx **out);
void slod_destroy(slod_ctx *ctx);
const char *slod_last_error(const slod_ctx *ctx);
/* message of the last failed slod_create (which has no handle to ask) */
const char *slod_last_create_error(void);

/* coefficients: replaces problem_parameter (include/Diffusion.h:7-54).  `cellwise` holds one value
 * per cell of a (2^eta_refinement)^dim grid, lexicographic x fastest.  field 0 = alpha | lambda,
 * field 1 = mu.  If the table is finer than the fine sub-cells (eta < h, as in tests/Poisson_LOD_Example)
 * it is sampled at the Gauss points like the reference does; that mode is available in 2-D. */
int slod_set_coefficient(slod_ctx *ctx, int field, int eta_refinement, const double *cellwise, size_t n);

/* integer maps (bit-exact with the reference / the oracle) -----------------------------------------*/
int slod_patch_count(const slod_ctx *ctx, int64_t *n_patches);
/* sizes of one patch: cells, fine DoFs, interior DoFs, patch-boundary DoFs (id 99), domain-boundary
 * DoFs (id 0), coarse DoFs; and the node box m[a] cells / lo[a] first coarse cell per axis. */
int slod_get_patch_info(const slod_ctx *ctx, int64_t patch, int32_t *n_cells, int32_t *n_fine, int32_t *n_internal,
                        int32_t *n_boundary, int32_t *n_domain_boundary, int32_t *n_coarse, int32_t lo[3],
                        int32_t m[3]);
/* Patch::cells (active-cell ids, centre first, x-offset outer)            source/LOD.cc:151-201 */
int slod_get_patch_cells(const slod_ctx *ctx, int64_t patch, uint32_t *cells, int32_t *n);
/* lexicographic patch DoF -> deal.II global fine DoF (dof_handler_fine)    source/LOD.cc:925-929 */
int slod_get_patch_fine_dofs(const slod_ctx *ctx, int64_t patch, uint64_t *global_dofs, int32_t *n);
/* lexicographic patch DoF -> deal.II patch-local DoF (dh_fine_patch)       source/LOD.cc:365-366 */
int slod_get_patch_local_dofs(const slod_ctx *ctx, int64_t patch, uint32_t *local_dofs, int32_t *n);
/* fill_dofs_indices_vector (lexicographic patch numbering, ascending)  include/LODtools.h:334-375
 * which: 0 internal, 1 patch boundary (id 99), 2 domain boundary (id 0). */
int slod_get_patch_dof_class(const slod_ctx *ctx, int64_t patch, int which, uint32_t *dofs, int32_t *n);

/* the hot path, host buffers ------------------------------------------------------------------------*/
/* stages assembly .. premultiply for every patch (source/LOD.cc:345-767), batched on the GPU. */
int slod_compute_basis(slod_ctx *ctx);
/* Patch::basis_function[comp] and basis_function_premultiplied[comp] (include/LOD.h:79-80) in
 * lexicographic patch numbering; either pointer may be NULL. n_fine values each. */
int slod_get_basis(const slod_ctx *ctx, int64_t patch, int comp, double *phi, double *A_phi);
/* all patches at once: arrays [n_patches][spacedim][stride] with stride = slod_basis_stride(). */
int slod_basis_stride(const slod_ctx *ctx, int64_t *stride);
int slod_get_all_basis(const slod_ctx *ctx, double *phi, double *A_phi);
/* global_stiffness_matrix = C^T (A C)  (source/LOD.cc:860-973).  The kernels are enqueued and the call returns: every
 * consumer of the matrix (slod_get_coarse_csr, slod_coarse_solve, ...) is ordered behind them and reports an execution
 * error at its own synchronisation, and slod_get_all_basis, called in between, copies while they run. */
int slod_assemble_coarse(slod_ctx *ctx);
/* CSR of the coarse matrix (rows/cols = spacedim*patch + comp, columns ascending; structural zeros of
 * the sparse product are kept).  Call with rowptr == NULL to query n_rows and nnz. */
int slod_get_coarse_csr(const slod_ctx *ctx, int64_t *rowptr, int64_t *col, double *val, int64_t *n_rows,
                        int64_t *nnz);

/* ---- online phase on the handle's basis and coarse matrix (replaces LOD::solve(),
 * source/LOD.cc:975-1001, and the prolongation in LOD::compare_lod_with_fem(), source/LOD.cc:1251).
 * Fine vectors are lexicographic: index = node * spacedim + component, node = x + G (y + G z), G = N n + 1 nodes per
 * axis; slod_get_patch_fine_dofs gives the deal.II numbering of the same nodes.  Coarse vectors: spacedim*patch + comp. */
int slod_fine_size(const slod_ctx *ctx, int64_t *n_fine);
/* system_rhs = C^T fem_rhs  (basis_matrix_transposed.Tvmult(system_rhs, fem_rhs), source/LOD.cc:981). */
int slod_coarse_rhs(slod_ctx *ctx, const double *f_fine, double *rhs_coarse);
/* K u = rhs by conjugate gradients from u = 0 (SolverCG, source/LOD.cc:994-998), stopped like deal.II's
 * ReductionControl ("Coarse solver control", include/LOD.h:109): ||r||_2 <= tolerance or ||r||_2 <= reduction * ||r_0||_2.
 * Not converged within max_steps -> SLOD_ERR_NUMERIC (deal.II throws SolverControl::NoConvergence).  The preconditioner
 * is diag(K), not the reference's SSOR(1.2): same solution to the tolerance, different step counts.
 * steps / residual (may be NULL) receive the number of steps taken and the final ||r||_2. */
int slod_coarse_solve(slod_ctx *ctx, const double *rhs_coarse, double *u_coarse, int32_t max_steps, double tolerance,
                      double reduction, int32_t *steps, double *residual);
/* lod_solution = C u  (basis_matrix_transposed.vmult(lod_solution, solution), source/LOD.cc:1251). */
int slod_prolongate(slod_ctx *ctx, const double *u_coarse, double *u_fine);

/* ---- fine-scale reference problem (the solve of LOD::assemble_and_solve_fem_problem,
 * source/LOD.cc:1004-1094, and the norms of compare_lod_with_fem, source/LOD.cc:1252).  Needs only the coefficient.
 * A u = f on the global Q_iso_Q1 grid with homogeneous Dirichlet conditions (boundary rows of f are ignored, u = 0
 * there), matrix free, by diag-preconditioned conjugate gradients (the reference: CG + AMG) with the same stopping
 * rule and error behaviour as slod_coarse_solve. */
int slod_fem_solve(slod_ctx *ctx, const double *f_fine, double *u_fine, int32_t max_steps, double tolerance,
                   double reduction, int32_t *steps, double *residual);
/* Norms of a fine vector (outputs may be NULL): ||v||_L2 = sqrt(v.Mv) and |v|_H1 = sqrt(v.Lv) with the exact Q1 mass
 * and Laplace matrices of the sub-cell grid, and the energy norm sqrt(v.Av) of the problem's own bilinear form.  The
 * reference integrates |u_fem - u_lod| with a Gauss rule on the coarse cells (ParsedConvergenceTable::difference),
 * which is not exact for Q_iso_Q1 functions; these are the exact values of the same norms. */
int slod_fine_norms(slod_ctx *ctx, const double *v_fine, double *l2, double *h1_semi, double *energy);
/* The same norms the way the reference's error tables compute them (ParsedConvergenceTable::difference,
 * source/LOD.cc:1252; include/LOD.h:111-115): VectorTools::integrate_difference with QGauss((degree + 1) * 2) on the cells
 * of dof_handler_fine, i.e. 2 (n_subdivisions + 1) Gauss points per direction on every COARSE cell; norms of the table's
 * default list: L2_norm, Linfty_norm (maximum over the quadrature points and components), H1_norm (the full norm
 * sqrt(L2^2 + |.|_H1^2)).  Outputs may be NULL. */
int slod_fine_norms_reference(slod_ctx *ctx, const double *v_fine, double *l2, double *linfty, double *h1);

/* ---- checkpoint of the offline phase (the reference has none: it recomputes every run) ----------
 * slod_save_state writes the parameters, the basis (phi, A*phi) and, if assembled, the coarse matrix to one binary
 * file; slod_load_state restores them into a handle created with the SAME parameters (else SLOD_ERR_INVALID), after
 * which the online entry points (slod_coarse_rhs / _solve / slod_prolongate, slod_get_basis, slod_get_coarse_csr) work
 * without recomputing anything.  The coefficient is not part of the file (slod_fem_solve / slod_fine_norms need it set). */
int slod_save_state(slod_ctx *ctx, const char *path);
int slod_load_state(slod_ctx *ctx, const char *path);

/* Page-locked host memory for the caller-owned output buffers of slod_get_all_basis / slod_get_coarse_csr: copies
 * into pageable memory work too but run at a fraction of the link speed.  No reference counterpart. */
int slod_alloc_host(size_t bytes, void **out);
int slod_free_host(void *p);

/* diagnostics: per patch and component 8 doubles:
 *   [0] ||d||_inf before truncation  [1] truncation steps  [2] sigma_0  [3] smallest kept sigma
 *   [4] reserved  [5] selection path (0 LOD branch, 1 Cholesky fast path, 2 Jacobi fallback, 3 tridiagonal QL)
 *   [6] QL iterations / Jacobi sweeps [7] status bits */
int slod_get_patch_diagnostics(const slod_ctx *ctx, int64_t patch, int comp, double out[8]);
/* stage intermediates of one patch for staged parity tests (recomputed on demand for that patch):
 * X = A_ii^{-1} P_i  (n_internal x n_coarse, row-major), Minv (n_coarse^2), BD^T BD (n_coarse^2). Any may be NULL. */
int slod_debug_patch_stages(slod_ctx *ctx, int64_t patch, double *X, double *Minv, double *G);

/* FP64 throughput of the handle's device in TFLOP/s, measured now with two register-resident probe kernels: a DFMA chain
 * and an mma.sync.m8n8k4.f64 chain (the instruction of the tensor-core kernels).  The benchmark's roofline denominator. */
int slod_measure_fp64_peak(slod_ctx *ctx, double *dfma_tflops, double *dmma_tflops);

/* per-kernel device times of the last compute/assemble (CUDA events), ms:
 *   [0] patch solve  [1] dense (M, BD, Gram)  [2] selection (eigen)  [3] finish (phi, A phi)
 *   [4] coarse matrix  [5] factorisation share of [0] (split solver, first chunk)  [6], [7] reserved */
int slod_get_timings(const slod_ctx *ctx, double *ms, int n);

/* the hot path, device buffers (multi-GPU plumbing) ----------------------------------------------------*/
/* Compute the basis of patches [patch_begin, patch_end) into device arrays laid out
 * [n_patches][spacedim][stride]; only the rows of the range are written.  `stream` is a cudaStream_t. */
int slod_compute_basis_device(slod_ctx *ctx, int64_t patch_begin, int64_t patch_end, double *d_phi, double *d_A_phi,
                              void *stream);
/* Coarse-matrix rows of patches [patch_begin, patch_end) in block-ELL form: d_K is
 * [n_patches*spacedim][ell_width] with ell_width = (4*oversampling+3)^dim * spacedim; slot of neighbour
 * offset D and component e is ((Dz+w)*(2w+1)+(Dy+w))*(2w+1)+(Dx+w))*spacedim+e, w = 2*oversampling+1.
 * Needs d_phi of the range and d_A_phi of ALL patches (all-gathered by the caller). */
int slod_assemble_coarse_device(slod_ctx *ctx, int64_t patch_begin, int64_t patch_end, const double *d_phi,
                                const double *d_A_phi, double *d_K, void *stream);
/* Synchronisation point of the two calls above: waits for what they enqueued, makes slod_get_timings valid and returns
 * SLOD_ERR_NUMERIC (first offending patch in slod_last_error) if a patch of the last basis range reported a status
 * (A_ii or M not positive definite, eigen-solver not converged) -- the same mapping slod_compute_basis does. */
int slod_synchronize(slod_ctx *ctx);
int slod_ell_width(const slod_ctx *ctx, int64_t *width);

/* ---- one handle per GPU, NCCL inside the library (multi-process multi-GPU) ---------------------------------------*/
/* [begin, end) of the patch ids owned by `rank` of `world`: create_evenly_distributed_partitioning
 * (source/LOD.cc:116-118) -- contiguous blocks, the first n % world ranks hold one more. */
int slod_owned_range(const slod_ctx *ctx, int rank, int world, int64_t *patch_begin, int64_t *patch_end);
/* 128 bytes of an ncclUniqueId, to be created by one rank and handed to the others by the caller (MPI_Bcast,
 * torch.distributed.broadcast, a file, ...). */
int slod_comm_unique_id(void *id128);
/* Join this handle (its device) into a communicator of `world` ranks. */
int slod_comm_init(slod_ctx *ctx, int rank, int world, const void *id128);
/* The offline phase of this rank on device arrays of the full layout ([n_patches][spacedim][stride] and
 * [n_patches * spacedim][ell_width], as in the *_device calls): basis of the owned range, all-gather of A*phi (and of
 * phi when gather_phi != 0), coarse-matrix rows of the owned range, all-gather of the row blocks when gather_K != 0.
 * Everything is enqueued on `stream`; slod_synchronize() is the synchronisation point. */
int slod_offline_distributed(slod_ctx *ctx, double *d_phi, double *d_A_phi, double *d_K, int gather_phi, int gather_K,
                             void *stream);
/* Host destinations for the rank's OWN rows ([rows of slod_owned_range][spacedim][stride] for phi / A*phi,
 * [rows * spacedim][ell_width] for K; page-locked memory recommended; any pointer may be NULL).  When set,
 * slod_offline_distributed copies the rows on the handle's own stream as soon as the producing kernels are through --
 * phi while the A*phi all-gather and the coarse-matrix kernel still run -- and slod_synchronize waits for the copies. */
int slod_set_host_outputs(slod_ctx *ctx, double *h_phi_rows, double *h_A_phi_rows, double *h_K_rows);
/* convert a host copy of the block-ELL matrix into CSR (same contract as slod_get_coarse_csr) */
int slod_ell_to_csr(const slod_ctx *ctx, const double *h_K, int64_t *rowptr, int64_t *col, double *val,
                    int64_t *n_rows, int64_t *nnz);
/* number of kernels launched by this handle so far */
int slod_launch_count(const slod_ctx *ctx, int64_t *n);

#ifdef __cplusplus
}
#endif
#endif /* SLOD_B200_H */
