// Launch interface between the host orchestration (capi.cu) and the kernels (kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include "geom.h"

namespace slod {

constexpr int kSolveNB = 8;  // panel width of the blocked banded Cholesky

struct SolveLayout {
  int threads;
  int R;             // window rows = bw_max + NB
  int ldw;           // band row length = bw_max + 1
  int ldr;           // RHS window leading dimension (>= NcdMax)
  int bw_max;
  int coef_doubles;  // shared doubles reserved for the patch coefficients
  int ldx;           // leading dimension of X rows (NcdMax)
  long long x_stride;     // doubles per patch in Xbuf (NiMax * ldx)
  long long lws_per_cta;  // doubles of L workspace per CTA
};
struct DenseLayout {
  int threads;
  int ncd_max, nb_max;
  int coef_doubles;
  int ldx;
  long long x_stride;
  long long m_stride;  // ncd_max^2
};
struct SelectLayout {
  int threads;
  int fast_path;  // try the Cholesky solve of G d = -g before the eigen-solver
  int ncd_max;
  long long m_stride;
};
struct FinishLayout {
  int coef_doubles;
  int nf_max, ncd_max;
  int ldx;
  long long x_stride;
};

cudaError_t upload_params(const Params &p);

cudaError_t launch_patch_solve(int grid, size_t smem, cudaStream_t st, const int *ids, int n_work, const double *coef,
                               double *X, double *Lws, int *status, const SolveLayout &lay);
// tensor-core (mma.sync f64) solver; variant 0: <RBMAX 13, 16 warps> (3-D), 1: <4, 4 warps>, 2: <4, 8 warps> (2-D)
size_t solve_mma_smem(int variant, int coef_doubles, int nip_max, int stw);
cudaError_t launch_patch_solve_mma(int variant, int grid, size_t smem, cudaStream_t st, const int *ids, int n_work,
                                   const double *coef, double *X, double *Lws, int *status, int coef_doubles, int ldx,
                                   long long x_stride, long long lws_per_cta, int nip_max, int stw);
cudaError_t launch_patch_dense(int grid, size_t smem, cudaStream_t st, const int *ids, int n_work, const double *coef,
                               const double *X, double *Minv, double *G, double *diag, int *status,
                               const DenseLayout &lay);
// tensor-core dense stage, ntile in {4, 8, 16} (8*ntile >= coarse dofs per patch)
size_t dense_mma_smem(int ntile, int coef_doubles, int nb_max);
cudaError_t launch_patch_dense_mma(int ntile, int grid, size_t smem, cudaStream_t st, const int *ids, int n_work,
                                   const double *coef, const double *X, double *Minv, double *G, double *diag,
                                   int *status, const DenseLayout &lay);
cudaError_t launch_patch_select(int grid, size_t smem, cudaStream_t st, const int *ids, int n_work, const double *Minv,
                                const double *G, double *cvec, double *diag, int *status, int *work_counter,
                                const SelectLayout &lay);
cudaError_t launch_patch_finish(int grid, size_t smem, cudaStream_t st, const int *ids, int n_work, const double *coef,
                                const double *X, const double *cvec, double *phi, double *aphi,
                                const FinishLayout &lay);
cudaError_t launch_coarse(int grid, size_t smem, cudaStream_t st, int p0, int p1, const double *phi, const double *aphi,
                          double *Kell, const FinishLayout &lay);

}  // namespace slod
