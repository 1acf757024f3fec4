// Selection of the super-localized basis function (source/LOD.cc:620-743):  d = -G^+ g with the reference's
// singular-value threshold (1e-15 sigma_0, :667) and its truncation loop (:703-725), then
// c = M^{-1}(e_d + sum_k d_k e_other[k]) (:727-743).
//
//   k_select_fast    LOD branch, and the SLOD fast path: when no singular value is thresholded and the truncation
//                    loop cannot fire, d solves the SPD system G d = -g (Cholesky in shared memory).  Everything
//                    else is appended to a work list for the eigen pipeline.
//   k_eig_tridiag    Householder tridiagonalisation G[o,o] = Q T Q^T in shared memory (one CTA per item),
//                    g_h = Q^T g.
//   k_eig_ql         implicit QL on T, one warp per item: eigenvalues, ghat = Z^T g_h on the fly, and the log of
//                    the plane rotations (Z is never formed).
//   k_eig_finish     d(r) = -Q Z W_r ghat for 32 candidate truncation counts r at once (lane = candidate): the
//                    rotation log replayed backwards on 32 vectors, the Householder reflectors applied, the first
//                    r with ||d(r)||_inf < 0.5 is the reference's answer.  Then c.
//   k_select_jacobi  cyclic Jacobi eigen-solver; only the fallback for items the QL pipeline gives up on.
// Included by kernels.cu.
#pragma once

namespace slod {

constexpr int kEigCand = 32;  // truncation counts tried per replay of the rotation log
constexpr int kEigLd = 36;    // row stride of the candidate block (== 4 mod 16: conflict-free 4 x 8 lane split)

// item = w * s + d  (w: index in the chunk's work list, d: component)
__device__ __forceinline__ int sym_idx(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }

// c = M^{-1} (e_d + sum_k z_k e_other[k])  (source/LOD.cc:727-743): eight rows of M^{-1} per warp pass (their loads are
// in flight together) and one packed reduction for the eight sums
__device__ __forceinline__ void minv_times_selection(const double *__restrict__ Minv, const double *z, int ncd, int n,
                                                     int d, double *__restrict__ cv, int warp, int nwarp, int lane) {
  for (int i0 = 0; i0 < ncd; i0 += 8 * nwarp) {
    double acc[8];
    const double *row[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      acc[u] = 0.0;
      row[u] = Minv + (size_t)min(i0 + warp + nwarp * u, ncd - 1) * ncd;   // rows past the end: loaded, not stored
    }
    for (int k = lane; k < n; k += 32) {
      const double zk = z[k];
      const int col = k + (k >= d);
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[u] += zk * row[u][col];
    }
    const double tot = warp_sum_packed(acc, lane);   // lane L holds row index warp_sum_index<8>(L)
    const int i = i0 + warp + nwarp * warp_sum_index<8>(lane);
    if ((lane & 3) == 0 && i < ncd) cv[i] = Minv[i * ncd + d] + tot;
  }
}

// ------------------------------------------------------------------------------------------------
// k_select_fast
// ------------------------------------------------------------------------------------------------
// 8 x 8 tiles of the packed lower triangle: tile (I, J), I >= J, at tri(I) + J; inside a tile element (r, c) sits at
// r * 8 + (c ^ 4 * ((r >> 1) & 1)): A-fragment, transposed-B-fragment and C-fragment accesses are bank-conflict free.
__device__ __forceinline__ int tile_off(int I, int J) { return (I * (I + 1) / 2 + J) * 64; }
// row I of entry `idx` of the packed lower triangle (idx = I (I + 1) / 2 + J, J <= I): closed form with one
// correction step instead of a linear search (the search took a quarter of k_select_fast's issue slots)
__device__ __forceinline__ int tri_row(int idx) {
  int I = (int)((sqrtf(8.0f * (float)idx + 1.0f) - 1.0f) * 0.5f);
  if ((I + 1) * (I + 2) / 2 <= idx) ++I;
  if (I * (I + 1) / 2 > idx) --I;
  return I;
}
__device__ __forceinline__ int tile_el(int r, int c) { return r * 8 + (c ^ (((r >> 1) & 1) << 2)); }

__global__ void __launch_bounds__(256, 3)
k_select_fast(const int *__restrict__ patch_ids, int n_work, const double *__restrict__ Minv_in,
              const double *__restrict__ G_in, double *__restrict__ cvec, double *__restrict__ diag,
              int *__restrict__ counters, int *__restrict__ work_list, int list_slot, SelectLayout lay) {
  extern __shared__ double smem[];
  const int nmax = lay.ncd_max;
  const int nbk_max = (nmax + 6) >> 3;                    // blocks of 8 covering n = ncd - 1 <= nmax - 1
  double *sL = smem;                                      // tiles of the lower triangle (Cholesky in place)
  double *sz = sL + (size_t)nbk_max * (nbk_max + 1) / 2 * 64;   // [8 nbk_max] right-hand side -> solution
  __shared__ int sFlag;
  __shared__ int sWork;
  __shared__ double sStat[4];
  const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5, NWARP = NT >> 5;
  const int g = lane >> 2, t = lane & 3;
  for (;;) {
    __syncthreads();
    if (tid == 0) sWork = atomicAdd(&counters[0], 1);
    __syncthreads();
    const int w = sWork;
    if (w >= n_work) break;
    const int pid = patch_ids[w];
    const Geom geo = make_geom(cP, pid);
    const int ncd = geo.Ncd, s = cP.s;
    const double *Minv = Minv_in + (size_t)w * lay.m_stride;
    const double *Gf = G_in + (size_t)w * lay.m_stride;
    for (int d = 0; d < s; ++d) {
      double *cv = cvec + ((size_t)w * s + d) * lay.ncd_max;
      double *dg = diag + ((size_t)pid * s + d) * 8;
      __syncthreads();
      if (!geo.slod) {  // LOD branch: c = M^{-1} e_d (source/LOD.cc:570-593)
        for (int i = tid; i < ncd; i += NT) cv[i] = Minv[i * ncd + d];
        if (tid == 0) { dg[0] = 0; dg[1] = 0; dg[2] = 0; dg[3] = 0; dg[5] = 0; dg[6] = 0; }
        continue;
      }
      const int n = ncd - 1;  // considered_candidates: all coarse dofs but d (source/LOD.cc:637-640)
      bool done = false;
      if (lay.fast_path) {
        // ---- fast path: if no singular value is thresholded and the truncation loop does not fire, the reference's
        // d = -G^+ g (source/LOD.cc:667-671) solves the SPD system G d = -g.  Blocked Cholesky (8 x 8 blocks, DMMA
        // panel and trailing updates on shared-memory tiles), then blocked triangular solves.  Any doubt (tiny pivot,
        // ||d||_inf close to or above 0.5) goes to the eigen pipeline. ----
        const int nbk = (n + 7) >> 3;
        double dmax = 0.0;
        for (int i = lane; i < n; i += 32) dmax = fmax(dmax, Gf[(i + (i >= d)) * ncd + (i + (i >= d))]);
        dmax = warp_max(dmax);  // every warp computes it (same value)
        for (int idx = tid; idx < nbk * (nbk + 1) / 2 * 64; idx += NT) {
          const int tile = idx >> 6, e = idx & 63, r = e >> 3, c = e & 7;
          const int I = tri_row(tile);
          const int J = tile - I * (I + 1) / 2;
          const int i = 8 * I + r, j = 8 * J + c;
          double v;
          if (i < n && j < n) v = Gf[(i + (i >= d)) * ncd + (j + (j >= d))];
          else v = (i == j) ? dmax : 0.0;   // padding: pivots equal to the largest diagonal entry
          sL[tile * 64 + tile_el(r, c)] = v;
        }
        for (int i = tid; i < 8 * nbk; i += NT) sz[i] = (i < n) ? -Gf[(i + (i >= d)) * ncd + d] : 0.0;
        if (tid == 0) { sFlag = 0; sStat[0] = 1e300; sStat[1] = 0.0; }
        __syncthreads();
        double pmin = 1e300;
        int bad = 0;
        for (int K = 0; K < nbk; ++K) {
          double *dK = sL + tile_off(K, K);
          if (warp == 0) {  // L_KK^{-1} replaces the diagonal tile (plain row-major)
            const double2 pv = make_double2(dK[tile_el(g, 2 * t)], dK[tile_el(g, 2 * t + 1)]);
            double pm;
            __syncwarp();
            bad |= chol8_inv(pv.x, pv.y, lane, dK, &pm);
            pmin = fmin(pmin, pm);
          }
          __syncthreads();
          // panel: L_IK = A_IK L_KK^{-T}
          for (int I = K + 1 + warp; I < nbk; I += NWARP) {
            double *tp = sL + tile_off(I, K);
            const double a0 = tp[tile_el(g, t)], a1 = tp[tile_el(g, 4 + t)];
            double p0 = 0.0, p1 = 0.0;
            dmma884(p0, p1, a0, dK[g * 8 + t]);        // B[k][n] = Linv[n][k]
            dmma884(p0, p1, a1, dK[g * 8 + 4 + t]);
            __syncwarp();
            tp[tile_el(g, 2 * t)] = p0;
            tp[tile_el(g, 2 * t + 1)] = p1;
          }
          __syncthreads();
          // trailing update: A_IJ -= L_IK L_JK^T for K < J <= I
          const int mrem = nbk - K - 1, ntile = mrem * (mrem + 1) / 2;
          for (int tt = warp; tt < ntile; tt += 2 * NWARP) {
            double c0[2], c1[2], a0[2], a1[2], b0[2], b1[2];
            double *ct[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              int tu = tt + u * NWARP;
              if (tu >= ntile) tu = tt;
              const int oi = tri_row(tu);
              const int oj = tu - oi * (oi + 1) / 2;
              const int I = K + 1 + oi, J = K + 1 + oj;
              ct[u] = sL + tile_off(I, J);
              const double *ti = sL + tile_off(I, K), *tj = sL + tile_off(J, K);
              c0[u] = ct[u][tile_el(g, 2 * t)];
              c1[u] = ct[u][tile_el(g, 2 * t + 1)];
              a0[u] = -ti[tile_el(g, t)];
              a1[u] = -ti[tile_el(g, 4 + t)];
              b0[u] = tj[tile_el(g, t)];       // B[k][n] = L_JK[n][k]
              b1[u] = tj[tile_el(g, 4 + t)];
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) dmma884(c0[u], c1[u], a0[u], b0[u]);
#pragma unroll
            for (int u = 0; u < 2; ++u) dmma884(c0[u], c1[u], a1[u], b1[u]);
#pragma unroll
            for (int u = 0; u < 2; ++u)
              if (tt + u * NWARP < ntile) {
                ct[u][tile_el(g, 2 * t)] = c0[u];
                ct[u][tile_el(g, 2 * t + 1)] = c1[u];
              }
          }
          __syncthreads();
        }
        if (warp == 0 && lane == 0) { sStat[0] = pmin; sStat[1] = bad; }
        // ---- L z = -g (forward), L^T x = z (backward), block by block ----
        for (int K = 0; K < nbk; ++K) {
          const double *dK = sL + tile_off(K, K);
          __syncthreads();
          double zr = 0.0;
          if (tid < 8) {
#pragma unroll
            for (int c = 0; c < 8; ++c) zr += dK[tid * 8 + c] * sz[8 * K + c];
          }
          __syncthreads();
          if (tid < 8) sz[8 * K + tid] = zr;
          __syncthreads();
          for (int idx = tid; idx < (nbk - K - 1) * 8; idx += NT) {
            const int I = K + 1 + (idx >> 3), r = idx & 7;
            const double *tp = sL + tile_off(I, K);
            double acc = 0.0;
#pragma unroll
            for (int c = 0; c < 8; ++c) acc += tp[tile_el(r, c)] * sz[8 * K + c];
            sz[8 * I + r] -= acc;
          }
        }
        for (int K = nbk - 1; K >= 0; --K) {
          const double *dK = sL + tile_off(K, K);
          __syncthreads();
          double xr = 0.0;
          if (tid < 8) {
#pragma unroll
            for (int r = 0; r < 8; ++r) xr += dK[r * 8 + tid] * sz[8 * K + r];   // L_KK^{-T}
          }
          __syncthreads();
          if (tid < 8) sz[8 * K + tid] = xr;
          __syncthreads();
          for (int idx = tid; idx < K * 8; idx += NT) {
            const int J = idx >> 3, c = idx & 7;
            const double *tp = sL + tile_off(K, J);
            double acc = 0.0;
#pragma unroll
            for (int r = 0; r < 8; ++r) acc += tp[tile_el(r, c)] * sz[8 * K + r];
            sz[8 * J + c] -= acc;
          }
        }
        __syncthreads();
        pmin = sStat[0];
        const bool spd_ok = (sStat[1] == 0.0) && (pmin > 1e-12 * dmax);
        if (spd_ok && warp == 0) {
          double m = 0.0;
          for (int r = lane; r < n; r += 32) m = fmax(m, fabs(sz[r]));
          m = warp_max(m);
          if (lane == 0 && m < 0.49) {
            sFlag = 1;
            dg[0] = m; dg[1] = 0; dg[2] = dmax; dg[3] = pmin; dg[5] = 1; dg[6] = 0;
          }
        }
        __syncthreads();
        if (sFlag) {
          minv_times_selection(Minv, sz, ncd, n, d, cv, warp, NWARP, lane);
          done = true;
        }
      }
      if (!done && tid == 0) work_list[atomicAdd(&counters[list_slot], 1)] = w * s + d;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// k_eig_tridiag : G[o,o] = Q T Q^T  (Householder, LAPACK dsytd2 'L' conventions)
// ------------------------------------------------------------------------------------------------
// Per item in HBM:  Hbuf [n][ldh]  row k = Householder vector v_k (v_k[0] = 1, length n-k-1)
//                   Vbuf [6][nmax] d | e | tau | g_h | (lambda) | (ghat)

template <int QN>   // 32 QN >= n: matrix columns per lane
__global__ void __launch_bounds__(512, 1)
k_eig_tridiag(const int *__restrict__ patch_ids, const int *__restrict__ counters, const int *__restrict__ eig_list,
              int round_off, const double *__restrict__ G_in, double *__restrict__ Hbuf, double *__restrict__ Vbuf,
              EigLayout lay) {
  // Two barriers per Householder step: every warp recomputes the (cheap) reflector and the vector w redundantly
  // and keeps v, w lane-distributed in registers (entry i in lane i % 32, slot i / 32), so the only data that has to
  // cross warps is p = tau A v (after the row-parallel symv) and the updated matrix itself.
  extern __shared__ double smem[];
  const int nmax = lay.nmax, ld = nmax | 1;
  double *sA = smem;                     // [nmax][ld] trailing matrix, both triangles kept up to date
  double *sv = sA + (size_t)nmax * ld;   // v  (identical copies written by every warp)
  double *sw = sv + nmax;                // w
  double *sgv = sw + nmax;               // g, becomes Q^T g (warp 0)
  double *sPv = sgv + nmax;              // p = tau A v
  const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5, NWARP = NT >> 5;
  int n_items = counters[1] - round_off;
  if (n_items > lay.cap_items) n_items = lay.cap_items;
  const int s = cP.s;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int it = eig_list[round_off + item];
    const int w = it / s, d = it - w * s;
    const Geom g = make_geom(cP, patch_ids[w]);
    const int ncd = g.Ncd, n = ncd - 1;
    const double *Gf = G_in + (size_t)w * lay.m_stride;
    double *H = Hbuf + (size_t)item * lay.h_stride;
    double *V = Vbuf + (size_t)item * lay.v_stride;
    __syncthreads();
    for (int i = warp; i < n; i += NWARP)
      for (int j = lane; j < n; j += 32) sA[i * ld + j] = Gf[(i + (i >= d)) * ncd + (j + (j >= d))];
    for (int i = tid; i < n; i += NT) sgv[i] = Gf[(i + (i >= d)) * ncd + d];
    __syncthreads();
    for (int k = 0; k < n - 1; ++k) {
      const int m = n - k - 1;  // x = A[k+1 .. n-1][k]
      // ---- reflector H = I - tau v v^T with H x = beta e_1 (every warp) ----
      double vq[QN];
      double sig = 0.0;
#pragma unroll
      for (int q = 0; q < QN; ++q) {
        const int i = lane + 32 * q;
        vq[q] = (i < m) ? sA[(k + 1 + i) * ld + k] : 0.0;
        if (i >= 1) sig += vq[q] * vq[q];
      }
      sig = warp_sum(sig);
      const double alpha = __shfl_sync(0xffffffffu, vq[0], 0);
      double tau = 0.0, beta = alpha, scal = 0.0;
      if (sig > 0.0) {
        beta = -copysign(sqrt(alpha * alpha + sig), alpha);
        tau = (beta - alpha) / beta;
        scal = 1.0 / (alpha - beta);
      }
#pragma unroll
      for (int q = 0; q < QN; ++q) {
        const int i = lane + 32 * q;
        vq[q] = (i < m) ? ((i == 0) ? 1.0 : vq[q] * scal) : 0.0;
        if (i < m) sv[i] = vq[q];
      }
      if (warp == 0) {
        double gd = 0.0;
#pragma unroll
        for (int q = 0; q < QN; ++q) {
          const int i = lane + 32 * q;
          if (i < m) {
            H[(size_t)k * lay.ldh + i] = vq[q];
            gd += vq[q] * sgv[k + 1 + i];
          }
        }
        gd = warp_sum(gd) * tau;  // g <- H g
#pragma unroll
        for (int q = 0; q < QN; ++q) {
          const int i = lane + 32 * q;
          if (i < m) sgv[k + 1 + i] -= gd * vq[q];
        }
        if (lane == 0) {
          V[k] = sA[k * ld + k];       // d_k
          V[nmax + k] = beta;          // e_k
          V[2 * nmax + k] = tau;
        }
      }
      if (tau != 0.0) {   // identical in every warp (same data, same arithmetic)
        // ---- p = tau A22 v : eight rows per warp pass (rows warp + NWARP u), lanes over the columns, one packed
        // reduction for the eight sums (9 shuffles instead of 40 on the latency path of every step) ----
        for (int i0 = 0; i0 < m; i0 += 8 * NWARP) {
          double acc[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            acc[u] = 0.0;
            const int i = i0 + warp + NWARP * u;
            if (i < m) {
              const double *row = sA + (k + 1 + i) * ld + (k + 1);
#pragma unroll
              for (int q = 0; q < QN; ++q) {
                const int j = lane + 32 * q;
                if (j < m) acc[u] += row[j] * vq[q];
              }
            }
          }
          const double tot = warp_sum_packed(acc, lane);   // lane L holds row index warp_sum_index<8>(L)
          const int i = i0 + warp + NWARP * warp_sum_index<8>(lane);
          if ((lane & 3) == 0 && i < m) sPv[i] = tau * tot;
        }
        __syncthreads();
        // ---- w = p - (tau/2)(p.v) v  (every warp), A22 <- A22 - v w^T - w v^T ----
        double wq[QN];
        double dot = 0.0;
#pragma unroll
        for (int q = 0; q < QN; ++q) {
          const int i = lane + 32 * q;
          wq[q] = (i < m) ? sPv[i] : 0.0;
          dot += wq[q] * vq[q];
        }
        dot = warp_sum(dot);
        const double al = -0.5 * tau * dot;
#pragma unroll
        for (int q = 0; q < QN; ++q) {
          const int i = lane + 32 * q;
          wq[q] += al * vq[q];
          if (i < m) sw[i] = wq[q];
        }
        __syncwarp();
        for (int i = warp; i < m; i += NWARP) {
          const double vi = sv[i], wi = sw[i];
          double *row = sA + (k + 1 + i) * ld + (k + 1);
#pragma unroll
          for (int q = 0; q < QN; ++q) {
            const int j = lane + 32 * q;
            if (j < m) row[j] -= vi * wq[q] + wi * vq[q];
          }
        }
      }
      __syncthreads();
    }
    if (tid == 0) {
      V[n - 1] = sA[(n - 1) * ld + (n - 1)];
      V[nmax + n - 1] = 0.0;
    }
    for (int i = tid; i < n; i += NT) V[3 * nmax + i] = sgv[i];
  }
}

// ------------------------------------------------------------------------------------------------
// k_eig_ql : implicit QL (EISPACK tql2 without the eigenvector matrix); one warp per item, lane 0 iterates.
// Every plane rotation is (i) applied to g_h on the fly (ghat = Z^T g_h) and (ii) appended to the log.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_eig_ql(const int *__restrict__ patch_ids, const int *__restrict__ counters, const int *__restrict__ eig_list,
         int round_off, double *__restrict__ Vbuf, double2 *__restrict__ rot_cs, unsigned short *__restrict__ rot_i,
         int *__restrict__ rot_n, EigLayout lay) {
  extern __shared__ double smem[];
  const int nmax = lay.nmax;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, NWARP = blockDim.x >> 5;
  double *sd = smem + (size_t)warp * 3 * nmax, *se = sd + nmax, *sgh = se + nmax;
  int n_items = counters[1] - round_off;
  if (n_items > lay.cap_items) n_items = lay.cap_items;
  const int s = cP.s;
  for (int item = blockIdx.x * NWARP + warp; item < n_items; item += gridDim.x * NWARP) {
    const int it = eig_list[round_off + item];
    const int w = it / s;
    const Geom g = make_geom(cP, patch_ids[w]);
    const int n = g.Ncd - 1;
    double *V = Vbuf + (size_t)item * lay.v_stride;
    __syncwarp();
    for (int i = lane; i < n; i += 32) { sd[i] = V[i]; se[i] = V[nmax + i]; sgh[i] = V[3 * nmax + i]; }
    __syncwarp();
    // Every lane runs the (identical) scalar recurrence -- same shared-memory addresses, same values, so control flow
    // stays warp uniform and the fp64 pipe sees the same instruction count as with one active lane -- which lets the
    // warp scan for the next negligible off-diagonal 32 entries at a time.  Lane 0 alone writes the rotation log.
    {
      double2 *cs = rot_cs + (size_t)item * lay.log_cap;
      unsigned short *ri = rot_i + (size_t)item * lay.log_cap;
      long long nrot = 0;
      int total_iter = 0;
      bool fail = false;
      for (int l = 0; l < n && !fail; ++l) {
        int iter = 0, m;
        do {
          m = n - 1;
          for (int m0 = l; m0 < n - 1; m0 += 32) {
            const int mm = m0 + lane;
            bool small = false;
            if (mm < n - 1) small = fabs(se[mm]) <= 2.220446049250313e-16 * (fabs(sd[mm]) + fabs(sd[mm + 1]));
            const unsigned hit = __ballot_sync(0xffffffffu, small);
            if (hit) { m = m0 + __ffs(hit) - 1; break; }
          }
          if (m != l) {
            if (++iter > 60 || nrot + (m - l) > lay.log_cap) { fail = true; break; }
            ++total_iter;
            const double d_l = sd[l];   // kept in a register: every lane executes this code, no read-modify-write on shared memory
            double gg = (sd[l + 1] - d_l) / (2.0 * se[l]);
            double r = sqrt(gg * gg + 1.0);
            gg = sd[m] - d_l + se[l] / (gg + copysign(r, gg));
            double sn = 1.0, c = 1.0, p = 0.0;
            int i = m - 1;
            // operands of rotation i travel in registers: the next ones are fetched one rotation ahead, and what
            // rotation i + 1 produced for position i + 1 is handed over instead of going through shared memory.
            // Running pointers instead of indices: the loop is bound by its instruction count (one warp, dependent
            // issue), not by any pipe.
            double *pe = se + i, *pd = sd + i, *pg = sgh + i;
            double2 *pcs = cs + nrot;
            unsigned short *pri = ri + nrot;
            double e_i = pe[0], d_lo = pd[0], d_hi = pd[1], gh_lo = pg[0], gh_hi = pg[1];
            for (; i >= l; --i, --pe, --pd, --pg) {
              double e_n = 0.0, d_n = 0.0, gh_n = 0.0;
              if (i > l) { e_n = pe[-1]; d_n = pd[-1]; gh_n = pg[-1]; }
              const double f = sn * e_i;
              const double b = c * e_i;
              const double h2 = f * f + gg * gg;
              const double ir = rsqrt(h2);   // one MUFU-based op on the dependency chain instead of sqrt + divide
              r = (h2 > 0.0) ? h2 * ir : 0.0;
              pe[1] = r;
              if (r == 0.0) {
                pd[1] = d_hi - p;
                se[m] = 0.0;
                break;
              }
              sn = f * ir;
              c = gg * ir;
              gg = d_hi - p;
              r = (d_lo - gg) * sn + 2.0 * c * b;
              p = sn * r;
              pd[1] = gg + p;
              gg = c * r - b;
              // Z <- Z R : columns (i, i+1); here applied as x <- R^T x to g_h
              pg[1] = sn * gh_lo + c * gh_hi;
              gh_hi = c * gh_lo - sn * gh_hi;
              if (lane == 0) {
                *pcs = make_double2(c, sn);
                *pri = (unsigned short)i;
              }
              ++pcs;
              ++pri;
              ++nrot;
              e_i = e_n; d_hi = d_lo; d_lo = d_n; gh_lo = gh_n;
            }
            sgh[i + 1] = gh_hi;   // position l after a full sweep, position i + 1 after an early exit
            if (r == 0.0 && i >= l) continue;
            sd[l] = d_l - p;
            se[l] = gg;
            se[m] = 0.0;
          }
        } while (m != l);
      }
      if (lane == 0) {
        rot_n[2 * item] = fail ? -1 : (int)nrot;
        rot_n[2 * item + 1] = total_iter;
      }
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) { V[4 * nmax + i] = sd[i]; V[5 * nmax + i] = sgh[i]; }
  }
}

// ------------------------------------------------------------------------------------------------
// k_eig_finish
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_eig_finish(const int *__restrict__ patch_ids, int *__restrict__ counters, const int *__restrict__ eig_list,
             int round_off, const double *__restrict__ Minv_in, const double *__restrict__ Hbuf,
             const double *__restrict__ Vbuf, const double2 *__restrict__ rot_cs,
             const unsigned short *__restrict__ rot_i, const int *__restrict__ rot_n, double *__restrict__ cvec,
             double *__restrict__ diag, int *__restrict__ jac_list, EigLayout lay) {
  extern __shared__ double smem[];
  const int nmax = lay.nmax;
  double *sY = smem;                            // [nmax][32]  candidate vectors, lane = candidate
  double *slam = sY + (size_t)nmax * kEigLd;  // eigenvalues
  double *sgh = slam + nmax;                    // ghat
  double *sd = sgh + nmax;                      // selected d
  double *sVr = sd + nmax;                      // [2][nmax] staged Householder vectors
  int *srank = (int *)(sVr + 2 * nmax);         // rank of eigenvalue k in descending |lambda| order
  __shared__ double sNorm[kEigCand];
  __shared__ double sStat[4];
  __shared__ int sPick;
  const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5;
  int n_items = counters[1] - round_off;
  if (n_items > lay.cap_items) n_items = lay.cap_items;
  const int s = cP.s;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int it = eig_list[round_off + item];
    const int w = it / s, d = it - w * s;
    const int pid = patch_ids[w];
    const Geom g = make_geom(cP, pid);
    const int ncd = g.Ncd, n = ncd - 1;
    const double *Minv = Minv_in + (size_t)w * lay.m_stride;
    const double *H = Hbuf + (size_t)item * lay.h_stride;
    const double *V = Vbuf + (size_t)item * lay.v_stride;
    const double2 *cs = rot_cs + (size_t)item * lay.log_cap;
    const unsigned short *ri = rot_i + (size_t)item * lay.log_cap;
    const int nrot = rot_n[2 * item];
    double *cv = cvec + (size_t)it * lay.ncd_max;
    double *dg = diag + ((size_t)pid * s + d) * 8;
    __syncthreads();
    if (nrot < 0) {  // QL gave up: Jacobi fallback
      if (tid == 0) jac_list[atomicAdd(&counters[2], 1)] = it;
      continue;
    }
    for (int i = tid; i < n; i += NT) { slam[i] = V[4 * nmax + i]; sgh[i] = V[5 * nmax + i]; }
    __syncthreads();
    for (int i = tid; i < n; i += NT) {
      const double li = fabs(slam[i]);
      int rank = 0;
      for (int j = 0; j < n; ++j) {
        const double lj = fabs(slam[j]);
        rank += (lj > li) || (lj == li && j < i);
      }
      srank[i] = rank;
      if (rank == 0) sStat[0] = li;  // sigma_0
    }
    if (tid == 0) sPick = -1;
    __syncthreads();
    const double sig0 = sStat[0], thr = 1e-15 * sig0;
    int pick = -1;
    for (int batch = 0; batch * kEigCand <= n && pick < 0; ++batch) {
      // candidate r = number of truncation steps: the r smallest singular values are dropped (source/LOD.cc:703-725)
      for (int idx = tid; idx < n * kEigCand; idx += NT) {
        const int k = idx >> 5, r = batch * kEigCand + (idx & 31);
        const double lk = slam[k];
        sY[k * kEigLd + (idx & 31)] = (srank[k] < n - r && fabs(lk) > thr) ? -sgh[k] / lk : 0.0;
      }
      __syncthreads();
      // ---- y <- Z y : replay the rotation log backwards, lane = candidate ----
      if (warp == 0) {
        int cur = -1;
        double carry = 0.0;
        // chunks of 32 rotations, one per lane, the next chunk is in flight while the current one is replayed
        double2 rcn = make_double2(1.0, 0.0);
        int idxn = 0;
        if (nrot - 1 - lane >= 0) { rcn = cs[nrot - 1 - lane]; idxn = ri[nrot - 1 - lane]; }
        for (int q0 = nrot; q0 > 0; q0 -= 32) {
          const double2 rc = rcn;
          const int idx = idxn;
          const int qn = q0 - 33 - lane;
          if (qn >= 0) { rcn = cs[qn]; idxn = ri[qn]; }
          const int cnt = min(32, q0);
          for (int j = 0; j < cnt; ++j) {
            const double c = __shfl_sync(0xffffffffu, rc.x, j), sn = __shfl_sync(0xffffffffu, rc.y, j);
            const int i = __shfl_sync(0xffffffffu, idx, j);
            double a;
            if (i != cur) {
              if (cur >= 0) sY[cur * kEigLd + lane] = carry;
              a = sY[i * kEigLd + lane];
            } else {
              a = carry;
            }
            const double b = sY[(i + 1) * kEigLd + lane];
            sY[i * kEigLd + lane] = c * a + sn * b;
            carry = c * b - sn * a;
            cur = i + 1;
          }
        }
        if (cur >= 0) sY[cur * kEigLd + lane] = carry;
      }
      __syncthreads();
      // ---- y <- Q y = H_0 H_1 ... H_{n-2} y ; warp handles 8 candidates, 4 lanes split the rows ----
      {
        const int cand = 8 * warp + (lane >> 2), part = lane & 3;
        int buf = 0;
        // stage v_{n-2}
        if (n >= 2) {
          const int k = n - 2, m = n - k - 1;
          for (int i = tid; i < m; i += NT) sVr[i] = H[(size_t)k * lay.ldh + i];
        }
        __syncthreads();
        for (int k = n - 2; k >= 0; --k) {
          const int m = n - k - 1;
          const double *vk = sVr + buf * nmax;
          if (k > 0) {  // prefetch v_{k-1} into the other buffer
            double *vn = sVr + (buf ^ 1) * nmax;
            for (int i = tid; i < m + 1; i += NT) __pipeline_memcpy_async(vn + i, H + (size_t)(k - 1) * lay.ldh + i, 8);
          }
          __pipeline_commit();
          const double tau = V[2 * nmax + k];
          double dot = 0.0;
          for (int i = part; i < m; i += 4) dot += vk[i] * sY[(k + 1 + i) * kEigLd + cand];
          dot += __shfl_xor_sync(0xffffffffu, dot, 1);
          dot += __shfl_xor_sync(0xffffffffu, dot, 2);
          dot *= tau;
          for (int i = part; i < m; i += 4) sY[(k + 1 + i) * kEigLd + cand] -= dot * vk[i];
          __pipeline_wait_prior(0);
          __syncthreads();
          buf ^= 1;
        }
      }
      // ---- ||d(r)||_inf per candidate ----
      {
        const int cand = 8 * warp + (lane >> 2), part = lane & 3;
        double mx = 0.0;
        for (int i = part; i < n; i += 4) mx = fmax(mx, fabs(sY[i * kEigLd + cand]));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        if (part == 0) sNorm[cand] = mx;
      }
      __syncthreads();
      if (tid == 0) {
        if (batch == 0) sStat[1] = sNorm[0];  // ||d||_inf before truncation
        for (int c = 0; c < kEigCand; ++c) {
          const int r = batch * kEigCand + c;
          if (r > n) break;
          if (sNorm[c] < 0.5) { sPick = r; break; }
        }
      }
      __syncthreads();
      pick = sPick;
      if (pick >= 0) {
        const int c = pick - batch * kEigCand;
        for (int i = tid; i < n; i += NT) sd[i] = sY[i * kEigLd + c];
      }
      __syncthreads();
    }
    if (pick < 0) {  // cannot happen (r = n gives d = 0); keep the fallback anyway
      if (tid == 0) jac_list[atomicAdd(&counters[2], 1)] = it;
      continue;
    }
    minv_times_selection(Minv, sd, ncd, n, d, cv, warp, NT >> 5, lane);
    if (tid == 0) {
      // smallest singular value still in use
      double kept = sig0;
      for (int k = 0; k < n; ++k) {
        const double lk = fabs(slam[k]);
        if (srank[k] < n - pick && lk > thr) kept = fmin(kept, lk);
      }
      dg[0] = sStat[1]; dg[1] = pick; dg[2] = sig0; dg[3] = kept; dg[5] = 3; dg[6] = rot_n[2 * item + 1];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// k_select_jacobi : thresholded pseudo-inverse through a cyclic Jacobi eigen-solver (fallback)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512, 1)
k_select_jacobi(const int *__restrict__ patch_ids, const int *__restrict__ counters, const int *__restrict__ jac_list,
                const double *__restrict__ Minv_in, const double *__restrict__ G_in, double *__restrict__ cvec,
                double *__restrict__ diag, int *__restrict__ status, SelectLayout lay) {
  extern __shared__ double smem[];
  const int nmax = lay.ncd_max;       // >= n + 1
  double *sG = smem;                                   // packed lower, n(n+1)/2
  double *sV = sG + (size_t)nmax * (nmax + 1) / 2;     // [n][n]
  double *sg = sV + (size_t)nmax * nmax;               // g
  double *sgh = sg + nmax;                             // V^T g
  double *sd = sgh + nmax;                             // d
  double *slam = sd + nmax;                            // eigenvalues
  double *sc = slam + nmax;                            // rotation cos per pair
  double *ss = sc + nmax;                              // rotation sin per pair
  int *sp = (int *)(ss + nmax);                        // pair p
  int *sq = sp + nmax;                                 // pair q
  int *sord = sq + nmax;                               // eigenvalue order (descending |lambda|)
  __shared__ double sRed[32];
  __shared__ double sOff;
  const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = NT >> 5;
  const int n_items = counters[2];
  const int s = cP.s;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int it = jac_list[item];
    const int w = it / s, d = it - w * s;
    const int pid = patch_ids[w];
    const Geom g = make_geom(cP, pid);
    const int ncd = g.Ncd;
    const double *Minv = Minv_in + (size_t)w * lay.m_stride;
    const double *Gf = G_in + (size_t)w * lay.m_stride;
    double *cv = cvec + (size_t)it * lay.ncd_max;
    double *dg = diag + ((size_t)pid * s + d) * 8;
    __syncthreads();
    const int n = ncd - 1;               // considered_candidates
    const int np = (n + 1) & ~1;         // even player count for the tournament
    const int half = np / 2;
    for (int idx = tid; idx < n * n; idx += NT) {
      const int i = idx / n, j = idx % n;
      if (j <= i) sG[i * (i + 1) / 2 + j] = Gf[(i + (i >= d)) * ncd + (j + (j >= d))];
      sV[idx] = (i == j) ? 1.0 : 0.0;
    }
    for (int i = tid; i < n; i += NT) sg[i] = Gf[(i + (i >= d)) * ncd + d];
    __syncthreads();
    double maxdiag = 0.0;
    for (int i = 0; i < n; ++i) maxdiag = fmax(maxdiag, fabs(sG[i * (i + 1) / 2 + i]));
    const double tol = 1e-17 * maxdiag;
    int sweeps = 0;
    for (; sweeps < 40; ++sweeps) {
      if (tid == 0) sOff = 0.0;
      __syncthreads();
      double myoff = 0.0;
      for (int step = 0; step < np - 1; ++step) {
        if (tid < half) {
          int a, b;
          if (tid == 0) { a = step; b = np - 1; }
          else { a = (step + tid) % (np - 1); b = (step - tid + (np - 1)) % (np - 1); }
          int p = min(a, b), q = max(a, b);
          double c = 1.0, sn = 0.0;
          if (q < n) {
            const double apq = sG[q * (q + 1) / 2 + p];
            myoff = fmax(myoff, fabs(apq));
            if (fabs(apq) > tol) {
              const double app = sG[p * (p + 1) / 2 + p], aqq = sG[q * (q + 1) / 2 + q];
              const double tau = (aqq - app) / (2.0 * apq);
              const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
              c = 1.0 / sqrt(1.0 + t * t);
              sn = t * c;
            }
          } else {
            q = -1;
          }
          sp[tid] = p; sq[tid] = q; sc[tid] = c; ss[tid] = sn;
        }
        __syncthreads();
        for (int idx = tid; idx < half * half; idx += NT) {
          const int k1 = idx / half, k2 = idx % half;
          if (k2 > k1) continue;
          const int p1 = sp[k1], q1 = sq[k1], p2 = sp[k2], q2 = sq[k2];
          const double c1 = sc[k1], s1 = ss[k1], c2 = sc[k2], s2 = ss[k2];
          if (q1 < 0 && q2 < 0) continue;
          if (k1 == k2) {
            if (q1 < 0) continue;
            const double app = sG[sym_idx(p1, p1)], aqq = sG[sym_idx(q1, q1)], apq = sG[sym_idx(q1, p1)];
            const double npp = c1 * c1 * app - 2.0 * c1 * s1 * apq + s1 * s1 * aqq;
            const double nqq = s1 * s1 * app + 2.0 * c1 * s1 * apq + c1 * c1 * aqq;
            const double npq = (c1 * c1 - s1 * s1) * apq + c1 * s1 * (app - aqq);
            sG[sym_idx(p1, p1)] = npp;
            sG[sym_idx(q1, q1)] = nqq;
            sG[sym_idx(q1, p1)] = (s1 != 0.0) ? 0.0 : npq;
            continue;
          }
          double b00 = sG[sym_idx(p1, p2)];
          double b01 = (q2 >= 0) ? sG[sym_idx(p1, q2)] : 0.0;
          double b10 = (q1 >= 0) ? sG[sym_idx(q1, p2)] : 0.0;
          double b11 = (q1 >= 0 && q2 >= 0) ? sG[sym_idx(q1, q2)] : 0.0;
          const double t00 = c1 * b00 - s1 * b10, t01 = c1 * b01 - s1 * b11;
          const double t10 = s1 * b00 + c1 * b10, t11 = s1 * b01 + c1 * b11;
          b00 = c2 * t00 - s2 * t01; b01 = s2 * t00 + c2 * t01;
          b10 = c2 * t10 - s2 * t11; b11 = s2 * t10 + c2 * t11;
          sG[sym_idx(p1, p2)] = b00;
          if (q2 >= 0) sG[sym_idx(p1, q2)] = b01;
          if (q1 >= 0) sG[sym_idx(q1, p2)] = b10;
          if (q1 >= 0 && q2 >= 0) sG[sym_idx(q1, q2)] = b11;
        }
        for (int idx = tid; idx < n * half; idx += NT) {
          const int r = idx / half, k = idx % half;
          const int q = sq[k];
          if (q < 0) continue;
          const int p = sp[k];
          const double c = sc[k], sn = ss[k];
          const double vp = sV[r * n + p], vq = sV[r * n + q];
          sV[r * n + p] = c * vp - sn * vq;
          sV[r * n + q] = sn * vp + c * vq;
        }
        __syncthreads();
      }
      myoff = warp_max(myoff);
      if (lane == 0) sRed[warp] = myoff;
      __syncthreads();
      if (tid == 0) {
        double m = 0.0;
        for (int i = 0; i < nwarp; ++i) m = fmax(m, sRed[i]);
        sOff = m;
      }
      __syncthreads();
      if (sOff <= tol) { ++sweeps; break; }
    }
    for (int i = tid; i < n; i += NT) {
      slam[i] = sG[i * (i + 1) / 2 + i];
      sord[i] = i;
    }
    __syncthreads();
    for (int i = tid; i < n; i += NT) {
      const double li = fabs(slam[i]);
      int rank = 0;
      for (int j = 0; j < n; ++j) {
        const double lj = fabs(slam[j]);
        rank += (lj > li) || (lj == li && j < i);
      }
      if (li == li) sord[rank] = i;
      double acc = 0.0;
      for (int r = 0; r < n; ++r) acc += sV[r * n + i] * sg[r];
      sgh[i] = acc;
    }
    __syncthreads();
    const double sig0 = fabs(slam[sord[0]]);
    for (int r = tid; r < n; r += NT) {
      double acc = 0.0;
      for (int k = 0; k < n; ++k) {
        const double lk = slam[k];
        if (fabs(lk) > 1e-15 * sig0) acc += sV[r * n + k] * (sgh[k] / lk);
      }
      sd[r] = -acc;
    }
    __syncthreads();
    if (warp == 0) {
      int steps = 0;
      double dinf0 = -1.0, kept = fabs(slam[sord[n - 1]]);
      int i = n - 1;
      for (; i >= 0; --i) {
        double m = 0.0;
        for (int r = lane; r < n; r += 32) m = fmax(m, fabs(sd[r]));
        m = warp_max(m);
        if (dinf0 < 0.0) dinf0 = m;
        if (m < 0.5) break;
        const int k = sord[i];
        const double lk = slam[k];
        if (fabs(lk) > 1e-15 * sig0) {
          const double f = sgh[k] / lk;
          for (int r = lane; r < n; r += 32) sd[r] += sV[r * n + k] * f;
        }
        __syncwarp();
        ++steps;
      }
      {
        int last = n - 1 - steps;
        if (last < 0) last = 0;
        while (last > 0 && !(fabs(slam[sord[last]]) > 1e-15 * sig0)) --last;
        kept = fabs(slam[sord[last]]);
      }
      if (lane == 0) {
        dg[0] = dinf0; dg[1] = steps; dg[2] = sig0; dg[3] = kept; dg[5] = 2; dg[6] = sweeps;
        if (sweeps >= 40) atomicOr(&status[pid], 4);
      }
    }
    __syncthreads();
    for (int i = tid; i < ncd; i += NT) {
      double acc = Minv[i * ncd + d];
      for (int k = 0; k < n; ++k) acc += sd[k] * Minv[i * ncd + (k + (k >= d))];
      cv[i] = acc;
    }
  }
}

}  // namespace slod
