// k_patch_solve_mma : X = A_ii^{-1} P_i for one patch per CTA with FP64 tensor-core (mma.sync m8n8k4) updates.
//
// Blocked (8 columns per step) banded Cholesky of A_ii on a circular (RB x RB blocks of 8x8) dense window
// in shared memory.  The multi-RHS window lives in REGISTERS as mma accumulator tiles: warp w owns the 8
// right-hand-side columns [8w, 8w+8) for every 8-row block of the window, so the forward substitution costs no
// shared-memory traffic for C.  The backward substitution keeps the X window in registers in mma B-fragment
// layout.  The 8x8 diagonal factorisation of step k+1 is done by warp 0 with shuffles while the other warps run
// the trailing update of step k (look-ahead), so there are two block barriers per step.
// Included by kernels.cu (needs cP and the helpers defined there).
#pragma once

namespace slod {

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// C-fragment (row g = lane>>2, cols 2t,2t+1 with t = lane&3) -> B-fragment of k-step j (row 4j+t, col g)
__device__ __forceinline__ double c_to_b(double c0, double c1, int lane, int j) {
  const int src = 16 * j + 4 * (lane & 3) + (lane >> 3);
  const double v0 = __shfl_sync(0xffffffffu, c0, src);
  const double v1 = __shfl_sync(0xffffffffu, c1, src);
  return ((lane >> 2) & 1) ? v1 : v0;
}

// In-register Cholesky of an 8x8 SPD tile held in C layout by one warp; writes L^{-1} (lower, row-major 8x8)
// to sLinv and uses sLd as scratch.  Returns nonzero if a pivot was not positive.
__device__ __forceinline__ int chol8_inv(double v0, double v1, int lane, double *sLd, double *sLinv) {
  const int g = lane >> 2, t = lane & 3;
  int bad = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const double sel = (k & 1) ? v1 : v0;
    const double dkk = __shfl_sync(0xffffffffu, sel, 4 * k + (k >> 1));
    if (!(dkk > 0.0)) bad = 1;
    const double inv = 1.0 / sqrt(dkk);
    const double lik = __shfl_sync(0xffffffffu, sel, 4 * g + (k >> 1)) * inv;
    const double lc0 = __shfl_sync(0xffffffffu, sel, 4 * (2 * t) + (k >> 1)) * inv;
    const double lc1 = __shfl_sync(0xffffffffu, sel, 4 * (2 * t + 1) + (k >> 1)) * inv;
    if (2 * t > k) v0 -= lik * lc0;
    else if (2 * t == k) v0 = lik;
    if (2 * t + 1 > k) v1 -= lik * lc1;
    else if (2 * t + 1 == k) v1 = lik;
  }
  sLd[g * 8 + 2 * t] = (2 * t <= g) ? v0 : 0.0;
  sLd[g * 8 + 2 * t + 1] = (2 * t + 1 <= g) ? v1 : 0.0;
  sLinv[g * 8 + 2 * t] = 0.0;
  sLinv[g * 8 + 2 * t + 1] = 0.0;
  __syncwarp();
  if (lane < 8) {  // column `lane` of L^{-1} by forward substitution
    const int j = lane;
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      double sum = (i == j) ? 1.0 : 0.0;
#pragma unroll
      for (int tt = 0; tt < 8; ++tt)
        if (tt < i) sum -= sLd[i * 8 + tt] * ((tt >= j) ? x[tt] : 0.0);
      x[i] = (i >= j) ? sum / sLd[i * 8 + i] : 0.0;
      if (i >= j) sLinv[i * 8 + j] = x[i];
    }
  }
  __syncwarp();
  return bad;
}

struct SolveMmaLayout {
  int coef_doubles;
  int ldx;                // leading dimension of X rows (8 * NW)
  long long x_stride;     // doubles per patch in Xbuf (NiPmax * ldx)
  long long lws_per_cta;  // doubles of L workspace per CTA  (steps_max * (64 + 8*(RBMAX-1)*8))
};

template <int RBMAX, int NW>
__global__ void __launch_bounds__(32 * NW, 1)
k_patch_solve_mma(const int *__restrict__ patch_ids, int n_work, const double *__restrict__ d_coef,
                  double *__restrict__ Xbuf, double *__restrict__ Lws, int *__restrict__ status, SolveMmaLayout lay) {
  constexpr int R = 8 * RBMAX;
  constexpr int LDWF = (R % 16 == 8) ? R : R + 8;  // row stride with LDWF % 16 == 8 : conflict-free C fragments
  constexpr int LDP = R + 4;                        // k-major panel copy, LDP % 16 in {4, 12}
  constexpr int NT = 32 * NW;
  constexpr int LSTEP = 64 + 8 * (RBMAX - 1) * 8;  // doubles per step in the L workspace: Linv + panel rows
  extern __shared__ double smem[];
  double *sCoef = smem;
  double *sWf = sCoef + lay.coef_doubles;   // [R][LDWF] circular dense window of the trailing matrix
  double *sLpT = sWf + R * LDWF;            // [8][LDP]  panel, k-major, slot rows
  double *sLpR = sLpT + 8 * LDP;            // [2][R*8]  panel rows (row-major by offset) for the backward pass
  double *sLd = sLpR + 2 * R * 8;           // [64] scratch
  double *sLinv = sLd + 64;                 // [2][64]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  double *myL = Lws + (size_t)blockIdx.x * lay.lws_per_cta;

  for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
    const int pid = patch_ids[w];
    const Geom geo = make_geom(cP, pid);
    const int Ni = geo.Ni, bw = geo.bw, ncd = geo.Ncd;
    const int NBLK = (Ni + 7) >> 3;
    int RB = (bw + 8 + 7) >> 3;
    if (RB > RBMAX) RB = RBMAX;  // host guarantees bw_max fits
    double *X = Xbuf + (size_t)w * lay.x_stride;
    __syncthreads();
    load_coef(geo, d_coef, sCoef);
    __syncthreads();

    // value of A_ii[r][c] (c <= r inside the band, else 0); rows >= Ni are identity rows
    auto a_entry = [&](int r, int c) -> double {
      if (r >= Ni || c >= Ni) return (r == c) ? 1.0 : 0.0;
      if (c > r || r - c > bw) return 0.0;
      int a[3], b[3], ca, cb;
      idof_to_node(geo, r, a, ca);
      idof_to_node(geo, c, b, cb);
      int dl[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
      if (dl[0] < -1 || dl[0] > 1 || dl[1] < -1 || dl[1] > 1 || dl[2] < -1 || dl[2] > 1) return 0.0;
      return stiff_entry(cP, geo, sCoef, a, dl, ca, cb);
    };
    // write the 8 rows of block `blk` (all live column slots) into its slot of the window
    auto assemble_block = [&](int blk) {
      const int slot = blk % RB;
      const int first_col_blk = blk - (RB - 1);  // oldest block that can be live together with blk
      for (int idx = tid; idx < 8 * 8 * RB; idx += NT) {
        const int i = idx / (8 * RB), cs = idx % (8 * RB);  // row in block, column slot index
        const int csb = cs >> 3;                            // column slot block
        // absolute block of this column slot: the one in [first_col_blk, blk] congruent to csb mod RB
        int cb_abs = first_col_blk + ((csb - (first_col_blk % RB + RB) % RB + RB) % RB);
        double v = 0.0;
        if (cb_abs >= 0) v = a_entry(8 * blk + i, 8 * cb_abs + (cs & 7));
        sWf[(8 * slot + i) * LDWF + cs] = v;
      }
    };
    // right-hand-side tile (C layout) of block blk for this warp's columns
    auto rhs_tile = [&](int blk, double &c0, double &c1) {
      const int r = 8 * blk + g, col = 8 * warp + 2 * t;
      c0 = c1 = 0.0;
      if (r < Ni) {
        int a[3], ca;
        idof_to_node(geo, r, a, ca);
        if (col < ncd) c0 = proj_entry(cP, geo, a, ca, col);
        if (col + 1 < ncd) c1 = proj_entry(cP, geo, a, ca, col + 1);
      }
    };

    double creg[RBMAX][2];
#pragma unroll
    for (int sb = 0; sb < RBMAX; ++sb) {
      creg[sb][0] = creg[sb][1] = 0.0;
      if (sb < RB && sb < NBLK) rhs_tile(sb, creg[sb][0], creg[sb][1]);
    }
    for (int b = 0; b < RB && b < NBLK; ++b) assemble_block(b);
    __syncthreads();
    int bad = 0;
    if (warp == 0) {
      const double v0 = sWf[g * LDWF + 2 * t], v1 = sWf[g * LDWF + 2 * t + 1];
      bad |= chol8_inv(v0, v1, lane, sLd, sLinv);
    }
    __syncthreads();

    // ============================ factorisation + forward substitution ============================
    for (int k = 0; k < NBLK; ++k) {
      const int kslot = k % RB, cur = k & 1;
      const double *Linv = sLinv + cur * 64;
      int nl = NBLK - 1 - k;  // live panel blocks below the diagonal block
      if (nl > RB - 1) nl = RB - 1;
      double *Ls = myL + (size_t)k * LSTEP;
      // ---- S2a: panel tiles  Lp_I = W[I, D] * Linv^T ----
      const double binv0 = Linv[g * 8 + t], binv1 = Linv[g * 8 + 4 + t];  // B[k][n] = Linv[n][k]
      for (int off = 1 + warp; off <= nl; off += NW) {
        int sI = kslot + off;
        if (sI >= RB) sI -= RB;
        const double *wt = sWf + (8 * sI + g) * LDWF + 8 * kslot;
        double p0 = 0.0, p1 = 0.0;
        dmma884(p0, p1, wt[t], binv0);
        dmma884(p0, p1, wt[4 + t], binv1);
        sLpT[(2 * t) * LDP + 8 * sI + g] = p0;
        sLpT[(2 * t + 1) * LDP + 8 * sI + g] = p1;
        *reinterpret_cast<double2 *>(Ls + 64 + (8 * (off - 1) + g) * 8 + 2 * t) = make_double2(p0, p1);
      }
      if (warp == NW - 1) {
        Ls[lane] = Linv[lane];
        Ls[32 + lane] = Linv[32 + lane];
      }
      // ---- S2b: Y_D = Linv * R_D for this warp's columns ----
      double rd0 = 0.0, rd1 = 0.0;
#pragma unroll
      for (int sb = 0; sb < RBMAX; ++sb)
        if (sb == kslot) { rd0 = creg[sb][0]; rd1 = creg[sb][1]; }
      double y0 = 0.0, y1 = 0.0;
      {
        const double b0 = c_to_b(rd0, rd1, lane, 0), b1 = c_to_b(rd0, rd1, lane, 1);
        dmma884(y0, y1, Linv[g * 8 + t], b0);
        dmma884(y0, y1, Linv[g * 8 + 4 + t], b1);
      }
      *reinterpret_cast<double2 *>(X + (size_t)(8 * k + g) * lay.ldx + 8 * warp + 2 * t) = make_double2(y0, y1);
      const double yb0 = -c_to_b(y0, y1, lane, 0), yb1 = -c_to_b(y0, y1, lane, 1);
      __syncthreads();
      // ---- S3: trailing update of the window (shared C), look-ahead factorisation, RHS tiles (register C) ----
      const int ntile = nl * (nl + 1) / 2;
      for (int tt = warp; tt < ntile; tt += NW) {
        // tt = offI (offI - 1) / 2 + offJ - 1
        int offI = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)tt)) * 0.5f);
        while (offI * (offI - 1) / 2 > tt) --offI;
        while ((offI + 1) * offI / 2 <= tt) ++offI;
        const int offJ = tt - offI * (offI - 1) / 2 + 1;
        int sI = kslot + offI, sJ = kslot + offJ;
        if (sI >= RB) sI -= RB;
        if (sJ >= RB) sJ -= RB;
        double *ct = sWf + (8 * sI + g) * LDWF + 8 * sJ + 2 * t;
        double2 c = *reinterpret_cast<double2 *>(ct);
        const double a0 = -sLpT[t * LDP + 8 * sI + g], a1 = -sLpT[(4 + t) * LDP + 8 * sI + g];
        const double b0 = sLpT[t * LDP + 8 * sJ + g], b1 = sLpT[(4 + t) * LDP + 8 * sJ + g];
        dmma884(c.x, c.y, a0, b0);
        dmma884(c.x, c.y, a1, b1);
        *reinterpret_cast<double2 *>(ct) = c;
        if (tt == 0) {  // next diagonal block: factor it now (look-ahead), warp 0 only
          bad |= chol8_inv(c.x, c.y, lane, sLd, sLinv + (cur ^ 1) * 64);
        }
      }
#pragma unroll
      for (int sb = 0; sb < RBMAX; ++sb) {
        if (sb < RB) {
          int off = sb - kslot;
          if (off < 0) off += RB;
          if (off >= 1 && off <= nl) {
            const double a0 = sLpT[t * LDP + 8 * sb + g], a1 = sLpT[(4 + t) * LDP + 8 * sb + g];
            dmma884(creg[sb][0], creg[sb][1], a0, yb0);
            dmma884(creg[sb][0], creg[sb][1], a1, yb1);
          }
        }
      }
      // ---- slide: block k + RB takes the slot of block k ----
      if (k + RB < NBLK) {
        assemble_block(k + RB);
        double n0, n1;
        rhs_tile(k + RB, n0, n1);
#pragma unroll
        for (int sb = 0; sb < RBMAX; ++sb)
          if (sb == kslot) { creg[sb][0] = n0; creg[sb][1] = n1; }
      }
      __syncthreads();
    }
    if (bad && lane == 0) atomicOr(&status[pid], 1);

    // ====================================== backward substitution ======================================
    double xb[RBMAX][2];
#pragma unroll
    for (int sb = 0; sb < RBMAX; ++sb) xb[sb][0] = xb[sb][1] = 0.0;
    auto stage_step = [&](int k, int buf) {  // L workspace of step k -> shared (Linv + panel rows)
      const double *Ls = myL + (size_t)k * LSTEP;
      int nl = NBLK - 1 - k;
      if (nl > RB - 1) nl = RB - 1;
      double *dst = sLpR + buf * R * 8;
      for (int idx = tid; idx < 64 + nl * 64; idx += NT) {
        if (idx < 64) sLinv[buf * 64 + idx] = Ls[idx];
        else dst[idx - 64] = Ls[idx];
      }
    };
    stage_step(NBLK - 1, (NBLK - 1) & 1);
    __syncthreads();
    for (int k = NBLK - 1; k >= 0; --k) {
      const int kslot = k % RB, buf = k & 1;
      int nl = NBLK - 1 - k;
      if (nl > RB - 1) nl = RB - 1;
      if (k > 0) stage_step(k - 1, buf ^ 1);
      const double *Lp = sLpR + buf * R * 8;
      const double *Linv = sLinv + buf * 64;
      double2 yv = *reinterpret_cast<const double2 *>(X + (size_t)(8 * k + g) * lay.ldx + 8 * warp + 2 * t);
      double c0 = yv.x, c1 = yv.y;
#pragma unroll
      for (int sb = 0; sb < RBMAX; ++sb) {
        if (sb < RB) {
          int off = sb - kslot;
          if (off < 0) off += RB;
          if (off >= 1 && off <= nl) {
            // A = Lp^T : A[m = g][kk = 4j + t] = Lp[8 (off-1) + 4j + t][g]
            const double a0 = -Lp[(8 * (off - 1) + t) * 8 + g], a1 = -Lp[(8 * (off - 1) + 4 + t) * 8 + g];
            dmma884(c0, c1, a0, xb[sb][0]);
            dmma884(c0, c1, a1, xb[sb][1]);
          }
        }
      }
      // X_D = Linv^T T
      const double tb0 = c_to_b(c0, c1, lane, 0), tb1 = c_to_b(c0, c1, lane, 1);
      double x0 = 0.0, x1 = 0.0;
      dmma884(x0, x1, Linv[t * 8 + g], tb0);
      dmma884(x0, x1, Linv[(4 + t) * 8 + g], tb1);
      *reinterpret_cast<double2 *>(X + (size_t)(8 * k + g) * lay.ldx + 8 * warp + 2 * t) = make_double2(x0, x1);
      const double nb0 = c_to_b(x0, x1, lane, 0), nb1 = c_to_b(x0, x1, lane, 1);
#pragma unroll
      for (int sb = 0; sb < RBMAX; ++sb)
        if (sb == kslot) { xb[sb][0] = nb0; xb[sb][1] = nb1; }
      __syncthreads();
    }
  }
}

}  // namespace slod
