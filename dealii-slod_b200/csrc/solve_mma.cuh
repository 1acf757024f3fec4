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
#include <cuda_pipeline.h>

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
  double rdiag[8];  // 1 / L_kk
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const double sel = (k & 1) ? v1 : v0;
    const double dkk = __shfl_sync(0xffffffffu, sel, 4 * k + (k >> 1));
    if (!(dkk > 0.0)) bad = 1;
    const double inv = rsqrt(dkk);
    rdiag[k] = inv;
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
      x[i] = (i >= j) ? sum * rdiag[i] : 0.0;
      if (i >= j) sLinv[i * 8 + j] = x[i];
    }
  }
  __syncwarp();
  return bad;
}

struct SolveMmaLayout {
  int coef_doubles;
  int nip_max;            // interior dofs per patch rounded up to 8
  int stw;                // lower-stencil entries per dof row: spacedim * (13 | 4) + spacedim
  int ldx;                // leading dimension of X rows (8 * NW)
  long long x_stride;     // doubles per patch in Xbuf (NiPmax * ldx)
  long long lws_per_cta;  // doubles of L workspace per CTA  (steps_max * (64 + 8*(RBMAX-1)*8))
};

// number of "lower" stencil nodes (linear offset < 0): 13 in 3-D, 4 in 2-D
__device__ __forceinline__ int n_lower_nodes() { return cP.dim == 3 ? 13 : 4; }
// e-th lower stencil node offset (dx,dy,dz); e == n_lower_nodes() is the node itself
__device__ __forceinline__ void lower_offset(int e, int dl[3]) {
  // enumeration of {-1,0,1}^dim in x-fastest order, truncated before the centre: e = (dx+1) + 3 (dy+1) + 9 (dz+1)
  if (cP.dim == 3) {
    dl[0] = e % 3 - 1; dl[1] = (e / 3) % 3 - 1; dl[2] = e / 9 - 1;
  } else {
    dl[0] = e % 3 - 1; dl[1] = e / 3 - 1; dl[2] = 0;
  }
}

template <int RBMAX, int NW>
__global__ void __launch_bounds__(32 * NW, 1)
k_patch_solve_mma(const int *__restrict__ patch_ids, int n_work, const double *__restrict__ d_coef,
                  double *__restrict__ Xbuf, double *__restrict__ Lws, int *__restrict__ status, SolveMmaLayout lay) {
  constexpr int R = 8 * RBMAX;
  constexpr int LDWF = (R % 16 == 8) ? R : R + 8;  // row stride with LDWF % 16 == 8 : conflict-free C fragments
  constexpr int LDP = R + 4;                        // k-major panel copy, LDP % 16 in {4, 12}
  constexpr int NT = 32 * NW;
  constexpr int NC = 8 * NW;
  constexpr int LSTEP = 64 + 8 * (RBMAX - 1) * 8;  // doubles per step in the L workspace: Linv + panel rows
  constexpr int NTT = RBMAX * (RBMAX - 1) / 2;
  extern __shared__ double smem[];
  double *sCoef = smem;
  double *sWf = sCoef + lay.coef_doubles;   // [R][LDWF] circular dense window of the trailing matrix
  double *sLpT = sWf + R * LDWF;            // [8][LDP]  panel, k-major, slot rows
  double *sLpR = sLpT + 8 * LDP;            // [2][R*8]  panel rows (row-major by offset) for the backward pass
  double *sLd = sLpR + 2 * R * 8;           // [64] scratch
  double *sLinv = sLd + 64;                 // [2][64]
  double *sSt = sLinv + 128;                // [nip_max][stw] lower stencil values of A_ii, one row per dof
  int *sRowPk = (int *)(sSt + (size_t)lay.nip_max * lay.stw);  // [nip_max] packed node coords / comp / mask
  int *sColCell = sRowPk + lay.nip_max;     // [NC] packed cell coords of each coarse column (or -1)
  int *sTileTab = sColCell + NC;            // [NTT] (offI | offJ << 8) of the trailing-update tiles
  int *sDOff = sTileTab + NTT;              // [32] dof offset of each lower stencil slot
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  double *myL = Lws + (size_t)blockIdx.x * lay.lws_per_cta;
  const int sdim = cP.s, stw = lay.stw, nlow = n_lower_nodes();

  for (int tt = tid; tt < NTT; tt += NT) {
    int offI = 1;
    while ((offI + 1) * offI / 2 <= tt) ++offI;
    sTileTab[tt] = offI | ((tt - offI * (offI - 1) / 2 + 1) << 8);
  }

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

    // ---- per-patch tables: the lower stencil row of every interior dof, packed coordinates, column cells ----
    for (int r = tid; r < 8 * NBLK; r += NT) {
      int pk = 0;
      if (r < Ni) {
        int a[3], ca;
        idof_to_node(geo, r, a, ca);
        int mask = 0;
        for (int e = 0; e <= nlow; ++e) {
          int dl[3];
          lower_offset(e, dl);
          int b[3] = {a[0] + dl[0], a[1] + dl[1], a[2] + dl[2]};
          bool inside = true;
          _Pragma("unroll") for (int x = 0; x < 3; ++x) if (x < cP.dim) inside = inside && (b[x] >= 1 && b[x] <= geo.p[x] - 2);
          for (int cb = 0; cb < sdim; ++cb) {
            const bool use = inside && (e < nlow || cb <= ca);
            sSt[(size_t)r * stw + e * sdim + cb] = use ? stiff_entry(cP, geo, sCoef, a, dl, ca, cb) : 0.0;
          }
          if (inside) mask |= 1 << e;
        }
        pk = a[0] | (a[1] << 5) | (a[2] << 10) | (ca << 15) | (mask << 16) | (1 << 31);
      }
      sRowPk[r] = pk;
    }
    for (int col = tid; col < NC; col += NT) {
      int v = -1;
      if (col < ncd) {
        int kc[3];
        col_to_cell(cP, geo, col / sdim, kc);
        v = kc[0] | (kc[1] << 5) | (kc[2] << 10) | ((col % sdim) << 15);
      }
      sColCell[col] = v;
    }
    if (tid <= nlow) {
      int dl[3];
      lower_offset(tid, dl);
      sDOff[tid] = (dl[0] + geo.q[0] * (dl[1] + geo.q[1] * dl[2])) * sdim;
    }
    __syncthreads();

    // write the 8 rows of block `blk` (slot sB) for all live column slots; entries outside the band are zero,
    // rows >= Ni are identity rows.  One (row, column-slot-block) pair per thread, no divisions.
    auto assemble_block = [&](int blk, int sB) {
      for (int pair = tid; pair < 8 * RB; pair += NT) {
        const int i = pair & 7, sb = pair >> 3;
        int back = sB - sb;
        if (back < 0) back += RB;
        const int cb_abs = blk - back;  // absolute column block living in slot sb
        const int r = 8 * blk + i;
        double out[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) out[c] = 0.0;
        const int pk = sRowPk[r];
        if (pk < 0) {  // bit 31: a real dof row
          const int ca = (pk >> 15) & 1, mask = (pk >> 16) & 0x7fff;
          for (int e = 0; e <= nlow; ++e) {
            if (!((mask >> e) & 1)) continue;
            for (int cb = 0; cb < sdim; ++cb) {
              if (e == nlow && cb > ca) continue;
              const int c = r + sDOff[e] + (cb - ca);
              if ((c >> 3) == cb_abs) {
                const double v = sSt[(size_t)r * stw + e * sdim + cb];
#pragma unroll
                for (int cc = 0; cc < 8; ++cc)
                  if (cc == (c & 7)) out[cc] = v;
              }
            }
          }
        } else if (back == 0) {
#pragma unroll
          for (int cc = 0; cc < 8; ++cc)
            if (cc == i) out[cc] = 1.0;
        }
        double2 *dst = reinterpret_cast<double2 *>(sWf + (8 * sB + i) * LDWF + 8 * sb);
        dst[0] = make_double2(out[0], out[1]);
        dst[1] = make_double2(out[2], out[3]);
        dst[2] = make_double2(out[4], out[5]);
        dst[3] = make_double2(out[6], out[7]);
      }
    };
    // right-hand-side tile (C layout) of block blk for this warp's columns: P_i entries from the tables
    auto rhs_tile = [&](int blk, double &c0, double &c1) {
      c0 = c1 = 0.0;
      const int pk = sRowPk[8 * blk + g];
      if (pk >= 0) return;
      const int ca = (pk >> 15) & 1;
      const int n = cP.n;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int cc = sColCell[8 * warp + 2 * t + h];
        if (cc < 0 || ((cc >> 15) & 1) != ca) continue;
        double wgt = cP.pw;
        bool in = true;
        _Pragma("unroll") for (int x = 0; x < 3; ++x) if (x < cP.dim) {
          const int tt = ((pk >> (5 * x)) & 31) - n * ((cc >> (5 * x)) & 31);
          if (tt < 0 || tt > n) in = false;
          if (tt != 0 && tt != n) wgt *= 2.0;
        }
        if (in) { if (h == 0) c0 = wgt; else c1 = wgt; }
      }
    };

    double creg[RBMAX][2];
#pragma unroll
    for (int sb = 0; sb < RBMAX; ++sb) {
      creg[sb][0] = creg[sb][1] = 0.0;
      if (sb < RB && sb < NBLK) rhs_tile(sb, creg[sb][0], creg[sb][1]);
    }
    for (int b = 0; b < RB && b < NBLK; ++b) assemble_block(b, b);
    __syncthreads();
    int bad = 0;
    if (warp == 0) {
      const double v0 = sWf[g * LDWF + 2 * t], v1 = sWf[g * LDWF + 2 * t + 1];
      bad |= chol8_inv(v0, v1, lane, sLd, sLinv);
    }
    __syncthreads();

    // ============================ factorisation + forward substitution ============================
    for (int k = 0, kslot = 0; k < NBLK; ++k, kslot = (kslot + 1 == RB) ? 0 : kslot + 1) {
      const int cur = k & 1;
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
      // warp 0 owns the look-ahead chain (next diagonal tile + its factorisation); the other tiles go round
      // robin over warps 1..NW-1
      for (int tt = (warp == 0) ? 0 : warp; tt < ntile; tt += (warp == 0) ? ntile : NW - 1) {
        const int tab = sTileTab[tt];
        const int offI = tab & 0xff, offJ = tab >> 8;
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
        assemble_block(k + RB, kslot);
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
      for (int idx = tid; idx < 32 + nl * 32; idx += NT) {  // 16-byte pieces, asynchronous (LDGSTS)
        double *d = (idx < 32) ? (sLinv + buf * 64 + 2 * idx) : (dst + 2 * (idx - 32));
        __pipeline_memcpy_async(d, Ls + 2 * idx, 16);
      }
      __pipeline_commit();
    };
    stage_step(NBLK - 1, (NBLK - 1) & 1);
    __pipeline_wait_prior(0);
    __syncthreads();
    for (int k = NBLK - 1, kslot = (NBLK - 1) % RB; k >= 0; --k, kslot = (kslot == 0) ? RB - 1 : kslot - 1) {
      const int buf = k & 1;
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
      __pipeline_wait_prior(0);
      __syncthreads();
    }
  }
}

}  // namespace slod
