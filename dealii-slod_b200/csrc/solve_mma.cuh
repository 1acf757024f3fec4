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

#ifdef SLOD_PHASE_CLOCKS
#define PH_DECL long long ph_acc[12] = {0}; long long ph_last = clock64(); const bool ph_on = (blockIdx.x == 0) && ((threadIdx.x & 31) == 0) && ((threadIdx.x >> 5) == 0 || (threadIdx.x >> 5) == 5);
#define PH(i) if (ph_on) { const long long c_ = clock64(); ph_acc[i] += c_ - ph_last; ph_last = c_; }
#define PH_PRINT(tag) if (ph_on) printf(tag " warp %d: %lld %lld %lld %lld %lld %lld | %lld %lld %lld %lld %lld %lld\n", (int)(threadIdx.x >> 5), ph_acc[0], ph_acc[1], ph_acc[2], ph_acc[3], ph_acc[4], ph_acc[5], ph_acc[6], ph_acc[7], ph_acc[8], ph_acc[9], ph_acc[10], ph_acc[11]);
#else
#define PH_DECL
#define PH(i)
#define PH_PRINT(tag)
#endif

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

// L^{-1} of the Cholesky factor of an 8x8 SPD tile held in C layout by one warp, all in registers: the row
// operations of the (LDL^T) elimination are applied to an identity tile M alongside, so that after 8 steps
// M = Ltilde^{-1} and L^{-1} = D^{-1/2} M.  Only the lower triangle of the tile is read.  Writes L^{-1} (row-major
// 8x8) to sLinv; returns nonzero if a pivot was not positive.
__device__ __forceinline__ int chol8_inv(double v0, double v1, int lane, double *sLinv, double *pivot_min = nullptr) {
  const int g = lane >> 2, t = lane & 3;
  double m0 = (g == 2 * t) ? 1.0 : 0.0, m1 = (g == 2 * t + 1) ? 1.0 : 0.0;
  double myinv = 0.0, pmin = 1e300;
  int bad = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const double sel = (k & 1) ? v1 : v0;  // column k lives in the lanes with t == k >> 1
    const double dkk = __shfl_sync(0xffffffffu, sel, 4 * k + (k >> 1));
    const double agk = __shfl_sync(0xffffffffu, sel, 4 * g + (k >> 1));
    const double ac0 = __shfl_sync(0xffffffffu, sel, 4 * (2 * t) + (k >> 1));
    const double ac1 = __shfl_sync(0xffffffffu, sel, 4 * (2 * t + 1) + (k >> 1));
    const double mk0 = __shfl_sync(0xffffffffu, m0, 4 * k + t), mk1 = __shfl_sync(0xffffffffu, m1, 4 * k + t);
    if (!(dkk > 0.0)) bad = 1;
    pmin = fmin(pmin, dkk);
    const double rs = rsqrt(dkk);  // off the dependency chain (only the final scaling needs it)
    myinv = (g == k) ? rs : myinv;
    // 1 / dkk on the chain: MUFU seed + one cubic Newton step (error ~ e^3, e ~ 2^-20)
    double rc;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rc) : "d"(dkk));
    const double er = fma(-dkk, rc, 1.0);
    rc = fma(rc, fma(er, er, er), rc);
    const double mg = (g > k) ? agk * rc : 0.0;
    v0 -= mg * ac0;
    v1 -= mg * ac1;
    m0 -= mg * mk0;
    m1 -= mg * mk1;
  }
  *reinterpret_cast<double2 *>(sLinv + g * 8 + 2 * t) = make_double2(m0 * myinv, m1 * myinv);
  if (pivot_min) *pivot_min = pmin;
  __syncwarp();
  return bad;
}

// The same factorisation without any shuffle: every lane loads the lower triangle of the tile (broadcast reads of
// shared memory, row stride ld) into registers and runs the whole LDL^T elimination redundantly -- all indices are
// compile-time constants, the dependency chain per pivot is one MUFU reciprocal + Newton step, one multiplier and one
// update, about half the latency of the shuffle version.  Lane l carries column (l & 7) of the transform M alongside;
// lanes 0..7 write L^{-1} = D^{-1/2} M (row-major 8 x 8, zeros above the diagonal) to sLinv.  Returns nonzero if a
// pivot was not positive.
__device__ __forceinline__ int chol8_inv_reg(const double *T, int ld, int lane, double *sLinv) {
  double a[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int j = 0; j + 1 <= i; j += 2) {
      const double2 v = *reinterpret_cast<const double2 *>(T + i * ld + j);
      a[i][j] = v.x;
      a[i][j + 1] = v.y;
    }
    if ((i & 1) == 0) a[i][i] = T[i * ld + i];
  }
  const int col = lane & 7;
  double m[8], rs[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = (i == col) ? 1.0 : 0.0;
  int bad = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const double dkk = a[k][k];
    if (!(dkk > 0.0)) bad = 1;
    rs[k] = dkk;          // scaled after the loop: the eight inverse square roots are independent of each other
    double rc;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rc) : "d"(dkk));
    const double er = fma(-dkk, rc, 1.0);
    rc = fma(rc, fma(er, er, er), rc);
    double mg[8];
#pragma unroll
    for (int i = k + 1; i < 8; ++i) mg[i] = a[i][k] * rc;
    // the next pivot first: it heads the dependency chain
#pragma unroll
    for (int i = k + 1; i < 8; ++i)
#pragma unroll
      for (int j = k + 1; j <= i; ++j) a[i][j] = fma(-mg[i], a[j][k], a[i][j]);
#pragma unroll
    for (int i = k + 1; i < 8; ++i) m[i] = fma(-mg[i], m[k], m[i]);
  }
  // 1 / sqrt(d_k): MUFU seed (2^-22) + two Newton steps, branch free (the pivots are positive normal numbers), so the
  // eight sequences interleave instead of running one after the other like eight calls of rsqrt() would
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(rs[k]));
    const double hx = 0.5 * rs[k];
    double e = fma(-hx * y, y, 0.5);
    y = fma(y, e, y);
    e = fma(-hx * y, y, 0.5);
    rs[k] = fma(y, e, y);
  }
  if (lane < 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) sLinv[i * 8 + col] = (i >= col) ? m[i] * rs[i] : 0.0;
  }
  __syncwarp();
  return bad;
}

constexpr int kSolveStages = 4;  // backward sweep: L panels staged this many steps ahead (they come back from HBM)

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
                  double *__restrict__ Xbuf, double *__restrict__ Lws, int *__restrict__ status, SolveMmaLayout lay,
                  int *work_counter) {
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
  double *sLpR = sLpT + 8 * LDP;            // [NSTG][R*8]  panel rows (row-major by offset) for the backward pass
  double *sLinv = sLpR + kSolveStages * R * 8;  // [NSTG][64]
  double *sSt = sLinv + kSolveStages * 64;                // [nip_max][stw] lower stencil values of A_ii, one row per dof
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

  __shared__ int sNextWork;
  SLOD_WORK_LOOP(w, n_work, work_counter, sNextWork) {
    fetch_work_item(w, work_counter, &sNextWork);
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

    // the two coarse columns of this thread's C-fragment entries (constant per patch)
    const int mycc0 = sColCell[8 * warp + 2 * t], mycc1 = sColCell[8 * warp + 2 * t + 1];

    // Block row `sB` of the window is rebuilt in two phases that an existing barrier separates: all threads zero
    // it, then the last warps scatter the (at most 14 * spacedim) lower-stencil entries of each of its 8 rows;
    // rows >= Ni are identity rows.
    auto zero_block_row = [&](int sB) {
      for (int item = tid; item < 32 * RB; item += NT) {
        const int i = item & 7, q = item >> 3;  // row, column pair
        *reinterpret_cast<double2 *>(sWf + (8 * sB + i) * LDWF + 2 * q) = make_double2(0.0, 0.0);
      }
    };
    auto scatter_block_row = [&](int blk, int sB) {
      const int per_row = (nlow + 1) * sdim;
      for (int item = NT - 1 - tid; item < 8 * per_row; item += NT) {
        const int i = item & 7, ee = item >> 3;
        const int e = ee / sdim, cb = ee - e * sdim;
        const int r = 8 * blk + i;
        const int pk = sRowPk[r];
        if (pk < 0) {  // bit 31: a real dof row
          const int ca = (pk >> 15) & 1;
          if (!((pk >> (16 + e)) & 1) || (e == nlow && cb > ca)) continue;
          const int c = r + sDOff[e] + (cb - ca);
          int sb = sB - (blk - (c >> 3));
          if (sb < 0) sb += RB;
          sWf[(8 * sB + i) * LDWF + 8 * sb + (c & 7)] = sSt[(size_t)r * stw + ee];
        } else if (ee == 0) {
          sWf[(8 * sB + i) * LDWF + 8 * sB + i] = 1.0;
        }
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
        const int cc = h ? mycc1 : mycc0;
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

    // cr[off]: right-hand-side tile (C layout) of block k + off; the window rotates by register moves so that every
    // index below is static
    double cr[RBMAX][2];
#pragma unroll
    for (int off = 0; off < RBMAX; ++off) {
      cr[off][0] = cr[off][1] = 0.0;
      if (off < RB && off < NBLK) rhs_tile(off, cr[off][0], cr[off][1]);
    }
    for (int b = 0; b < RB && b < NBLK; ++b) zero_block_row(b);
    __syncthreads();
    for (int b = 0; b < RB && b < NBLK; ++b) scatter_block_row(b, b);
    __syncthreads();
    int bad = 0;
    if (warp == 0) {
      const double v0 = sWf[g * LDWF + 2 * t], v1 = sWf[g * LDWF + 2 * t + 1];
      bad |= chol8_inv(v0, v1, lane, sLinv);
    }
    __syncthreads();

    // ============================ factorisation + forward substitution ============================
    PH_DECL
    for (int k = 0, kslot = 0; k < NBLK; ++k, kslot = (kslot + 1 == RB) ? 0 : kslot + 1) {
      PH(0)
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
      double y0 = 0.0, y1 = 0.0;
      {
        const double b0 = c_to_b(cr[0][0], cr[0][1], lane, 0), b1 = c_to_b(cr[0][0], cr[0][1], lane, 1);
        dmma884(y0, y1, Linv[g * 8 + t], b0);
        dmma884(y0, y1, Linv[g * 8 + 4 + t], b1);
      }
      *reinterpret_cast<double2 *>(X + (size_t)(8 * k + g) * lay.ldx + 8 * warp + 2 * t) = make_double2(y0, y1);
      const double yb0 = -c_to_b(y0, y1, lane, 0), yb1 = -c_to_b(y0, y1, lane, 1);
      // block row kslot (block k) is dead: only its diagonal tile was still needed, by the factorisation of the
      // previous step.  Clear it for block k + RB; the scatter follows after the barrier.
      if (k + RB < NBLK) zero_block_row(kslot);
      PH(1)
      __syncthreads();
      PH(2)
      // ---- S3: trailing update of the window (shared C), look-ahead factorisation, RHS tiles (register C) ----
      const int ntile = nl * (nl + 1) / 2;
      if (warp == 0) {
        // warp 0 owns the look-ahead chain: next diagonal tile, then its factorisation
        if (ntile > 0) {
          int sI = kslot + 1;
          if (sI >= RB) sI -= RB;
          double *ct = sWf + (8 * sI + g) * LDWF + 8 * sI + 2 * t;
          double2 c = *reinterpret_cast<double2 *>(ct);
          const double a0 = -sLpT[t * LDP + 8 * sI + g], a1 = -sLpT[(4 + t) * LDP + 8 * sI + g];
          dmma884(c.x, c.y, a0, -a0);
          dmma884(c.x, c.y, a1, -a1);
          *reinterpret_cast<double2 *>(ct) = c;
          bad |= chol8_inv(c.x, c.y, lane, sLinv + (cur ^ 1) * 64);
        }
      }
      // The warps that share warp 0's scheduler hold their DMMAs back until the factorisation chain is through:
      // its dependent fp64 operations would otherwise queue behind them on the shared fp64 pipe.
      if ((warp & 3) == 0) asm volatile("bar.sync 1, %0;" ::"n"(32 * (NW / 4)) : "memory");
      // the other tiles go round robin over the warps of the three schedulers warp 0 does not sit on (its fp64
      // dependency chain shares the pipe with their DMMAs), three at a time for instruction-level parallelism
      auto do_trailing = [&]() {
        constexpr int NTW = NW - NW / 4;
        const int wrank = warp - 1 - (warp >> 2);  // 0 .. NTW-1
        for (int tt = 1 + wrank; tt < ntile; tt += 3 * NTW) {
          double *ct[3];
          double2 c[3];
          double a0[3], a1[3], b0[3], b1[3];
#pragma unroll
          for (int u = 0; u < 3; ++u) {
            const int tu = tt + u * NTW;
            const int tab = sTileTab[tu < ntile ? tu : tt];
            const int offI = tab & 0xff, offJ = tab >> 8;
            int sI = kslot + offI, sJ = kslot + offJ;
            if (sI >= RB) sI -= RB;
            if (sJ >= RB) sJ -= RB;
            ct[u] = sWf + (8 * sI + g) * LDWF + 8 * sJ + 2 * t;
            c[u] = *reinterpret_cast<double2 *>(ct[u]);
            a0[u] = -sLpT[t * LDP + 8 * sI + g];
            a1[u] = -sLpT[(4 + t) * LDP + 8 * sI + g];
            b0[u] = sLpT[t * LDP + 8 * sJ + g];
            b1[u] = sLpT[(4 + t) * LDP + 8 * sJ + g];
          }
#pragma unroll
          for (int u = 0; u < 3; ++u) dmma884(c[u].x, c[u].y, a0[u], b0[u]);
#pragma unroll
          for (int u = 0; u < 3; ++u) dmma884(c[u].x, c[u].y, a1[u], b1[u]);
#pragma unroll
          for (int u = 0; u < 3; ++u)
            if (tt + u * NTW < ntile) *reinterpret_cast<double2 *>(ct[u]) = c[u];
        }
      };
      auto do_rhs = [&]() {
#pragma unroll
        for (int off = 1; off < RBMAX; ++off) {
          if (off <= nl) {
            int sI = kslot + off;
            if (sI >= RB) sI -= RB;
            const double a0 = sLpT[t * LDP + 8 * sI + g], a1 = sLpT[(4 + t) * LDP + 8 * sI + g];
            dmma884(cr[off][0], cr[off][1], a0, yb0);
            dmma884(cr[off][0], cr[off][1], a1, yb1);
          }
        }
      };
      // Skew: on each scheduler half of the warps run the (latency-bound) trailing tiles first and the (DMMA-bound)
      // right-hand-side update second, the other half the other way round, so that the two overlap on the fp64 pipe.
      if ((warp & 3) == 0) {
        do_rhs();
      } else if ((warp >> 2) & 1) {
        do_rhs();
        do_trailing();
      } else {
        do_trailing();
        do_rhs();
      }
      PH(3)
      PH(4)
      // ---- slide: block k + RB takes the slot of block k; the register window moves up by one ----
      double n0 = 0.0, n1 = 0.0;
      if (k + RB < NBLK) {
        scatter_block_row(k + RB, kslot);
        rhs_tile(k + RB, n0, n1);
      }
#pragma unroll
      for (int off = 0; off < RBMAX - 1; ++off) {
        cr[off][0] = (off == RB - 1) ? n0 : cr[off + 1][0];
        cr[off][1] = (off == RB - 1) ? n1 : cr[off + 1][1];
      }
      if (RB == RBMAX) { cr[RBMAX - 1][0] = n0; cr[RBMAX - 1][1] = n1; }
      PH(5)
      __syncthreads();
    }
    PH(0)
    if (bad && lane == 0) atomicOr(&status[pid], 1);

    // ====================================== backward substitution ======================================
    // xr[off]: B-fragments of the solved block k + off
    double xr[RBMAX][2];
#pragma unroll
    for (int off = 0; off < RBMAX; ++off) xr[off][0] = xr[off][1] = 0.0;
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
    // Two block rows per iteration (k and k - 1): both use the same X fragments for all but one term, so a
    // single pass over the register window feeds four accumulator chains, and the staging / barrier / shift overhead
    // is paid once per pair.  Stages: buffers (k & 3) and ((k - 1) & 3) are live while the next pair is in flight.
    static_assert(kSolveStages == 4, "the pairwise backward sweep uses two live and two in-flight stages");
    auto y_tile = [&](int k) -> double2 {
      return (k >= 0) ? *reinterpret_cast<const double2 *>(X + (size_t)(8 * k + g) * lay.ldx + 8 * warp + 2 * t)
                      : make_double2(0.0, 0.0);
    };
    for (int j = 0; j < 2; ++j) {
      if (NBLK - 1 - j >= 0) stage_step(NBLK - 1 - j, (NBLK - 1 - j) & 3);
      else __pipeline_commit();
    }
    double2 ya = y_tile(NBLK - 1), yb = y_tile(NBLK - 2);
    __pipeline_wait_prior(0);
    __syncthreads();
    for (int k = NBLK - 1; k >= 0; k -= 2) {
      PH(6)
      const bool two = (k >= 1);
      int nl = NBLK - 1 - k;   // live panel blocks of block row k; block row k - 1 has one more (capped)
      if (nl > RB - 1) nl = RB - 1;
      int nl2 = NBLK - k;
      if (nl2 > RB - 1) nl2 = RB - 1;
      // next pair: stages and Y tiles
      if (k - 2 >= 0) stage_step(k - 2, (k - 2) & 3); else __pipeline_commit();
      if (k - 3 >= 0) stage_step(k - 3, (k - 3) & 3); else __pipeline_commit();
      const double2 yna = y_tile(k - 2), ynb = y_tile(k - 3);
      const double *Lp = sLpR + (k & 3) * R * 8, *Linv = sLinv + (k & 3) * 64;
      const double *Lp2 = sLpR + ((k - 1) & 3) * R * 8, *Linv2 = sLinv + ((k - 1) & 3) * 64;
      double c0 = ya.x, c1 = ya.y, e0 = 0.0, e1 = 0.0;     // block row k     (two chains)
      double p0 = yb.x, p1 = yb.y, q0 = 0.0, q1 = 0.0;     // block row k - 1 (two chains)
      PH(7)
#pragma unroll
      for (int off = 1; off < RBMAX; ++off) {
        // xr[off] = X_{k + off}: offset off for row k, offset off + 1 for row k - 1
        if (off <= nl) {
          // A = Lp^T : A[m = g][kk = 4j + t] = Lp[8 (off-1) + 4j + t][g]
          const double a0 = -Lp[(8 * (off - 1) + t) * 8 + g], a1 = -Lp[(8 * (off - 1) + 4 + t) * 8 + g];
          if (off & 1) {
            dmma884(c0, c1, a0, xr[off][0]);
            dmma884(c0, c1, a1, xr[off][1]);
          } else {
            dmma884(e0, e1, a0, xr[off][0]);
            dmma884(e0, e1, a1, xr[off][1]);
          }
        }
        if (two && off + 1 <= nl2) {
          const double a0 = -Lp2[(8 * off + t) * 8 + g], a1 = -Lp2[(8 * off + 4 + t) * 8 + g];
          if (off & 1) {
            dmma884(p0, p1, a0, xr[off][0]);
            dmma884(p0, p1, a1, xr[off][1]);
          } else {
            dmma884(q0, q1, a0, xr[off][0]);
            dmma884(q0, q1, a1, xr[off][1]);
          }
        }
      }
      c0 += e0;
      c1 += e1;
      PH(8)
      // X_k = Linv_k^T T
      const double tb0 = c_to_b(c0, c1, lane, 0), tb1 = c_to_b(c0, c1, lane, 1);
      double x0 = 0.0, x1 = 0.0;
      dmma884(x0, x1, Linv[t * 8 + g], tb0);
      dmma884(x0, x1, Linv[(4 + t) * 8 + g], tb1);
      *reinterpret_cast<double2 *>(X + (size_t)(8 * k + g) * lay.ldx + 8 * warp + 2 * t) = make_double2(x0, x1);
      const double nb0 = c_to_b(x0, x1, lane, 0), nb1 = c_to_b(x0, x1, lane, 1);
      double mb0 = 0.0, mb1 = 0.0;
      if (two) {
        // the one term of row k - 1 that needs X_k (offset 1), then X_{k-1} = Linv_{k-1}^T T
        p0 += q0;
        p1 += q1;
        dmma884(p0, p1, -Lp2[t * 8 + g], nb0);
        dmma884(p0, p1, -Lp2[(4 + t) * 8 + g], nb1);
        const double ub0 = c_to_b(p0, p1, lane, 0), ub1 = c_to_b(p0, p1, lane, 1);
        double z0 = 0.0, z1 = 0.0;
        dmma884(z0, z1, Linv2[t * 8 + g], ub0);
        dmma884(z0, z1, Linv2[(4 + t) * 8 + g], ub1);
        *reinterpret_cast<double2 *>(X + (size_t)(8 * (k - 1) + g) * lay.ldx + 8 * warp + 2 * t) = make_double2(z0, z1);
        mb0 = c_to_b(z0, z1, lane, 0);
        mb1 = c_to_b(z0, z1, lane, 1);
      }
      // the register window moves down by two block rows
#pragma unroll
      for (int off = RBMAX - 1; off >= 3; --off) { xr[off][0] = xr[off - 2][0]; xr[off][1] = xr[off - 2][1]; }
      if (RBMAX > 2) { xr[2][0] = nb0; xr[2][1] = nb1; }
      xr[1][0] = mb0;
      xr[1][1] = mb1;
      ya = yna;
      yb = ynb;
      PH(9)
      __pipeline_wait_prior(0);
      __syncthreads();
      PH(10)
    }
    PH_PRINT("solve")
  }
}

}  // namespace slod
