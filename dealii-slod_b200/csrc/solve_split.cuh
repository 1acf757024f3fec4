// Split patch solver for the large 3-D patches (cfg 4/5: Ni = 729, half band width 91, 125 right-hand sides):
//
//   k_patch_factor    blocked banded Cholesky of A_ii alone.  Small working set (the 104 x 104 circular window), two
//                     CTAs of 8 warps per SM so that the latency-bound look-ahead chain of one patch (8 x 8 diagonal
//                     factorisation + in-register inverse) hides behind the tensor-core trailing update of the other.
//                     The factor leaves the SM as one record per 8-column step: 13 tiles of 8 x 8 (the inverse of the
//                     diagonal factor, then the 12 panel tiles), each stored in mma.m8n8k4 A-FRAGMENT ORDER -- lane l
//                     owns the two doubles {T[g][t], T[g][4+t]} (g = l >> 2, t = l & 3) -- so that the consumer reads
//                     its A operand with one conflict-free 16-byte load per tile (forward sweep) or two conflict-free
//                     8-byte loads (backward sweep: T^T[g][t] = T[t][g] sits at double 8t + 2(g&3) + (g>>2)).
//   k_patch_trisolve  forward and backward substitution with all right-hand sides.  Pure tensor-core streaming: the 16
//                     warps own 8 right-hand-side columns each and never synchronise with each other -- the records
//                     arrive through an 8-stage ring filled by bulk asynchronous copies (cp.async.bulk + mbarrier
//                     transaction counts), every warp waits on the "full" barrier of a stage and releases it through its
//                     "empty" barrier, there is no block-wide barrier inside a patch.  The right-hand-side window lives
//                     in registers as accumulator tiles (as in k_patch_solve_mma).
//
// Together they replace assemble_stiffness + Gauss_elimination of the reference (source/LOD.cc:433-546,
// include/LODtools.h:511-595).  The fused kernel k_patch_solve_mma remains for the small 2-D patches, where one CTA
// holds everything and the factorisation chain is short.  Included by kernels.cu after solve_mma.cuh.
#pragma once
#include <cuda/std/cstdint>

namespace slod {

// ---- mbarrier / bulk-copy primitives (PTX ISA 8.x, sm_90+) ---------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// global -> shared bulk copy (TMA engine, no tensor map), completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

constexpr int kTriStages = 8;   // ring depth of k_patch_trisolve (8 x 6.5 KB)

struct SplitLayout {
  int coef_doubles;
  int nip_max;            // interior dofs per patch rounded up to 8
  int ldx;                // leading dimension of X rows (8 * NW of the triangular solver)
  long long x_stride;     // doubles per patch in Xbuf
  long long rec_stride;   // doubles per patch in the factor records: (nip_max / 8) * RBMAX * 64
};

// position of element (r, c) of an 8 x 8 tile in A-fragment order
__device__ __forceinline__ int frag_pos(int r, int c) { return 8 * r + 2 * (c & 3) + (c >> 2); }

// Entry of the unconstrained stiffness matrix between interior node a and a + dl for the scalar 3-D problem with one
// coefficient per sub-cell; the 8 x 8 reference matrix comes from shared memory (dynamic indexing of __constant__
// memory serialises over the distinct addresses of a warp).
__device__ __forceinline__ double stiff_entry_3d(const Geom &g, const double *sCoef, const double *sK, int n,
                                                 const int a[3], const int dl[3]) {
  int o0[3], o1[3];
#pragma unroll
  for (int x = 0; x < 3; ++x) {
    const int msub = g.m[x] * n;
    int lo_o = (dl[x] == 1) ? a[x] : a[x] - 1;
    int hi_o = (dl[x] == -1) ? a[x] - 1 : a[x];
    if (lo_o < 0) lo_o = 0;
    if (hi_o > msub - 1) hi_o = msub - 1;
    o0[x] = lo_o;
    o1[x] = hi_o;
  }
  const int msx = g.m[0] * n, msy = g.m[1] * n;
  double acc = 0.0;
  for (int oz = o0[2]; oz <= o1[2]; ++oz)
    for (int oy = o0[1]; oy <= o1[1]; ++oy)
      for (int ox = o0[0]; ox <= o1[0]; ++ox) {
        const int sc = (oz * msy + oy) * msx + ox;
        const int la = (a[0] - ox) + 2 * (a[1] - oy) + 4 * (a[2] - oz);
        const int lb = (a[0] + dl[0] - ox) + 2 * (a[1] + dl[1] - oy) + 4 * (a[2] + dl[2] - oz);
        acc += sCoef[sc] * sK[la * 8 + lb];
      }
  return acc;
}

// =====================================================================================================================
// k_patch_factor
// =====================================================================================================================
template <int RBMAX, int NW>
__global__ void __launch_bounds__(32 * NW, 2)
k_patch_factor(const int *__restrict__ patch_ids, int n_work, const double *__restrict__ d_coef,
               double *__restrict__ Lrec, int *__restrict__ status, SplitLayout lay, int *work_counter) {
  constexpr int R = 8 * RBMAX;
  constexpr int LDWF = (R % 16 == 8) ? R : R + 8;  // row stride with LDWF % 16 == 8 : conflict-free C fragments
  constexpr int LDP = R + 4;                        // k-major panel copy
  constexpr int NT = 32 * NW;
  constexpr int LSTEP = RBMAX * 64;                 // doubles per step record
  constexpr int NTT = RBMAX * (RBMAX - 1) / 2;
  extern __shared__ double smem[];
  double *sCoef = smem;
  double *sWf = sCoef + lay.coef_doubles;   // [R][LDWF] circular dense window of the trailing matrix
  double *sLpT = sWf + R * LDWF;            // [8][LDP]  panel, k-major, slot rows
  double *sLinv = sLpT + 8 * LDP;           // [2][64]
  double *sK = sLinv + 2 * 64;              // [64] reference sub-cell matrix
  int *sRowPk = (int *)(sK + 64);           // [nip_max] packed node coords / mask
  int *sTileTab = sRowPk + lay.nip_max;     // [NTT] (offI | offJ << 8) of the trailing-update tiles
  int *sDOff = sTileTab + NTT;              // [16] dof offset of each lower stencil slot
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  constexpr int nlow = 13;                  // lower stencil nodes in 3-D

  for (int tt = tid; tt < NTT; tt += NT) {
    int offI = 1;
    while ((offI + 1) * offI / 2 <= tt) ++offI;
    sTileTab[tt] = offI | ((tt - offI * (offI - 1) / 2 + 1) << 8);
  }
  for (int i = tid; i < 64; i += NT) sK[i] = cP.Kref[i];

  __shared__ int sNextWork;
  SLOD_WORK_LOOP(w, n_work, work_counter, sNextWork) {
    fetch_work_item(w, work_counter, &sNextWork);
    const int pid = patch_ids[w];
    const Geom geo = make_geom(cP, pid);
    const int Ni = geo.Ni, bw = geo.bw, n = cP.n;
    const int NBLK = (Ni + 7) >> 3;
    int RB = (bw + 8 + 7) >> 3;
    if (RB > RBMAX) RB = RBMAX;  // host guarantees bw_max fits
    double *rec = Lrec + (size_t)w * lay.rec_stride;
    __syncthreads();
    load_coef(geo, d_coef, sCoef);
    for (int r = tid; r < 8 * NBLK; r += NT) {
      int pk = 0;
      if (r < Ni) {
        int a[3];
        interior_coords(geo, r, a);
        int mask = 0;
        for (int e = 0; e <= nlow; ++e) {
          int dl[3];
          lower_offset(e, dl);
          bool inside = true;
#pragma unroll
          for (int x = 0; x < 3; ++x) inside = inside && (a[x] + dl[x] >= 1 && a[x] + dl[x] <= geo.p[x] - 2);
          if (inside) mask |= 1 << e;
        }
        pk = a[0] | (a[1] << 5) | (a[2] << 10) | (mask << 16) | (1 << 31);
      }
      sRowPk[r] = pk;
    }
    if (tid <= nlow) {
      int dl[3];
      lower_offset(tid, dl);
      sDOff[tid] = dl[0] + geo.q[0] * (dl[1] + geo.q[1] * dl[2]);
    }
    __syncthreads();

    auto zero_block_row = [&](int sB) {
      for (int item = tid; item < 32 * RB; item += NT) {
        const int i = item & 7, q = item >> 3;
        *reinterpret_cast<double2 *>(sWf + (8 * sB + i) * LDWF + 2 * q) = make_double2(0.0, 0.0);
      }
    };
    // the (at most 14) lower-stencil entries of each of the 8 rows of block `blk` into block row slot sB, computed
    // on the fly from the sub-cell coefficients; item0 / nitem: the threads taking part
    auto scatter_block_row = [&](int blk, int sB, int item0, int nitem) {
      for (int item = item0; item < 8 * (nlow + 1); item += nitem) {
        const int i = item & 7, e = item >> 3;
        const int r = 8 * blk + i;
        const int pk = sRowPk[r];
        if (pk < 0) {  // bit 31: a real dof row
          if (!((pk >> (16 + e)) & 1)) continue;
          const int a[3] = {pk & 31, (pk >> 5) & 31, (pk >> 10) & 31};
          int dl[3];
          lower_offset(e, dl);
          const int c = r + sDOff[e];
          int sb = sB - (blk - (c >> 3));
          if (sb < 0) sb += RB;
          sWf[(8 * sB + i) * LDWF + 8 * sb + (c & 7)] = stiff_entry_3d(geo, sCoef, sK, n, a, dl);
        } else if (e == 0) {
          sWf[(8 * sB + i) * LDWF + 8 * sB + i] = 1.0;
        }
      }
    };

    for (int b = 0; b < RB && b < NBLK; ++b) zero_block_row(b);
    __syncthreads();
    for (int b = 0; b < RB && b < NBLK; ++b) scatter_block_row(b, b, tid, NT);
    __syncthreads();
    int bad = 0;
    if (warp == 0) {
      const double v0 = sWf[g * LDWF + 2 * t], v1 = sWf[g * LDWF + 2 * t + 1];
      bad |= chol8_inv(v0, v1, lane, sLinv);
    }
    __syncthreads();

    for (int k = 0, kslot = 0; k < NBLK; ++k, kslot = (kslot + 1 == RB) ? 0 : kslot + 1) {
      const int cur = k & 1;
      const double *Linv = sLinv + cur * 64;
      int nl = NBLK - 1 - k;  // live panel blocks below the diagonal block
      if (nl > RB - 1) nl = RB - 1;
      double *Ls = rec + (size_t)k * LSTEP;
      // ---- panel tiles  Lp_I = W[I, D] * Linv^T ----
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
        double *dst = Ls + 64 * off;      // element (g, 2t) and (g, 2t + 1) in fragment order
        dst[frag_pos(g, 2 * t)] = p0;
        dst[frag_pos(g, 2 * t + 1)] = p1;
      }
      if (warp == NW - 1)   // tile 0 of the record: Linv in fragment order
        *reinterpret_cast<double2 *>(Ls + 2 * lane) = make_double2(Linv[g * 8 + t], Linv[g * 8 + 4 + t]);
      // block row kslot (block k) is dead: only its diagonal tile was still needed, by the factorisation of the
      // previous step.  Clear it for block k + RB; the scatter follows after the barrier.
      if (k + RB < NBLK) zero_block_row(kslot);
      __syncthreads();
      // ---- trailing update of the window, look-ahead factorisation of the next diagonal tile, window slide ----
      const int ntile = nl * (nl + 1) / 2;
      if (warp == 0) {
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
      } else {
        // the last warp first refills the freed block row (it is not touched by this step's updates)
        if (warp == NW - 1 && k + RB < NBLK) scatter_block_row(k + RB, kslot, lane, 32);
        constexpr int NTW = NW - 1;
        const int wrank = warp - 1;
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
      }
      __syncthreads();
    }
    if (bad && lane == 0) atomicOr(&status[pid], 1);
  }
}

// =====================================================================================================================
// k_patch_trisolve
// =====================================================================================================================
template <int RBMAX, int NW>
__global__ void __launch_bounds__(32 * NW, 1)
k_patch_trisolve(const int *__restrict__ patch_ids, int n_work, const double *__restrict__ Lrec,
                 double *__restrict__ Xbuf, SplitLayout lay, int *work_counter) {
  constexpr int NT = 32 * NW;
  constexpr int NC = 8 * NW;
  constexpr int LSTEP = RBMAX * 64;
  constexpr int NSTG = kTriStages;
  extern __shared__ __align__(128) double smem[];
  double *sRing = smem;                                      // [NSTG][LSTEP]
  uint64_t *sFull = reinterpret_cast<uint64_t *>(sRing + NSTG * LSTEP);   // [NSTG]
  uint64_t *sEmpty = sFull + NSTG;                            // [NSTG]
  int *sRowPk = reinterpret_cast<int *>(sEmpty + NSTG);       // [nip_max] packed node coords of each interior dof
  int *sColCell = sRowPk + lay.nip_max;                       // [NC] packed cell coords of each coarse column (or -1)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int tpos = 8 * t + 2 * (g & 3) + (g >> 2);           // T^T fragment position, k-step 0 (+32 for k-step 1)

  if (tid == 0) {
    for (int s = 0; s < NSTG; ++s) {
      mbar_init(sFull + s, 1);
      mbar_init(sEmpty + s, NW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  uint32_t rec_base = 0;   // records consumed so far by this CTA (ring position and barrier phases follow from it)

  __shared__ int sNextWork;
  SLOD_WORK_LOOP(w, n_work, work_counter, sNextWork) {
    fetch_work_item(w, work_counter, &sNextWork);
    const int pid = patch_ids[w];
    const Geom geo = make_geom(cP, pid);
    const int Ni = geo.Ni, bw = geo.bw, ncd = geo.Ncd;
    const int NBLK = (Ni + 7) >> 3;
    int RB = (bw + 8 + 7) >> 3;
    if (RB > RBMAX) RB = RBMAX;
    double *X = Xbuf + (size_t)w * lay.x_stride;
    const double *rec = Lrec + (size_t)w * lay.rec_stride;
    __syncthreads();   // the tables of the previous patch are dead
    for (int r = tid; r < 8 * NBLK; r += NT) {
      int pk = 0;
      if (r < Ni) {
        int a[3];
        interior_coords(geo, r, a);
        pk = a[0] | (a[1] << 5) | (a[2] << 10) | (1 << 31);
      }
      sRowPk[r] = pk;
    }
    for (int col = tid; col < NC; col += NT) {
      int v = -1;
      if (col < ncd) {
        int kc[3];
        col_to_cell(cP, geo, col, kc);
        v = kc[0] | (kc[1] << 5) | (kc[2] << 10);
      }
      sColCell[col] = v;
    }
    __syncthreads();
    const int mycc0 = sColCell[8 * warp + 2 * t], mycc1 = sColCell[8 * warp + 2 * t + 1];
    const int n = cP.n;
    const double pw = cP.pw;
    // right-hand-side tile (C layout) of block blk for this warp's columns: P_i entries from the tables
    auto rhs_tile = [&](int blk, double &c0, double &c1) {
      c0 = c1 = 0.0;
      const int pk = sRowPk[8 * blk + g];
      if (pk >= 0) return;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int cc = h ? mycc1 : mycc0;
        if (cc < 0) continue;
        double wgt = pw;
        bool in = true;
#pragma unroll
        for (int x = 0; x < 3; ++x) {
          const int tt = ((pk >> (5 * x)) & 31) - n * ((cc >> (5 * x)) & 31);
          if (tt < 0 || tt > n) in = false;
          if (tt != 0 && tt != n) wgt *= 2.0;
        }
        if (in) { if (h == 0) c0 = wgt; else c1 = wgt; }
      }
    };

    // ---- record stream of this patch: forward steps 0 .. NBLK-1, then backward steps NBLK-1 .. 0 ----
    const int total = 2 * NBLK;
    int issued = 0;   // producer state (thread 0 only)
    auto step_of = [&](int i) { return i < NBLK ? i : 2 * NBLK - 1 - i; };
    auto top_up = [&](int upto) {   // thread 0: keep the ring filled up to record `upto` (exclusive)
      if (upto > total) upto = total;
      while (issued < upto) {
        const uint32_t r = rec_base + issued;
        const int slot = r % NSTG;
        const uint32_t use = r / NSTG;
        if (use > 0) mbar_wait(sEmpty + slot, (use - 1) & 1);   // every warp is through with the previous occupant
        const int k = step_of(issued);
        int nl = NBLK - 1 - k;
        if (nl > RB - 1) nl = RB - 1;
        const uint32_t bytes = 8u * 64u * (1 + nl);
        mbar_expect_tx(sFull + slot, bytes);
        bulk_g2s(sRing + slot * LSTEP, rec + (size_t)k * LSTEP, bytes, sFull + slot);
        ++issued;
      }
    };
    if (tid == 0) top_up(NSTG - 1);

    // =============================== forward substitution  L Y = P_i ===============================
    double cr[RBMAX][2];   // cr[off]: right-hand-side tile of block k + off (rotating register window)
#pragma unroll
    for (int off = 0; off < RBMAX; ++off) {
      cr[off][0] = cr[off][1] = 0.0;
      if (off < RB && off < NBLK) rhs_tile(off, cr[off][0], cr[off][1]);
    }
    for (int k = 0; k < NBLK; ++k) {
      if (tid == 0) top_up(k + NSTG - 1);
      const uint32_t r = rec_base + k;
      const int slot = r % NSTG;
      int nl = NBLK - 1 - k;
      if (nl > RB - 1) nl = RB - 1;
      // independent of the record: the B fragments of the current right-hand-side tile, the tile entering the window
      const double b0 = c_to_b(cr[0][0], cr[0][1], lane, 0), b1 = c_to_b(cr[0][0], cr[0][1], lane, 1);
      double n0 = 0.0, n1 = 0.0;
      if (k + RB < NBLK) rhs_tile(k + RB, n0, n1);
      mbar_wait(sFull + slot, (r / NSTG) & 1);
      const double2 *F = reinterpret_cast<const double2 *>(sRing + slot * LSTEP);
      double y0 = 0.0, y1 = 0.0;
      {
        const double2 li = F[lane];
        dmma884(y0, y1, li.x, b0);
        dmma884(y0, y1, li.y, b1);
      }
      *reinterpret_cast<double2 *>(X + (size_t)(8 * k + g) * lay.ldx + 8 * warp + 2 * t) = make_double2(y0, y1);
      const double yb0 = -c_to_b(y0, y1, lane, 0), yb1 = -c_to_b(y0, y1, lane, 1);
#pragma unroll
      for (int o4 = 1; o4 < RBMAX; o4 += 4) {
        double2 a[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (o4 + u < RBMAX && o4 + u <= nl) a[u] = F[32 * (o4 + u) + lane];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (o4 + u < RBMAX && o4 + u <= nl) dmma884(cr[o4 + u][0], cr[o4 + u][1], a[u].x, yb0);
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (o4 + u < RBMAX && o4 + u <= nl) dmma884(cr[o4 + u][0], cr[o4 + u][1], a[u].y, yb1);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(sEmpty + slot);
      // the register window moves up by one block
#pragma unroll
      for (int off = 0; off < RBMAX - 1; ++off) {
        cr[off][0] = (off == RB - 1) ? n0 : cr[off + 1][0];
        cr[off][1] = (off == RB - 1) ? n1 : cr[off + 1][1];
      }
      if (RB == RBMAX) { cr[RBMAX - 1][0] = n0; cr[RBMAX - 1][1] = n1; }
    }

    // =============================== backward substitution  L^T X = Y ===============================
    // xr[off]: NEGATED B fragments of the solved block k + off
    double xr[RBMAX][2];
#pragma unroll
    for (int off = 0; off < RBMAX; ++off) xr[off][0] = xr[off][1] = 0.0;
    auto y_tile = [&](int k) -> double2 {
      return (k >= 0) ? *reinterpret_cast<const double2 *>(X + (size_t)(8 * k + g) * lay.ldx + 8 * warp + 2 * t)
                      : make_double2(0.0, 0.0);
    };
    double2 ya = y_tile(NBLK - 1), yb = y_tile(NBLK - 2);
    for (int k = NBLK - 1; k >= 0; --k) {
      const int i = 2 * NBLK - 1 - k;   // record index
      if (tid == 0) top_up(i + NSTG - 1);
      const uint32_t r = rec_base + i;
      const int slot = r % NSTG;
      int nl = NBLK - 1 - k;
      if (nl > RB - 1) nl = RB - 1;
      const double2 yn = y_tile(k - 2);
      mbar_wait(sFull + slot, (r / NSTG) & 1);
      const double *F = sRing + slot * LSTEP;
      double c0 = ya.x, c1 = ya.y, e0 = 0.0, e1 = 0.0;   // two accumulation chains
#pragma unroll
      for (int o2 = 1; o2 < RBMAX; o2 += 2) {
        // A = Lp^T : A[m = g][kk = 4 j + t] = Lp[4 j + t][g]
        double a0[2], a1[2];
#pragma unroll
        for (int u = 0; u < 2; ++u)
          if (o2 + u < RBMAX && o2 + u <= nl) {
            a0[u] = F[64 * (o2 + u) + tpos];
            a1[u] = F[64 * (o2 + u) + 32 + tpos];
          }
        if (o2 <= nl) dmma884(c0, c1, a0[0], xr[o2][0]);
        if (o2 + 1 < RBMAX && o2 + 1 <= nl) dmma884(e0, e1, a0[1], xr[o2 + 1][0]);
        if (o2 <= nl) dmma884(c0, c1, a1[0], xr[o2][1]);
        if (o2 + 1 < RBMAX && o2 + 1 <= nl) dmma884(e0, e1, a1[1], xr[o2 + 1][1]);
      }
      c0 += e0;
      c1 += e1;
      // X_k = Linv_k^T T
      const double tb0 = c_to_b(c0, c1, lane, 0), tb1 = c_to_b(c0, c1, lane, 1);
      double x0 = 0.0, x1 = 0.0;
      dmma884(x0, x1, F[tpos], tb0);
      dmma884(x0, x1, F[32 + tpos], tb1);
      __syncwarp();
      if (lane == 0) mbar_arrive(sEmpty + slot);
      *reinterpret_cast<double2 *>(X + (size_t)(8 * k + g) * lay.ldx + 8 * warp + 2 * t) = make_double2(x0, x1);
      const double nb0 = -c_to_b(x0, x1, lane, 0), nb1 = -c_to_b(x0, x1, lane, 1);
#pragma unroll
      for (int off = RBMAX - 1; off >= 2; --off) { xr[off][0] = xr[off - 1][0]; xr[off][1] = xr[off - 1][1]; }
      xr[1][0] = nb0;
      xr[1][1] = nb1;
      ya = yb;
      yb = yn;
    }
    rec_base += (uint32_t)total;
  }
}

}  // namespace slod
