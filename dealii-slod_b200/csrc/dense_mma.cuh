// k_patch_dense_mma : M, M^{-1}, BD = (S_b X - P_b) M^{-1} and G = BD^T BD.
//
// One CTA (NTILE warps) per patch.
//  * M = P_i^T X / H^d is accumulated straight into a register tile (4 x NTILE/2 entries per thread) from a
//    per-patch table of the interior X rows of every coarse cell;
//  * M^{-1}: blocked Gauss-Jordan (8 x 8 pivot blocks) with DMMA tile updates in shared memory (replaces
//    FullMatrix::gauss_jordan, source/LOD.cc:553);
//  * BD and the Gram matrix use FP64 mma.sync tiles: warp w owns the coarse-column tile [8w, 8w+8) of BD and one half
//    of the lower-triangle tiles of the tile-row pair (r, NTILE-1-r), r = w mod NTILE/2 (every pair has NTILE + 1
//    tiles, so the NTILE (NTILE + 1) / 2 tiles of the triangle are spread evenly and each is computed once); the
//    accumulators stay in registers while the boundary rows stream through shared memory in tiles of 32.
// Included by kernels.cu.
#pragma once

namespace slod {

constexpr int kDTB = 32;    // boundary rows per tile

// one node layer of X (contiguous rows) -> shared memory as four bulk copies on one barrier (several requests in flight)
__device__ __forceinline__ void stage_layer(double *dst, const double *src, size_t doubles, uint64_t *bar) {
  mbar_expect_tx(bar, (uint32_t)(doubles * sizeof(double)));
  const size_t piece = ((doubles + 3) / 4 + 1) & ~(size_t)1;   // even number of doubles: 16-byte granules
  for (size_t off = 0; off < doubles; off += piece) {
    const size_t cnt = (doubles - off < piece) ? doubles - off : piece;
    bulk_g2s(dst + off, src + off, (uint32_t)(cnt * sizeof(double)), bar);
  }
}

template <int NTILE>
__global__ void __launch_bounds__(32 * NTILE, 1)
k_patch_dense_mma(const int *__restrict__ patch_ids, int n_work, const double *__restrict__ d_coef,
                  const double *__restrict__ Xbuf, const double *__restrict__ Wbuf, double *__restrict__ Minv_out,
                  double *__restrict__ G_out, double *__restrict__ diag, int *__restrict__ status, DenseLayout lay,
                  int *work_counter) {
  constexpr int NT = 32 * NTILE;
  constexpr int NC = 8 * NTILE;   // padded coarse dimension
  constexpr int LDM = NC + 4;     // LDM % 16 == 4 : conflict-free fragment loads
  constexpr int TW = NTILE / 2;   // register tile: 4 rows x TW columns per thread; thread tx owns columns tx + 16 j
                                  // (interleaved: the pivot-row reads of a half warp are 16 consecutive doubles)
  extern __shared__ double smem[];
  double *sM = smem;                      // [NC][LDM]   M, then M^{-1} for the mma B operand
  double *sT = sM + NC * LDM;             // [kDTB][LDM] W tile, then BD tile (even tiles); scratch of the Gauss-Jordan
  double *sT2 = sT + kDTB * LDM;          // [kDTB][LDM] the same for the odd tiles
  int *sMList = (int *)(sT2 + kDTB * LDM);  // [NC][32] X rows under each coarse row: (xrow << 2 | log2 weight) or -1
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int ty = tid >> 4, tx = tid & 15;
  // coarse rows of this thread in the M accumulation: ty, ty + NT/16, ... (interleaved: rows near the patch boundary
  // have fewer interior nodes under them, consecutive rows would leave some warps with much less to gather)
  constexpr int RSTR = 2 * NTILE;

  unsigned long long table_key = ~0ull;   // shape key (geom.h) of the patch the gather table was built for
  __shared__ int sNextWork;
  // staged accumulation of M (3-D scalar problems): two node layers of X in flight as bulk copies
  constexpr int KZ = 5;                     // coarse cells per axis of a patch the staged path handles (l <= 2)
  constexpr size_t kStageCap = (size_t)NC * LDM + 2 * (size_t)kDTB * LDM + (NC * 32) / 2;   // doubles from sM on
  __shared__ __align__(8) uint64_t sBar[2];
  uint32_t bar_phase = 0;                   // bit b: phase of sBar[b] the next wait expects (uniform over the CTA)
  if (tid == 0) {
    mbar_init(&sBar[0], 1);
    mbar_init(&sBar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  SLOD_WORK_LOOP(w, n_work, work_counter, sNextWork) {
    fetch_work_item(w, work_counter, &sNextWork);
    const int pid = patch_ids[w];
    const Geom geo = make_geom(cP, pid);
    const int ncd = geo.Ncd, s = cP.s;
    const double *X = Xbuf + (size_t)w * lay.x_stride;
    const size_t layer_doubles = (size_t)geo.q[0] * geo.q[1] * lay.ldx;
    const bool staged = (TW % 2 == 0) && cP.dim == 3 && s == 1 && geo.m[0] * geo.m[1] <= NT / 16 && geo.m[2] <= KZ &&
                        2 * layer_doubles <= kStageCap && (smem_u32(sM) & 15u) == 0 && lay.ldx >= NC;
    const bool rebuild = !staged && (shape_key(geo) != table_key);   // the table is integer geometry: per shape, not per patch
    table_key = staged ? ~0ull : shape_key(geo);                     // the staging buffers overlay the table
    __syncthreads();
    PH_DECL
    if (staged && tid == 0) {
      // the previous patch used the region through the generic proxy (everybody is past the barrier above)
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      stage_layer(sM, X, layer_doubles, &sBar[0]);
      if (geo.q[2] > 1) stage_layer(sM + layer_doubles, X + layer_doubles, layer_doubles, &sBar[1]);
    }
    // ---- table: interior X rows (and weights 1,2,4,8) under every coarse row; one (row, local node) pair per thread,
    // compacted per row by ballot. ----
    const int npc = cP.n + 1;
    const int nloc = (cP.dim == 3) ? npc * npc * npc : npc * npc;   // <= 27 (host guarantees)
    if (rebuild)
    for (int idx = tid; idx < NC * 32; idx += NT) {
      const int row = idx >> 5, l = idx & 31;
      int e = -1;
      if (row < ncd && l < nloc) {
        const int comp = row % s;
        int k[3];
        if (lay.zmajor) zcol_to_cell(geo, row / s, k);
        else col_to_cell(cP, geo, row / s, k);
        int tt[3] = {l % npc, (l / npc) % npc, (cP.dim == 3) ? l / (npc * npc) : 0};
        int a[3] = {k[0] * cP.n + tt[0], k[1] * cP.n + tt[1], (cP.dim == 3) ? k[2] * cP.n + tt[2] : 0};
        if (node_class(cP, geo, a) == 0) {
          int lg = 0;
          _Pragma("unroll") for (int x = 0; x < 3; ++x) if (x < cP.dim)
            if (tt[x] != 0 && tt[x] != cP.n) ++lg;
          e = ((interior_index(geo, a) * s + comp) << 2) | lg;
        }
      }
      // a warp = one row: compact the valid entries to the front (slot order kept), count in slot 31
      const unsigned mask = __ballot_sync(0xffffffffu, e >= 0);
      if (e >= 0) sMList[(row << 5) + __popc(mask & ((1u << l) - 1u))] = e;
      if (l == 31) sMList[idx] = __popc(mask);
    }
    __syncthreads();

    PH(0)
    if (staged) {
      // ---- M = P_i^T X / H^d, X read exactly once: the interior rows of one node layer (z fixed) are contiguous in X
      // and arrive as one bulk copy (two layers in flight).  Thread = (cell column (kx, ky), 8 coarse columns); per
      // layer it adds the (at most 3 x 3) interior nodes of its cell column with the x-y projection weights and hands
      // the sum to the (at most two) cells of the column that contain the layer. ----
      const int cxy = tid >> 4, tx2 = 2 * (tid & 15);
      const int kx = cxy / geo.m[1], ky = cxy - kx * geo.m[1];
      const bool active = cxy < geo.m[0] * geo.m[1];
      const int nn = cP.n, q0 = geo.q[0];
      double acc[KZ][TW];
#pragma unroll
      for (int kz = 0; kz < KZ; ++kz)
#pragma unroll
        for (int j = 0; j < TW; ++j) acc[kz][j] = 0.0;
      for (int zi = 0; zi < geo.q[2]; ++zi) {
        const int b = zi & 1;
        mbar_wait(&sBar[b], (bar_phase >> b) & 1u);
        bar_phase ^= 1u << b;
        const double *L = sM + (size_t)b * layer_doubles;
        double t2[TW];
#pragma unroll
        for (int j = 0; j < TW; ++j) t2[j] = 0.0;
        if (active) {
          for (int ty_ = 0; ty_ <= nn; ++ty_) {
            const int ay = nn * ky + ty_;
            if (ay < 1 || ay > geo.p[1] - 2) continue;
            const double wy = (ty_ != 0 && ty_ != nn) ? 2.0 : 1.0;
            for (int tx_ = 0; tx_ <= nn; ++tx_) {
              const int ax = nn * kx + tx_;
              if (ax < 1 || ax > geo.p[0] - 2) continue;
              const double wxy = (tx_ != 0 && tx_ != nn) ? 2.0 * wy : wy;
              const double2 *src = reinterpret_cast<const double2 *>(L + (size_t)((ay - 1) * q0 + (ax - 1)) * lay.ldx + tx2);
#pragma unroll
              for (int j = 0; j < TW / 2; ++j) {
                const double2 v = src[16 * j];
                t2[2 * j] += wxy * v.x;
                t2[2 * j + 1] += wxy * v.y;
              }
            }
          }
          const int az = zi + 1;
#pragma unroll
          for (int kz = 0; kz < KZ; ++kz)
            if (az >= nn * kz && az <= nn * kz + nn) {
              const double wz = (az != nn * kz && az != nn * kz + nn) ? 2.0 : 1.0;
#pragma unroll
              for (int j = 0; j < TW; ++j) acc[kz][j] += wz * t2[j];
            }
        }
        __syncthreads();   // everybody is done with buffer b
        if (tid == 0 && zi + 2 < geo.q[2])
          stage_layer(sM + (size_t)b * layer_doubles, X + (size_t)(zi + 2) * layer_doubles, layer_doubles, &sBar[b]);
      }
      // all copies have landed and been consumed: M replaces the staging buffers (padding rows/cols: identity)
      const double scale = cP.pw / cP.Hd;
      if (active) {
#pragma unroll
        for (int kz = 0; kz < KZ; ++kz)
          if (kz < geo.m[2]) {
            const int kc[3] = {kx, ky, kz};
            const int row = lay.zmajor ? zcell_to_col(geo, kc) : cell_to_col(cP, geo, kc);
#pragma unroll
            for (int j = 0; j < TW; ++j) {
              const int col = tx2 + 32 * (j >> 1) + (j & 1);
              sM[row * LDM + col] = (col < ncd) ? acc[kz][j] * scale : 0.0;
            }
          }
      }
      for (int idx = tid; idx < (NC - ncd) * NC; idx += NT) {
        const int row = ncd + idx / NC, col = idx % NC;
        sM[row * LDM + col] = (row == col) ? 1.0 : 0.0;
      }
    } else {
    // ---- M = P_i^T X / H^d into the register tile (padding rows/cols: identity) ----
    double m[4][TW];
    {
      const double scale = cP.pw / cP.Hd;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < TW; ++j) m[i][j] = 0.0;
        const int row = ty + RSTR * i;
        const int cnt = sMList[row * 32 + 31];
        for (int l = 0; l < cnt; ++l) {
          const int e = sMList[row * 32 + l];
          const double wgt = (double)(1 << (e & 3));
          const double *xr = X + (size_t)(e >> 2) * lay.ldx + tx;
#pragma unroll
          for (int j = 0; j < TW; ++j) m[i][j] += wgt * xr[16 * j];
        }
#pragma unroll
        for (int j = 0; j < TW; ++j) {
          m[i][j] *= scale;
          if (row >= ncd || tx + 16 * j >= ncd) m[i][j] = (row == tx + 16 * j) ? 1.0 : 0.0;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < TW; ++j) sM[(ty + RSTR * i) * LDM + tx + 16 * j] = m[i][j];
    }
    PH(1)
    // ---- M^{-1}: blocked Gauss-Jordan (8 x 8 pivot blocks, no pivoting: M is SPD) with DMMA tile updates in
    // shared memory; replaces FullMatrix::gauss_jordan (source/LOD.cc:553).  Step K:
    //   Pinv = M_KK^{-1};  R_KJ = Pinv M_KJ;  M_IJ -= M_IK R_KJ (I, J != K);  M_IK = -M_IK Pinv;  M_KJ = R_KJ;  M_KK = Pinv.
    // Warp w owns block row I = w; the pivot block of the next step is inverted one step ahead (below).
    {
      __syncthreads();
      int badpiv = 0;
      const int nblk = (ncd + 7) >> 3;
      double *sPinvB = sT;           // [2][64] Pinv of step K in buffer K & 1 (row-major); sT is free until the W tiles
      double *sLi = sT + 128;        // [64] scratch: L^{-1}
      double *sTile = sT + 192;      // [64] the next pivot tile, row-major (published by its owner, inverted by warp K)
      double *sRowB = sT + 256;      // [2][8][LDM] block row K (buffer K & 1): first as it is, then multiplied with Pinv
      auto invert_pivot = [&](double *Pinv) {  // one warp: Pinv = (L^{-1})^T L^{-1} of the tile in sTile
        badpiv |= chol8_inv_reg(sTile, 8, lane, sLi);   // shuffle-free: the shortest dependency chain (solve_mma.cuh)
        __syncwarp();
        double p0 = 0.0, p1 = 0.0;
        const double f0 = sLi[t * 8 + g], f1 = sLi[(4 + t) * 8 + g];  // A[m][k] = Li[k][m] and B[k][n] = Li[k][n]
        dmma884(p0, p1, f0, f0);
        dmma884(p0, p1, f1, f1);
        *reinterpret_cast<double2 *>(Pinv + g * 8 + 2 * t) = make_double2(p0, p1);
      };
      // Warp I keeps block row I in registers for the whole elimination (NTILE tiles in C-fragment layout, 2 doubles
      // each).  With the rows in shared memory every tile of the rank-8 sweep cost a 16-byte load and store (2-way
      // bank conflicts: the row stride suits the 8-byte fragment loads of the BD phase) on top of its B operand, and
      // the sweep was bound by shared-memory wavefronts, not by the tensor pipe.  Only the pivot row passes through
      // shared memory: its owner writes it at the end of the previous step, every warp multiplies one tile with Pinv
      // in place, the sweeps read their B operands from it, and the owner takes it back at the end of the step.
      double2 c[NTILE];
      const bool rowwarp = warp < nblk;
      const int frag = g * LDM + 2 * t;   // C-fragment position of this lane inside a block row
#pragma unroll
      for (int J = 0; J < NTILE; ++J)
        c[J] = (rowwarp && J < nblk) ? *reinterpret_cast<const double2 *>(sM + 8 * warp * LDM + frag + 8 * J)
                                     : make_double2(0.0, 0.0);
      if (warp == 0) {
        *reinterpret_cast<double2 *>(sTile + g * 8 + 2 * t) = c[0];
#pragma unroll
        for (int J = 0; J < NTILE; ++J)
          if (J < nblk) *reinterpret_cast<double2 *>(sRowB + frag + 8 * J) = c[J];
        __syncwarp();
        invert_pivot(sPinvB);
      }
      __syncthreads();
      for (int K = 0; K < nblk; ++K) {
        PH(2)
        const double *sPinv = sPinvB + (K & 1) * 64;
        double *sRow = sRowB + (K & 1) * 8 * LDM;
        const bool have_next = K + 1 < nblk;
        // ---- R_KJ = Pinv M_KJ, in place in the row buffer; warp J handles tile J ----
        if (rowwarp && warp != K) {
          const int J = warp;
          const double b0 = sRow[t * LDM + 8 * J + g], b1 = sRow[(4 + t) * LDM + 8 * J + g];
          double r0_ = 0.0, r1_ = 0.0;
          dmma884(r0_, r1_, sPinv[g * 8 + t], b0);
          dmma884(r0_, r1_, sPinv[g * 8 + 4 + t], b1);
          __syncwarp();
          *reinterpret_cast<double2 *>(sRow + frag + 8 * J) = make_double2(r0_, r1_);
        }
        __syncthreads();
        PH(9)
        if (rowwarp && warp != K) {
          // A fragments {(g, t), (g, 4 + t)} of the own column-K tile out of its C-fragment registers: element (g, col)
          // sits in lane 4 g + (col >> 1), slot col & 1
          double2 ck = c[0];
#pragma unroll
          for (int J = 1; J < NTILE; ++J) ck = (J == K) ? c[J] : ck;
          const double x0 = __shfl_sync(0xffffffffu, ck.x, 4 * g + (t >> 1)), y0 = __shfl_sync(0xffffffffu, ck.y, 4 * g + (t >> 1));
          const double x1 = __shfl_sync(0xffffffffu, ck.x, 4 * g + 2 + (t >> 1)), y1 = __shfl_sync(0xffffffffu, ck.y, 4 * g + 2 + (t >> 1));
          const double a0 = -((t & 1) ? y0 : x0), a1 = -((t & 1) ? y1 : x1);
          const double *br0 = sRow + t * LDM + g, *br1 = br0 + 4 * LDM;
          const bool is_next = have_next && warp == K + 1;
          if (is_next) {
            // the next pivot block first: publish it for warp K's look-ahead inversion
            double2 cn = c[0];
#pragma unroll
            for (int J = 1; J < NTILE; ++J) cn = (J == K + 1) ? c[J] : cn;
            dmma884(cn.x, cn.y, a0, br0[8 * (K + 1)]);
            dmma884(cn.x, cn.y, a1, br1[8 * (K + 1)]);
#pragma unroll
            for (int J = 0; J < NTILE; ++J) c[J] = (J == K + 1) ? cn : c[J];
            *reinterpret_cast<double2 *>(sTile + g * 8 + 2 * t) = cn;
            __syncwarp();
            asm volatile("bar.arrive 1, 64;" ::: "memory");
          }
          const int jskip = is_next ? K + 1 : -1;
          // ---- M_IJ -= M_IK R_KJ over the row, four tiles in flight ----
#pragma unroll
          for (int J0 = 0; J0 < NTILE; J0 += 4) {
            if (J0 < nblk) {
              double b0[4], b1[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                b0[u] = br0[8 * (J0 + u)];
                b1[u] = br1[8 * (J0 + u)];
              }
#pragma unroll
              for (int u = 0; u < 4; ++u)
                if (J0 + u != K && J0 + u != jskip && J0 + u < nblk) dmma884(c[J0 + u].x, c[J0 + u].y, a0, b0[u]);
#pragma unroll
              for (int u = 0; u < 4; ++u)
                if (J0 + u != K && J0 + u != jskip && J0 + u < nblk) dmma884(c[J0 + u].x, c[J0 + u].y, a1, b1[u]);
            }
          }
          // ---- column block: M_IK = -M_IK Pinv ----
          double2 newcol = make_double2(0.0, 0.0);
          dmma884(newcol.x, newcol.y, a0, sPinv[t * 8 + g]);
          dmma884(newcol.x, newcol.y, a1, sPinv[(4 + t) * 8 + g]);
#pragma unroll
          for (int J = 0; J < NTILE; ++J) c[J] = (J == K) ? newcol : c[J];
          if (is_next) {   // the pivot row of the next step goes to the other row buffer
            double *nrow = sRowB + ((K + 1) & 1) * 8 * LDM + frag;
#pragma unroll
            for (int J = 0; J < NTILE; ++J)
              if (J < nblk) *reinterpret_cast<double2 *>(nrow + 8 * J) = c[J];
          }
        } else if (warp == K) {
          // the look-ahead: invert the next pivot block into the other Pinv buffer (the row registers are dead here) ...
          if (have_next) {
            asm volatile("bar.sync 1, 64;" ::: "memory");
            invert_pivot(sPinvB + ((K + 1) & 1) * 64);
          }
          // ... then take the row back: M_KJ = R_KJ, M_KK = Pinv
#pragma unroll
          for (int J = 0; J < NTILE; ++J)   // every register is overwritten: nothing of the row stays live across the inversion
            c[J] = (J < nblk) ? *reinterpret_cast<const double2 *>(sRow + frag + 8 * J) : make_double2(0.0, 0.0);
          const double2 pk = *reinterpret_cast<const double2 *>(sPinv + g * 8 + 2 * t);
#pragma unroll
          for (int J = 0; J < NTILE; ++J) c[J] = (J == K) ? pk : c[J];
        }
#ifdef SLOD_PHASE_CLOCKS
        if (ph_on) { const long long c_ = clock64(); ph_acc[(warp == K) ? 4 : 11] += c_ - ph_last; ph_last = c_; }
#endif
        __syncthreads();
        PH(10)
      }
      if (rowwarp) {
#pragma unroll
        for (int J = 0; J < NTILE; ++J)
          if (J < nblk) *reinterpret_cast<double2 *>(sM + 8 * warp * LDM + frag + 8 * J) = c[J];
      }
      __syncthreads();
      if (badpiv && lane == 0) atomicOr(&status[pid], 2);
      double *Mo = Minv_out + (size_t)w * lay.m_stride;
      for (int idx = tid; idx < ncd * NC; idx += NT) {
        const int row = idx / NC, col = idx - row * NC;
        if (col < ncd) {
          Mo[row * ncd + col] = sM[row * LDM + col];
        } else {
          sM[row * LDM + col] = 0.0;   // padding columns / rows must not contribute to BD
        }
      }
      for (int idx = tid; idx < (NC - ncd) * NC; idx += NT) sM[(ncd + idx / NC) * LDM + idx % NC] = 0.0;
    }
    PH(2)
    if (!geo.slod) continue;
    __syncthreads();

    // ---- BD tiles and Gram accumulation ----
    constexpr int NGA = (NTILE + 2) / 2;   // Gram tiles per warp: half of the NTILE + 1 tiles of a tile-row pair
    double gacc[NGA][2];
#pragma unroll
    for (int e = 0; e < NGA; ++e) gacc[e][0] = gacc[e][1] = 0.0;
    int nbd = 0;   // patch-boundary dofs (id 99): nodes on a patch side that is not part of the domain boundary
    {
      int cnt_all = 1, cnt_nob = 1;
      for (int a = 0; a < cP.dim; ++a) {
        cnt_all *= geo.p[a];
        cnt_nob *= geo.p[a] - (geo.domlo[a] ? 0 : 1) - (geo.domhi[a] ? 0 : 1);
      }
      nbd = s * (cnt_all - cnt_nob);
    }
    const int ksteps = (ncd + 3) >> 2;
    const int I1 = warp % (NTILE / 2), I2 = NTILE - 1 - I1;
    const int e_base = (warp / (NTILE / 2)) * NGA;   // first entry of this warp in the pair's list of NTILE + 1 tiles
    int gcol[NGA];          // B column (doubles) of the el-th owned tile
    unsigned gfirst = 0;    // bit el: the tile belongs to tile row I1 (else I2)
#pragma unroll
    for (int el = 0; el < NGA; ++el) {
      const int e = e_base + el;
      const bool first = (e <= I1);
      gcol[el] = (e > NTILE) ? 0 : 8 * (first ? e : e - (I1 + 1));
      gfirst |= (first ? 1u : 0u) << el;
    }
    // W = S_b X - P_b comes from k_patch_flux, zero padded to whole tiles.  The tiles alternate between two buffers:
    // a thread moves 64 bytes of a tile with cp.async (NT * 8 = 32 * NC doubles), tile i + 1 is in flight while the
    // tensor phases of tile i run, and BD replaces W in the tile's own buffer.
    const double *Wp = Wbuf + (size_t)w * lay.w_stride;
    const int wrow = tid / (NC / 8), wcol = 8 * (tid % (NC / 8));
    auto issue_w = [&](int t0, double *buf) {
      const double *src = Wp + (size_t)(t0 + wrow) * NC + wcol;
      const uint32_t dst = smem_u32(buf + wrow * LDM + wcol);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16 * q), "l"(src + 2 * q) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (nbd > 0) issue_w(0, sT);
    for (int t0 = 0, it = 0; t0 < nbd; t0 += kDTB, ++it) {
      PH(3)
      double *sW = (it & 1) ? sT2 : sT;
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();   // tile `it` is complete and visible; nobody reads the other buffer (Gram of tile it - 1) any more
      if (t0 + kDTB < nbd) issue_w(t0 + kDTB, (it & 1) ? sT : sT2);
      PH(5)
      // BD tile = W tile * Minv : warp owns 8 columns, 4 row tiles
      double bd[4][2];
#pragma unroll
      for (int i = 0; i < 4; ++i) bd[i][0] = bd[i][1] = 0.0;
      for (int kk = 0; kk < ksteps; ++kk) {
        const double bf = sM[(4 * kk + t) * LDM + 8 * warp + g];
#pragma unroll
        for (int i = 0; i < 4; ++i) dmma884(bd[i][0], bd[i][1], sW[(8 * i + g) * LDM + 4 * kk + t], bf);
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *reinterpret_cast<double2 *>(sW + (8 * i + g) * LDM + 8 * warp + 2 * t) = make_double2(bd[i][0], bd[i][1]);
      __syncthreads();
      PH(6)
      // G += BD^T BD on the owned lower-triangle tiles.  One pointer per tile (its B column, fixed for the patch), the
      // rows advance by compile-time offsets: a tile step is one load, the select of its A operand and the tensor
      // instruction -- with four warps per scheduler the phase is bound by the issue rate as soon as a tile step
      // costs more than four instructions.  A warp's unused ninth slot recomputes tile 0 into an accumulator that is
      // never stored.
      {
        const double *rp = sW + t * LDM + g;
        const double *pa1 = rp + 8 * I1, *pa2 = rp + 8 * I2;
        const double *pb[NGA];
#pragma unroll
        for (int el = 0; el < NGA; ++el) pb[el] = rp + gcol[el];
#pragma unroll
        for (int jj = 0; jj < kDTB / 4; ++jj) {
          const double a1 = pa1[4 * jj * LDM], a2 = pa2[4 * jj * LDM];
#pragma unroll
          for (int el = 0; el < NGA; ++el)
            dmma884(gacc[el][0], gacc[el][1], ((gfirst >> el) & 1) ? a1 : a2, pb[el][4 * jj * LDM]);
        }
      }
      PH(7)
    }
    {
      double *Go = G_out + (size_t)w * lay.m_stride;
#pragma unroll
      for (int el = 0; el < NGA; ++el) {
        const int e = e_base + el;
        if (e > NTILE) continue;
        const bool first = (e <= I1);
        const int I = first ? I1 : I2;
        const int J = first ? e : e - (I1 + 1);
        const int i = 8 * I + g, j = 8 * J + 2 * t;
        if (i < ncd) {
          if (j < ncd) { Go[i * ncd + j] = gacc[el][0]; if (I != J) Go[j * ncd + i] = gacc[el][0]; }
          if (j + 1 < ncd) { Go[i * ncd + j + 1] = gacc[el][1]; if (I != J) Go[(j + 1) * ncd + i] = gacc[el][1]; }
        }
      }
    }
    PH(8)
    PH_PRINT("dense")
  }
}

}  // namespace slod
