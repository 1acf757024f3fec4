// k_patch_dense_mma : M, M^{-1}, BD = (S_b X - P_b) M^{-1} and G = BD^T BD with FP64 mma.sync tiles.
// One CTA (NTILE warps) per patch; warp w owns the coarse-column tile [8w, 8w+8) of BD and two tile rows
// (w and NTILE-1-w, lower triangle only) of the Gram matrix, whose accumulators stay in registers while the
// boundary rows stream through shared memory in tiles of 32.  Included by kernels.cu.
#pragma once

namespace slod {

constexpr int kDTB = 32;  // boundary rows per tile

template <int NTILE>
__global__ void __launch_bounds__(32 * NTILE, 1)
k_patch_dense_mma(const int *__restrict__ patch_ids, int n_work, const double *__restrict__ d_coef,
                  const double *__restrict__ Xbuf, double *__restrict__ Minv_out, double *__restrict__ G_out,
                  double *__restrict__ diag, int *__restrict__ status, DenseLayout lay) {
  constexpr int NT = 32 * NTILE;
  constexpr int NC = 8 * NTILE;   // padded coarse dimension
  constexpr int LDM = NC + 4;     // LDM % 16 == 4 : conflict-free fragment loads
  extern __shared__ double smem[];
  double *sCoef = smem;
  double *sM = sCoef + lay.coef_doubles;  // [NC][LDM]
  double *sT = sM + NC * LDM;             // [kDTB][LDM]
  double *sCol = sT + kDTB * LDM;         // [NC]
  double *sRow = sCol + NC;               // [NC]
  double *sArow = sRow + NC;              // [kDTB][54]
  int *sAnbr = (int *)(sArow + kDTB * 54);
  int *sBlist = sAnbr + kDTB * 54;
  __shared__ int sNb;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;

  for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
    const int pid = patch_ids[w];
    const Geom geo = make_geom(cP, pid);
    const int ncd = geo.Ncd, s = cP.s;
    const double *X = Xbuf + (size_t)w * lay.x_stride;
    __syncthreads();
    load_coef(geo, d_coef, sCoef);
    if (tid == 0) sNb = 0;
    for (int idx = tid; idx < NC * LDM; idx += NT) sM[idx] = 0.0;
    __syncthreads();

    // ---- M = P_i^T X / H^d ----
    const int npc = cP.n + 1;
    const int nloc = (cP.dim == 3) ? npc * npc * npc : npc * npc;
    const double scale = cP.pw / cP.Hd;
    for (int idx = tid; idx < ncd * NC; idx += NT) {
      const int row = idx / NC, col = idx % NC;
      if (col >= ncd) continue;
      const int comp = row % s;
      int k[3];
      col_to_cell(cP, geo, row / s, k);
      double acc = 0.0;
      for (int l = 0; l < nloc; ++l) {
        int tt[3] = {l % npc, (l / npc) % npc, (cP.dim == 3) ? l / (npc * npc) : 0};
        int a[3] = {k[0] * cP.n + tt[0], k[1] * cP.n + tt[1], (cP.dim == 3) ? k[2] * cP.n + tt[2] : 0};
        if (node_class(cP, geo, a) != 0) continue;
        double wgt = 1.0;
        for (int x = 0; x < cP.dim; ++x)
          if (tt[x] != 0 && tt[x] != cP.n) wgt *= 2.0;
        acc += wgt * X[(size_t)(interior_index(geo, a) * s + comp) * lay.ldx + col];
      }
      sM[row * LDM + col] = acc * scale;
    }
    if (geo.slod && tid < 32) {  // boundary dofs, ascending
      int count = 0;
      for (int base = 0; base < geo.nnodes; base += 32) {
        const int node = base + tid;
        bool isb = false;
        if (node < geo.nnodes) {
          int a[3];
          node_coords(geo, node, a);
          isb = (node_class(cP, geo, a) & 1) != 0;
        }
        const unsigned mask = __ballot_sync(0xffffffffu, isb);
        if (isb) {
          const int pos = (count + __popc(mask & ((1u << tid) - 1u))) * s;
          for (int c = 0; c < s; ++c) sBlist[pos + c] = node * s + c;
        }
        count += __popc(mask);
      }
      if (tid == 0) sNb = count * s;
    }
    __syncthreads();

    // ---- M^{-1}: in-place Gauss-Jordan sweeps ----
    for (int k = 0; k < ncd; ++k) {
      const double piv = sM[k * LDM + k];
      for (int j = tid; j < ncd; j += NT) {
        sCol[j] = sM[j * LDM + k];
        sRow[j] = sM[k * LDM + j] / piv;
      }
      if (tid == 0 && !(piv > 0.0)) atomicOr(&status[pid], 2);
      __syncthreads();
      for (int idx = tid; idx < ncd * NC; idx += NT) {
        const int i = idx / NC, j = idx % NC;
        if (j >= ncd) continue;
        double v;
        if (i == k && j == k) v = 1.0 / piv;
        else if (i == k) v = sRow[j];
        else if (j == k) v = -sCol[i] / piv;
        else v = sM[i * LDM + j] - sCol[i] * sRow[j];
        sM[i * LDM + j] = v;
      }
      __syncthreads();
    }
    {
      double *Mo = Minv_out + (size_t)w * lay.m_stride;
      for (int idx = tid; idx < ncd * ncd; idx += NT) Mo[idx] = sM[(idx / ncd) * LDM + idx % ncd];
    }
    if (!geo.slod) continue;

    // ---- BD tiles and Gram accumulation ----
    double gacc[NTILE + 1][2];
#pragma unroll
    for (int e = 0; e <= NTILE; ++e) gacc[e][0] = gacc[e][1] = 0.0;
    const int nbd = sNb;
    const int nst = (cP.dim == 3) ? 27 : 9;
    const int per_row = nst * s;
    const int ksteps = (ncd + 3) >> 2;
    const int I1 = warp, I2 = NTILE - 1 - warp;
    for (int t0 = 0; t0 < nbd; t0 += kDTB) {
      const int nt = min(kDTB, nbd - t0);
      for (int idx = tid; idx < nt * per_row; idx += NT) {
        const int rb = idx / per_row;
        int e = idx % per_row;
        const int cb = e % s;
        e /= s;
        int dl[3] = {e % 3 - 1, (e / 3) % 3 - 1, (cP.dim == 3) ? (e / 9 - 1) : 0};
        const int dof = sBlist[t0 + rb];
        int a[3];
        node_coords(geo, dof / s, a);
        int b[3] = {a[0] + dl[0], a[1] + dl[1], a[2] + dl[2]};
        bool ok = true;
        for (int x = 0; x < cP.dim; ++x) ok = ok && (b[x] >= 1 && b[x] <= geo.p[x] - 2);
        if (ok) {
          sAnbr[rb * 54 + (idx % per_row)] = interior_index(geo, b) * s + cb;
          sArow[rb * 54 + (idx % per_row)] = stiff_entry(cP, geo, sCoef, a, dl, dof % s, cb);
        } else {
          sAnbr[rb * 54 + (idx % per_row)] = -1;
        }
      }
      __syncthreads();
      // W tile = S_b X - P_b (zero padded to 32 x NC)
      for (int idx = tid; idx < kDTB * NC; idx += NT) {
        const int rb = idx / NC, col = idx % NC;
        double acc = 0.0;
        if (rb < nt && col < ncd) {
          const int dof = sBlist[t0 + rb];
          int a[3];
          node_coords(geo, dof / s, a);
          acc = -proj_entry(cP, geo, a, dof % s, col);
          for (int e = 0; e < per_row; ++e) {
            const int nb_ = sAnbr[rb * 54 + e];
            if (nb_ >= 0) acc += sArow[rb * 54 + e] * X[(size_t)nb_ * lay.ldx + col];
          }
        }
        sT[rb * LDM + col] = acc;
      }
      __syncthreads();
      // BD tile = W tile * Minv : warp owns 8 columns, 4 row tiles
      double bd[4][2];
#pragma unroll
      for (int i = 0; i < 4; ++i) bd[i][0] = bd[i][1] = 0.0;
      for (int kk = 0; kk < ksteps; ++kk) {
        const double bf = sM[(4 * kk + t) * LDM + 8 * warp + g];
#pragma unroll
        for (int i = 0; i < 4; ++i) dmma884(bd[i][0], bd[i][1], sT[(8 * i + g) * LDM + 4 * kk + t], bf);
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *reinterpret_cast<double2 *>(sT + (8 * i + g) * LDM + 8 * warp + 2 * t) = make_double2(bd[i][0], bd[i][1]);
      __syncthreads();
      // G += BD^T BD on the owned lower-triangle tiles
#pragma unroll
      for (int jj = 0; jj < kDTB / 4; ++jj) {
        const double *rowp = sT + (4 * jj + t) * LDM + g;
        const double a1 = rowp[8 * I1], a2 = rowp[8 * I2];
#pragma unroll
        for (int e = 0; e <= NTILE; ++e) {
          const bool first = (e <= I1);
          const int J = first ? e : e - (I1 + 1);
          if (!first && J > I2) continue;
          dmma884(gacc[e][0], gacc[e][1], first ? a1 : a2, rowp[8 * J]);
        }
      }
    }
    {
      double *Go = G_out + (size_t)w * lay.m_stride;
#pragma unroll
      for (int e = 0; e <= NTILE; ++e) {
        const bool first = (e <= I1);
        const int I = first ? I1 : I2;
        const int J = first ? e : e - (I1 + 1);
        if (J > I) continue;
        const int i = 8 * I + g, j = 8 * J + 2 * t;
        if (i < ncd) {
          if (j < ncd) { Go[i * ncd + j] = gacc[e][0]; if (I != J) Go[j * ncd + i] = gacc[e][0]; }
          if (j + 1 < ncd) { Go[i * ncd + j + 1] = gacc[e][1]; if (I != J) Go[(j + 1) * ncd + i] = gacc[e][1]; }
        }
      }
    }
  }
}

}  // namespace slod
