// k_patch_flux : W = S_b X_i - P_b, the boundary flux of every candidate before the M^{-1} scaling
// (B_full - PT_boundary of source/LOD.cc:609-617; S_boundary = A[boundary, internal] :520-528 is applied matrix free).
//
// One CTA per patch, 64 boundary rows per pass: the compact stencil rows towards interior dofs are assembled by the
// warps, the gather  W[b, :] = sum_e A[b, e] X[e, :]  reads X through L1 and writes the rows straight to W (a warp
// covers 512 contiguous bytes), then the projection weights of the at most 2^dim coarse cells containing the boundary
// dof are subtracted by a read-modify-write of the rows just written.  No staging tile: 18 KB of shared memory, four
// CTAs per SM and a large L1 that turns the 9-fold reuse of every X row into cache hits; two barriers per pass.
// Rows are in ascending boundary-dof order, zero padded to a multiple of 32 rows, for k_patch_dense_mma to stream.
// Included by kernels.cu.
#pragma once

namespace slod {

constexpr int kFTB = 64;    // boundary rows per pass
constexpr int kFNB = 20;    // stencil slots per boundary row: only inward offsets couple, 3^(dim-1) * spacedim <= 18

__global__ void __launch_bounds__(256, 4)
k_patch_flux(const int *__restrict__ patch_ids, int n_work, const double *__restrict__ d_coef,
             const double *__restrict__ Xbuf, double *__restrict__ Wbuf, FluxLayout lay, int *work_counter) {
  extern __shared__ double smem[];
  const int NC = lay.ldx;
  double *sCoef = smem;
  double *sArow = sCoef + lay.coef_doubles;     // [kFTB][kFNB]
  int *sAnbr = (int *)(sArow + kFTB * kFNB);    // [kFTB][kFNB]
  int *sAcnt = sAnbr + kFTB * kFNB;             // [kFTB]
  int *sBlist = sAcnt + kFTB;                   // [nb_max]
  __shared__ int sNb;
  const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5, NWARP = NT >> 5;

  __shared__ int sNextWork;
  SLOD_WORK_LOOP(w, n_work, work_counter, sNextWork) {
    fetch_work_item(w, work_counter, &sNextWork);
    const Geom geo = make_geom(cP, patch_ids[w]);
    if (!geo.slod) continue;
    const int ncd = geo.Ncd, s = cP.s;
    const double *X = Xbuf + (size_t)w * lay.x_stride;
    double *W = Wbuf + (size_t)w * lay.w_stride;
    __syncthreads();
    if (tid >= 32) load_coef(geo, d_coef, sCoef, tid - 32, NT - 32);   // beside warp 0's serial list building
    if (tid < 32) {  // boundary dofs (patch boundary, id 99), ascending
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
    const int nbd = sNb;
    const int nbd_pad = (nbd + 31) & ~31;
    const int lgn = __ffs(cP.n) - 1;
    for (int t0 = 0; t0 < nbd; t0 += kFTB) {
      const int nt = min(kFTB, nbd - t0);
      // Compact stencil rows.  A boundary dof only couples to interior dofs through offsets that point inward on
      // every axis on which its node sits on a patch side, so at most 3^(dim-1) * s entries exist: several rows share
      // a warp (lane = row-in-group x entry), ballot compaction inside each group keeps the ascending slot order.
      {
        const int epr = ((cP.dim == 3) ? 9 : 3) * s;   // entry candidates per row
        const int rpw = 32 / epr;                       // rows per warp pass
        const int sub = lane / epr, j = lane - sub * epr;
        for (int rb0 = warp * rpw; rb0 < kFTB; rb0 += NWARP * rpw) {
          const int rb = rb0 + sub;
          bool ok = (sub < rpw) && (rb < nt);
          int nbr = 0;
          double val = 0.0;
          if (ok) {
            const int dof = sBlist[t0 + rb];
            int a[3];
            node_coords(geo, dof / s, a);
            const int cb = j % s;
            int q = j / s;
            int dl[3] = {0, 0, 0};
            _Pragma("unroll") for (int x = 0; x < 3; ++x) if (x < cP.dim) {
              if (a[x] == 0) dl[x] = 1;
              else if (a[x] == geo.p[x] - 1) dl[x] = -1;
              else { dl[x] = q % 3 - 1; q /= 3; }
            }
            if (q != 0) ok = false;   // fewer free axes than the enumeration allows for (edge / corner nodes)
            int b[3] = {a[0] + dl[0], a[1] + dl[1], a[2] + dl[2]};
            _Pragma("unroll") for (int x = 0; x < 3; ++x) if (x < cP.dim) ok = ok && (b[x] >= 1 && b[x] <= geo.p[x] - 2);
            if (ok) {
              nbr = interior_index(geo, b) * s + cb;
              val = stiff_entry(cP, geo, sCoef, a, dl, dof % s, cb);
            }
          }
          const unsigned mask = __ballot_sync(0xffffffffu, ok);
          const unsigned gmask = (epr >= 32) ? 0xffffffffu : (((1u << epr) - 1u) << (sub * epr));
          const int count = __popc(mask & gmask);
          if (ok) {
            const int pos = __popc(mask & gmask & ((1u << lane) - 1u));
            sAnbr[rb * kFNB + pos] = nbr;
            sArow[rb * kFNB + pos] = val;
          }
          if (sub < rpw && rb < kFTB) {
            for (int pos = count + j; pos < kFNB; pos += epr) {  // padding: value 0 times X row 0
              sAnbr[rb * kFNB + pos] = 0;
              sArow[rb * kFNB + pos] = 0.0;
            }
            if (j == 0) sAcnt[rb] = count;
          }
        }
      }
      __syncthreads();
      // gather: thread = (column pair, row group); four rows at a time, lists are zero padded (no predicates)
      {
        const int rows_per_pass = NT / (NC / 2);  // 4 for 128 columns
        const int c2 = tid % (NC / 2), rbase = tid / (NC / 2);
        for (int r0 = 4 * rbase; r0 < kFTB; r0 += 4 * rows_per_pass) {   // four CONSECUTIVE rows per thread
          double2 acc[4];
          int cmax = 0;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            acc[u] = make_double2(0.0, 0.0);
            const int rb = r0 + u;
            if (rb < kFTB) cmax = max(cmax, sAcnt[rb]);
          }
          if (2 * c2 >= ncd) cmax = 0;
#pragma unroll 3
          for (int e = 0; e < cmax; ++e) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int rb = min(r0 + u, kFTB - 1);
              const double av = sArow[rb * kFNB + e];
              const double2 xv = *reinterpret_cast<const double2 *>(X + (size_t)sAnbr[rb * kFNB + e] * lay.ldx + 2 * c2);
              acc[u].x += av * xv.x;
              acc[u].y += av * xv.y;
            }
          }
          // straight to W: a warp covers 512 contiguous bytes of a row; rows past the last dof of the pass and the
          // padding columns carry zeros (empty stencil lists / cmax = 0)
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int rb = r0 + u;
            if (rb < kFTB && t0 + rb < nbd_pad)   // W holds round_up(nbd, 32) rows: dense_mma streams tiles of 32
              *reinterpret_cast<double2 *>(W + (size_t)(t0 + rb) * lay.ldx + 2 * c2) = acc[u];
          }
        }
      }
      __syncthreads();   // the rows are visible to the whole CTA; the stencil lists may be rebuilt
      // ... - P_b : every boundary dof lies in at most 2^dim coarse cells, each in its own column: a read-modify-write
      // of the row just written (L2), not waited for -- the next pass touches other rows
      for (int idx = tid; idx < nt * 8; idx += NT) {
        const int rb = idx >> 3, corner = idx & 7;
        if (corner >= (1 << cP.dim)) continue;
        const int dof = sBlist[t0 + rb];
        int a[3];
        node_coords(geo, dof / s, a);
        int kc[3] = {0, 0, 0};
        double wgt = cP.pw;
        bool ok = true;
        _Pragma("unroll") for (int x = 0; x < 3; ++x) if (x < cP.dim) {
          const int q = a[x] >> lgn, rem = a[x] - (q << lgn);
          if ((corner >> x) & 1) {
            if (rem != 0 || q < 1) ok = false;
            kc[x] = q - 1;
          } else {
            if (q > geo.m[x] - 1) ok = false;
            kc[x] = q;
            if (rem != 0) wgt *= 2.0;
          }
        }
        if (ok) W[(size_t)(t0 + rb) * lay.ldx + (lay.zmajor ? zcell_to_col(geo, kc) : cell_to_col(cP, geo, kc)) * s + dof % s] -= wgt;
      }
    }
  }
}

}  // namespace slod
