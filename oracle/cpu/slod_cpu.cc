// slod_cpu.cc -- multi-threaded C++ CPU restatement of the SLOD offline phase.  TEST / BASELINE INFRASTRUCTURE,
// NOT PRODUCT CODE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
// the library built from this file (oracle/_build/libslod_cpu.so).  The product (dealii-slod_b200) never does.
//
// It is the second, independent CPU statement of the reference algorithm (the first is oracle/slod_oracle.py) and the
// timed "reference CPU path": the reference itself (camillabelponer/dealii-slod) needs deal.II 9.6 + Trilinos + LAPACK,
// none of which exist in this image, so it cannot be built here.  The library exports the subset of include/slod.h the
// offline phase needs (same names, same argument meaning, same error behaviour), so the ctypes binding and the C++ host
// mirror run on it unchanged.  Parity status: pinned to the numpy oracle at 1e-10 (tests/test_cpu_port.py), which in
// turn is pinned to the reference's golden files; the SLOD branch (source/LOD.cc:596-757) is executed by no reference
// test, so for it this file and the oracle are two independent readings of the same lines.
//
// Reference lines followed (paths relative to /root/reference):
//   create_patches                        source/LOD.cc:122-244     -> Patch::Patch
//   create_mesh_for_patch (face ids 0/99) source/LOD.cc:770-858     -> Patch::node_class
//   fill_dofs_indices_vector              include/LODtools.h:334-375 -> Patch::classify
//   problem_parameter::value              include/Diffusion.h:40-53 -> Ctx::coef_at
//   assemble_stiffness                    include/Diffusion.h:111-207, include/Elasticity.h:163-299 -> assemble()
//   projection_P1_P0 + scatter            include/LODtools.h:7-73, source/LOD.cc:329-342, 470-496   -> projection()
//   boundary prep                         source/LOD.cc:498-544
//   Gauss_elimination                     include/LODtools.h:511-595 -> band_cholesky / band_solve (A_ii is SPD)
//   Schur complement + inverse            source/LOD.cc:548-553
//   LOD branch                            source/LOD.cc:563-595
//   SLOD branch                           source/LOD.cc:596-757      -> select()
//   premultiply                           source/LOD.cc:758-765
//   assemble_global_matrix                source/LOD.cc:860-973      -> coarse_rows()
//
// Deliberately different from both other implementations where the mathematics allows it: the patch matrix is assembled
// sub-cell by sub-cell and Gauss point by Gauss point into stencil storage (the CUDA path tabulates stencil entries on the
// fly, the oracle builds a scipy COO matrix), the solve is an unblocked row-oriented banded Cholesky (CUDA: blocked,
// look-ahead; oracle: SuperLU), M^-1 is Gauss-Jordan with partial pivoting (CUDA: blocked without pivoting; oracle:
// LAPACK LU), and the SVD of the Gram matrix is Householder tridiagonalisation + implicit QL with accumulated
// eigenvectors (CUDA: Cholesky fast path / rotation log; oracle: LAPACK dgesdd).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/slod.h"

namespace {

std::string g_create_error;

struct PatchOut {
  bool done = false;
  int Nf = 0;
  std::vector<double> phi, aphi;   // [s][Nf]
  double diag[2][8] = {{0}};
  int status = 0;
};

struct Patch {
  int dim, s, n, ell, N;
  int c[3], lo[3], hi[3], m[3], p[3], q[3];
  bool domlo[3], domhi[3];
  int nnodes, Nf, Nc, Ncd, Ni, nsubp;
  bool slod;
  std::vector<int> cells_rel;   // [Nc][3], centre first, then x-outer sweep (source/LOD.cc:151-178)
  Patch(int dim_, int s_, int n_, int ell_, int ref, int stabilize, uint32_t pid) : dim(dim_), s(s_), n(n_), ell(ell_) {
    N = 1 << ref;
    for (int a = 0; a < 3; ++a) c[a] = 0;
    for (int b = 0; b < ref; ++b)
      for (int a = 0; a < dim; ++a) c[a] |= (int)((pid >> (dim * b + a)) & 1u) << b;   // Morton, x = low bit
    nnodes = 1; Nc = 1; Ni = 1; nsubp = 1;
    for (int a = 0; a < 3; ++a) {
      if (a < dim) {
        lo[a] = std::max(c[a] - ell, 0);
        hi[a] = std::min(c[a] + ell, N - 1);
        m[a] = hi[a] - lo[a] + 1;
        p[a] = m[a] * n + 1;
        q[a] = p[a] - 2;
        domlo[a] = lo[a] == 0;
        domhi[a] = hi[a] == N - 1;
      } else {
        lo[a] = hi[a] = 0; m[a] = 1; p[a] = 1; q[a] = 1; domlo[a] = domhi[a] = false;
      }
      nnodes *= p[a]; Nc *= m[a]; Ni *= q[a];
      nsubp *= (a < dim) ? m[a] * n : 1;
    }
    Nf = s * nnodes; Ncd = s * Nc; Ni *= s;
    long long total = 1;
    for (int a = 0; a < dim; ++a) total *= N;
    slod = stabilize && ell > 0 && (long long)Nc != total;   // source/LOD.cc:563-564
    cells_rel.reserve(3 * Nc);
    const int cr[3] = {c[0] - lo[0], c[1] - lo[1], c[2] - lo[2]};
    cells_rel.insert(cells_rel.end(), cr, cr + 3);
    for (int a = 0; a < m[0]; ++a)
      for (int b = 0; b < m[1]; ++b)
        for (int e = 0; e < m[2]; ++e) {
          if (a == cr[0] && b == cr[1] && e == cr[2]) continue;
          const int k[3] = {a, b, e};
          cells_rel.insert(cells_rel.end(), k, k + 3);
        }
  }
  int node(int x, int y, int z) const { return (z * p[1] + y) * p[0] + x; }
  void coords(int nd, int a[3]) const { a[0] = nd % p[0]; nd /= p[0]; a[1] = nd % p[1]; a[2] = nd / p[1]; }
  // bit 0: patch boundary (id 99), bit 1: domain boundary (id 0); both may be set (include/LODtools.h:367-369)
  int node_class(const int a[3]) const {
    int cls = 0;
    for (int x = 0; x < dim; ++x) {
      if (a[x] == 0) cls |= domlo[x] ? 2 : 1;
      if (a[x] == p[x] - 1) cls |= domhi[x] ? 2 : 1;
    }
    return cls;
  }
  uint32_t cell_id(int k, int ref) const {
    uint32_t code = 0;
    for (int b = 0; b < ref; ++b)
      for (int a = 0; a < dim; ++a) code |= (uint32_t)(((lo[a] + cells_rel[3 * k + a]) >> b) & 1) << (dim * b + a);
    return code;
  }
};

}  // namespace

struct slod_ctx {
  slod_params par{};
  int dim = 2, s = 1, ref = 0, n = 1, ell = 0, N = 1;
  double H = 1, h = 1;
  int n_fields = 1;
  int64_t n_patches = 0;
  int NfMax = 0, w = 0, ell_width = 0;
  std::vector<double> table[2];
  int table_r[2] = {0, 0};
  bool table_set[2] = {false, false};
  int nthreads = 1;
  // sub-cell matrices per Gauss point (q = qx + 2 qy + 4 qz), local dof = s * (lx + 2 ly + 4 lz) + comp
  std::vector<double> Kq, Klamq;
  double gp[2];
  std::vector<PatchOut> out;
  bool basis_all = false, coarse_done = false;
  std::vector<int64_t> csr_rowptr, csr_col;
  std::vector<double> csr_val;
  std::vector<double> presaved;   // quirk B: stencil matrix of the first full-size patch (source/LOD.cc:354-362)
  bool has_presaved = false;
  double ms[8] = {0};
  mutable std::string err;
};

namespace {

int fail(const slod_ctx *c, int code, const std::string &m) {
  if (c) c->err = m;
  return code;
}

void local_matrices(slod_ctx *C) {
  const int dim = C->dim, s = C->s, nn = 1 << dim, nl = nn * s;
  const double h = C->h, g = 0.5 / std::sqrt(3.0);
  C->gp[0] = 0.5 - g; C->gp[1] = 0.5 + g;
  const double jxw = std::pow(h / 2.0, dim);
  C->Kq.assign((size_t)nn * nl * nl, 0.0);
  C->Klamq.assign((size_t)nn * nl * nl, 0.0);
  for (int q = 0; q < nn; ++q) {
    const double x[3] = {C->gp[q & 1], C->gp[(q >> 1) & 1], C->gp[(q >> 2) & 1]};
    double G[8][3];
    for (int i = 0; i < nn; ++i) {
      const int nd[3] = {i & 1, (i >> 1) & 1, (i >> 2) & 1};
      for (int a = 0; a < dim; ++a) {
        double v = 1.0;
        for (int b = 0; b < dim; ++b) v *= (b == a) ? (nd[b] ? 1.0 : -1.0) / h : (nd[b] ? x[b] : 1.0 - x[b]);
        G[i][a] = v;
      }
    }
    double *K = C->Kq.data() + (size_t)q * nl * nl, *L = C->Klamq.data() + (size_t)q * nl * nl;
    if (C->par.problem == SLOD_PROBLEM_DIFFUSION) {
      for (int i = 0; i < nn; ++i)
        for (int j = 0; j < nn; ++j) {
          double d = 0;
          for (int a = 0; a < dim; ++a) d += G[i][a] * G[j][a];
          K[i * nl + j] = d * jxw;     // include/Diffusion.h:181-186
        }
    } else {
      for (int i = 0; i < nn; ++i)
        for (int ci = 0; ci < s; ++ci)
          for (int j = 0; j < nn; ++j)
            for (int cj = 0; cj < s; ++cj) {
              double ee = 0;   // eps(phi_i) : eps(phi_j), phi_i = N_i e_ci   (include/Elasticity.h:236-250)
              for (int a = 0; a < dim; ++a)
                for (int b = 0; b < dim; ++b) {
                  const double ei = 0.5 * ((a == ci ? G[i][b] : 0.0) + (b == ci ? G[i][a] : 0.0));
                  const double ej = 0.5 * ((a == cj ? G[j][b] : 0.0) + (b == cj ? G[j][a] : 0.0));
                  ee += ei * ej;
                }
              K[(i * s + ci) * nl + j * s + cj] = 2.0 * ee * jxw;
              L[(i * s + ci) * nl + j * s + cj] = G[i][ci] * G[j][cj] * jxw;
            }
    }
  }
}

// problem_parameter::value (include/Diffusion.h:40-53): table[floor(x/eta) + 2^r floor(y/eta) (+ 4^r floor(z/eta))]
inline double coef_at(const slod_ctx *C, int f, const double x[3]) {
  const int nl = 1 << C->table_r[f];
  const double eta = 1.0 / nl;
  size_t idx = 0, mul = 1;
  for (int a = 0; a < C->dim; ++a) {
    long long ia = (long long)std::floor(x[a] / eta);
    if (ia < 0) ia = 0;
    if (ia > nl - 1) ia = nl - 1;
    idx += mul * (size_t)ia;
    mul *= nl;
  }
  return C->table[f][idx];
}

// Stencil storage of the unconstrained patch matrix: A[(node * nst + e) * s*s + ca * s + cb] couples dof (node, ca)
// with dof (node + offset e, cb), e = (dx+1) + 3 (dy+1) + 9 (dz+1).
struct Stencil {
  int nst, s;
  std::vector<double> v;
};

void assemble(const slod_ctx *C, const Patch &P, const int clo[3], Stencil &A) {
  const int dim = P.dim, s = P.s, nn = 1 << dim, nl = nn * s, n = P.n;
  A.nst = (dim == 3) ? 27 : 9;
  A.s = s;
  A.v.assign((size_t)P.nnodes * A.nst * s * s, 0.0);
  const int ms[3] = {P.m[0] * n, P.m[1] * n, dim == 3 ? P.m[2] * n : 1};
  std::vector<double> loc((size_t)nl * nl);
  for (int oz = 0; oz < ms[2]; ++oz)
    for (int oy = 0; oy < ms[1]; ++oy)
      for (int ox = 0; ox < ms[0]; ++ox) {
        const int o[3] = {ox, oy, oz};
        std::fill(loc.begin(), loc.end(), 0.0);
        for (int q = 0; q < nn; ++q) {   // cells x sub-cells x Gauss points (include/Diffusion.h:143-193)
          double x[3] = {0, 0, 0};
          for (int a = 0; a < dim; ++a) x[a] = ((double)(clo[a] * n + o[a]) + C->gp[(q >> a) & 1]) * C->h;
          const double *K = C->Kq.data() + (size_t)q * nl * nl, *L = C->Klamq.data() + (size_t)q * nl * nl;
          if (C->par.problem == SLOD_PROBLEM_DIFFUSION) {
            const double a0 = coef_at(C, 0, x);
            for (int i = 0; i < nl * nl; ++i) loc[i] += a0 * K[i];
          } else {
            const double lam = coef_at(C, 0, x), mu = coef_at(C, 1, x);
            for (int i = 0; i < nl * nl; ++i) loc[i] += mu * K[i] + lam * L[i];
          }
        }
        for (int i = 0; i < nn; ++i) {
          const int ai[3] = {ox + (i & 1), oy + ((i >> 1) & 1), oz + ((i >> 2) & 1)};
          const int ni = P.node(ai[0], ai[1], ai[2]);
          for (int j = 0; j < nn; ++j) {
            const int e = ((j & 1) - (i & 1) + 1) + 3 * ((((j >> 1) & 1) - ((i >> 1) & 1)) + 1) +
                          ((dim == 3) ? 9 * ((((j >> 2) & 1) - ((i >> 2) & 1)) + 1) : 0);
            double *dst = A.v.data() + ((size_t)ni * A.nst + e) * s * s;
            for (int ca = 0; ca < s; ++ca)
              for (int cb = 0; cb < s; ++cb) dst[ca * s + cb] += loc[(i * s + ca) * nl + j * s + cb];
          }
        }
      }
}

// P^T (Nf x Ncd, row-major): cell-local weights 1 / 2 / 4 (/ 8) for vertex / line / face / interior nodes times
// h^d / 2^d, summed over the patch cells, column s * k + comp for the k-th cell of the list
// (include/LODtools.h:7-73, source/LOD.cc:329-342, 470-496).
void projection(const slod_ctx *C, const Patch &P, std::vector<double> &PT) {
  const int dim = P.dim, s = P.s, n = P.n;
  PT.assign((size_t)P.Nf * P.Ncd, 0.0);
  const double base = std::pow(C->h, dim) / (double)(1 << dim);
  const int nz = (dim == 3) ? n : 0;
  for (int k = 0; k < P.Nc; ++k) {
    const int *cr = &P.cells_rel[3 * k];
    for (int tz = 0; tz <= nz; ++tz)
      for (int ty = 0; ty <= n; ++ty)
        for (int tx = 0; tx <= n; ++tx) {
          double wgt = base;
          if (tx != 0 && tx != n) wgt *= 2.0;
          if (ty != 0 && ty != n) wgt *= 2.0;
          if (dim == 3 && tz != 0 && tz != n) wgt *= 2.0;
          const int nd = P.node(cr[0] * n + tx, cr[1] * n + ty, (dim == 3) ? cr[2] * n + tz : 0);
          for (int comp = 0; comp < s; ++comp) PT[(size_t)(nd * s + comp) * P.Ncd + s * k + comp] += wgt;
        }
  }
}

// ---- dense helpers --------------------------------------------------------------------------------------------
// in-place inverse by Gauss-Jordan elimination with partial pivoting; returns false if singular
bool gauss_jordan(std::vector<double> &M, int n) {
  std::vector<int> piv(n);
  for (int k = 0; k < n; ++k) {
    int r = k;
    double best = std::fabs(M[(size_t)k * n + k]);
    for (int i = k + 1; i < n; ++i)
      if (std::fabs(M[(size_t)i * n + k]) > best) { best = std::fabs(M[(size_t)i * n + k]); r = i; }
    if (!(best > 0.0)) return false;
    piv[k] = r;
    if (r != k)
      for (int j = 0; j < n; ++j) std::swap(M[(size_t)k * n + j], M[(size_t)r * n + j]);
    const double d = 1.0 / M[(size_t)k * n + k];
    M[(size_t)k * n + k] = 1.0;
    for (int j = 0; j < n; ++j) M[(size_t)k * n + j] *= d;
    for (int i = 0; i < n; ++i) {
      if (i == k) continue;
      const double f = M[(size_t)i * n + k];
      if (f == 0.0) continue;
      M[(size_t)i * n + k] = 0.0;
      double *ri = &M[(size_t)i * n];
      const double *rk = &M[(size_t)k * n];
      for (int j = 0; j < n; ++j) ri[j] -= f * rk[j];
    }
  }
  for (int k = n - 1; k >= 0; --k)
    if (piv[k] != k)
      for (int i = 0; i < n; ++i) std::swap(M[(size_t)i * n + k], M[(size_t)i * n + piv[k]]);
  return true;
}

// Eigen-decomposition of a symmetric matrix: Householder reduction to tridiagonal form with accumulation of the
// transformation, then implicit-shift QL.  a (n x n, row-major) is overwritten by the eigenvectors (columns), d
// receives the eigenvalues.  Returns false if QL does not converge.
bool sym_eig(std::vector<double> &a, int n, std::vector<double> &d) {
  std::vector<double> e(n, 0.0);
  d.assign(n, 0.0);
  auto A = [&](int i, int j) -> double & { return a[(size_t)i * n + j]; };
  for (int i = n - 1; i >= 1; --i) {
    const int l = i - 1;
    double hh = 0.0, scale = 0.0;
    if (l > 0) {
      for (int k = 0; k <= l; ++k) scale += std::fabs(A(i, k));
      if (scale == 0.0) {
        e[i] = A(i, l);
      } else {
        for (int k = 0; k <= l; ++k) {
          A(i, k) /= scale;
          hh += A(i, k) * A(i, k);
        }
        double f = A(i, l);
        double g = (f >= 0.0) ? -std::sqrt(hh) : std::sqrt(hh);
        e[i] = scale * g;
        hh -= f * g;
        A(i, l) = f - g;
        f = 0.0;
        for (int j = 0; j <= l; ++j) {
          A(j, i) = A(i, j) / hh;
          g = 0.0;
          for (int k = 0; k <= j; ++k) g += A(j, k) * A(i, k);
          for (int k = j + 1; k <= l; ++k) g += A(k, j) * A(i, k);
          e[j] = g / hh;
          f += e[j] * A(i, j);
        }
        const double hk = f / (hh + hh);
        for (int j = 0; j <= l; ++j) {
          f = A(i, j);
          e[j] = g = e[j] - hk * f;
          for (int k = 0; k <= j; ++k) A(j, k) -= (f * e[k] + g * A(i, k));
        }
      }
    } else {
      e[i] = A(i, l);
    }
    d[i] = hh;
  }
  d[0] = 0.0;
  e[0] = 0.0;
  for (int i = 0; i < n; ++i) {
    const int l = i - 1;
    if (d[i] != 0.0) {
      for (int j = 0; j <= l; ++j) {
        double g = 0.0;
        for (int k = 0; k <= l; ++k) g += A(i, k) * A(k, j);
        for (int k = 0; k <= l; ++k) A(k, j) -= g * A(k, i);
      }
    }
    d[i] = A(i, i);
    A(i, i) = 1.0;
    for (int j = 0; j <= l; ++j) A(j, i) = A(i, j) = 0.0;
  }
  for (int i = 1; i < n; ++i) e[i - 1] = e[i];
  e[n - 1] = 0.0;
  for (int l = 0; l < n; ++l) {
    int iter = 0, mm;
    do {
      for (mm = l; mm < n - 1; ++mm) {
        const double dd = std::fabs(d[mm]) + std::fabs(d[mm + 1]);
        if (std::fabs(e[mm]) <= 2.220446049250313e-16 * dd) break;
      }
      if (mm != l) {
        if (iter++ == 200) return false;
        double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
        double r = std::hypot(g, 1.0);
        g = d[mm] - d[l] + e[l] / (g + (g >= 0.0 ? std::fabs(r) : -std::fabs(r)));
        double sn = 1.0, cs = 1.0, pp = 0.0;
        int i;
        for (i = mm - 1; i >= l; --i) {
          double f = sn * e[i];
          const double b = cs * e[i];
          e[i + 1] = (r = std::hypot(f, g));
          if (r == 0.0) {
            d[i + 1] -= pp;
            e[mm] = 0.0;
            break;
          }
          sn = f / r;
          cs = g / r;
          g = d[i + 1] - pp;
          r = (d[i] - g) * sn + 2.0 * cs * b;
          d[i + 1] = g + (pp = sn * r);
          g = cs * r - b;
          for (int k = 0; k < n; ++k) {
            f = A(k, i + 1);
            A(k, i + 1) = sn * A(k, i) + cs * f;
            A(k, i) = cs * A(k, i) - sn * f;
          }
        }
        if (r == 0.0 && i >= l) continue;
        d[l] -= pp;
        e[l] = g;
        e[mm] = 0.0;
      }
    } while (mm != l);
  }
  return true;
}

// ---- one patch (source/LOD.cc:345-767) --------------------------------------------------------------------------
struct Stages {   // optional intermediates for staged parity checks
  std::vector<double> *X = nullptr, *Minv = nullptr, *G = nullptr;
};

void compute_patch(slod_ctx *C, uint32_t pid, PatchOut &out, const Stages *stg = nullptr) {
  const Patch P(C->dim, C->s, C->n, C->ell, C->ref, C->par.stabilize, pid);
  const int dim = P.dim, s = P.s, Nf = P.Nf, Ncd = P.Ncd;
  out.Nf = Nf;
  out.status = 0;
  // ---- node classes and index lists (include/LODtools.h:334-375) ----
  std::vector<int> cls(P.nnodes), internal, bnd;   // dof lists, ascending
  for (int nd = 0; nd < P.nnodes; ++nd) {
    int a[3];
    P.coords(nd, a);
    cls[nd] = P.node_class(a);
    for (int c = 0; c < s; ++c) {
      if (cls[nd] == 0) internal.push_back(nd * s + c);
      if (cls[nd] & 1) bnd.push_back(nd * s + c);
    }
  }
  const int Ni = (int)internal.size(), Nb = (int)bnd.size();
  std::vector<int> int_of(Nf, -1);
  for (int i = 0; i < Ni; ++i) int_of[internal[i]] = i;

  // ---- stiffness (quirk B: full-size patches re-use the first one's matrix, source/LOD.cc:354-362, 433-451) ----
  Stencil A;
  bool full = true;
  for (int a = 0; a < dim; ++a) full = full && (P.m[a] == 2 * P.ell + 1);
  if (C->par.quirk_presaved && full && C->has_presaved) {
    A.nst = (dim == 3) ? 27 : 9;
    A.s = s;
    A.v = C->presaved;
  } else {
    assemble(C, P, P.lo, A);
  }
  const int nst = A.nst;
  auto nbr_node = [&](int nd, int e, int &ok) {
    int a[3];
    P.coords(nd, a);
    const int b[3] = {a[0] + e % 3 - 1, a[1] + (e / 3) % 3 - 1, a[2] + ((dim == 3) ? e / 9 - 1 : 0)};
    ok = 1;
    for (int x = 0; x < dim; ++x) ok = ok && b[x] >= 0 && b[x] < P.p[x];
    return ok ? P.node(b[0], b[1], b[2]) : -1;
  };

  // ---- P^T, PT_boundary (before zeroing), zeroed rows (source/LOD.cc:470-518) ----
  std::vector<double> PT;
  projection(C, P, PT);
  std::vector<double> PTb;
  if (P.slod) {
    PTb.resize((size_t)Nb * Ncd);
    for (int r = 0; r < Nb; ++r) std::memcpy(&PTb[(size_t)r * Ncd], &PT[(size_t)bnd[r] * Ncd], sizeof(double) * Ncd);
  }

  // ---- A_ii in lower band storage; the solve A_0 X = P^T_zeroed reduces to A_ii X_i = P_i, X = 0 elsewhere ----
  int bw = 0;
  for (int i = 0; i < Ni; ++i) {
    const int dof = internal[i], nd = dof / s, ca = dof % s;
    for (int e = 0; e < nst; ++e) {
      int ok;
      const int nb = nbr_node(nd, e, ok);
      if (!ok) continue;
      for (int cb = 0; cb < s; ++cb) {
        const int j = int_of[nb * s + cb];
        if (j >= 0 && j <= i) bw = std::max(bw, i - j);   // structural band (entries may vanish by symmetry)
      }
    }
  }
  const int ldb = bw + 1;
  std::vector<double> Lb((size_t)Ni * ldb, 0.0);   // Lb[i][k] = L(i, i - bw + k)
  for (int i = 0; i < Ni; ++i) {
    const int dof = internal[i], nd = dof / s, ca = dof % s;
    for (int e = 0; e < nst; ++e) {
      int ok;
      const int nb = nbr_node(nd, e, ok);
      if (!ok) continue;
      for (int cb = 0; cb < s; ++cb) {
        const int j = int_of[nb * s + cb];
        if (j >= 0 && j <= i) Lb[(size_t)i * ldb + (j - i + bw)] = A.v[((size_t)nd * nst + e) * s * s + ca * s + cb];
      }
    }
  }
  // row-oriented Cholesky
  for (int i = 0; i < Ni; ++i) {
    double *li = &Lb[(size_t)i * ldb];
    const int j0 = std::max(0, i - bw);
    for (int j = j0; j <= i; ++j) {
      const double *lj = &Lb[(size_t)j * ldb];
      const int k0 = std::max(j0, j - bw);
      double sum = li[j - i + bw];
      const double *pi = li + (k0 - i + bw), *pj = lj + (k0 - j + bw);
      const int cnt = j - k0;
      double acc = 0.0;
      for (int k = 0; k < cnt; ++k) acc += pi[k] * pj[k];
      sum -= acc;
      if (j < i) {
        li[j - i + bw] = sum / lj[bw];
      } else {
        if (!(sum > 0.0)) { out.status |= 1; sum = std::fabs(sum) + 1e-300; }
        li[bw] = std::sqrt(sum);
      }
    }
  }
  // forward / backward substitution with all Ncd right-hand sides (rows of X contiguous)
  std::vector<double> X((size_t)Ni * Ncd);
  for (int i = 0; i < Ni; ++i) {
    double *xi = &X[(size_t)i * Ncd];
    std::memcpy(xi, &PT[(size_t)internal[i] * Ncd], sizeof(double) * Ncd);
    const double *li = &Lb[(size_t)i * ldb];
    for (int k = std::max(0, i - bw); k < i; ++k) {
      const double f = li[k - i + bw];
      if (f == 0.0) continue;
      const double *xk = &X[(size_t)k * Ncd];
      for (int c = 0; c < Ncd; ++c) xi[c] -= f * xk[c];
    }
    const double inv = 1.0 / li[bw];
    for (int c = 0; c < Ncd; ++c) xi[c] *= inv;
  }
  for (int i = Ni - 1; i >= 0; --i) {
    double *xi = &X[(size_t)i * Ncd];
    for (int k = i + 1; k <= std::min(Ni - 1, i + bw); ++k) {
      const double f = Lb[(size_t)k * ldb + (i - k + bw)];
      if (f == 0.0) continue;
      const double *xk = &X[(size_t)k * Ncd];
      for (int c = 0; c < Ncd; ++c) xi[c] -= f * xk[c];
    }
    const double inv = 1.0 / Lb[(size_t)i * ldb + bw];
    for (int c = 0; c < Ncd; ++c) xi[c] *= inv;
  }

  // ---- M = PT^T (A_0^-1 PT) / H^d, M^-1 (source/LOD.cc:548-553) ----
  std::vector<double> M((size_t)Ncd * Ncd, 0.0);
  for (int i = 0; i < Ni; ++i) {
    const double *pr = &PT[(size_t)internal[i] * Ncd], *xi = &X[(size_t)i * Ncd];
    for (int a = 0; a < Ncd; ++a) {
      const double f = pr[a];
      if (f == 0.0) continue;
      double *mr = &M[(size_t)a * Ncd];
      for (int c = 0; c < Ncd; ++c) mr[c] += f * xi[c];
    }
  }
  const double Hd = std::pow(C->H, dim);
  for (auto &v : M) v /= Hd;
  std::vector<double> Minv = M;
  if (!gauss_jordan(Minv, Ncd)) out.status |= 2;
  if (stg && stg->X) *stg->X = X;
  if (stg && stg->Minv) *stg->Minv = Minv;

  std::vector<double> cvec((size_t)s * Ncd, 0.0);
  for (int d = 0; d < s; ++d)
    for (int k = 0; k < 8; ++k) out.diag[d][k] = 0.0;
  if (!P.slod) {
    for (int d = 0; d < s; ++d)
      for (int r = 0; r < Ncd; ++r) cvec[(size_t)d * Ncd + r] = Minv[(size_t)r * Ncd + d];   // source/LOD.cc:570-576
  } else {
    // ---- BD = S_b X_i M^-1 - PT_b M^-1 (source/LOD.cc:520-528, 609-618) ----
    std::vector<double> Bfull((size_t)Nb * Ncd, 0.0);
    for (int r = 0; r < Nb; ++r) {
      const int dof = bnd[r], nd = dof / s, ca = dof % s;
      double *br = &Bfull[(size_t)r * Ncd];
      for (int e = 0; e < nst; ++e) {
        int ok;
        const int nb = nbr_node(nd, e, ok);
        if (!ok) continue;
        for (int cb = 0; cb < s; ++cb) {
          const int j = int_of[nb * s + cb];
          if (j < 0) continue;
          const double f = A.v[((size_t)nd * nst + e) * s * s + ca * s + cb];   // S_boundary = A[b, internal], unconstrained
          if (f == 0.0) continue;
          const double *xj = &X[(size_t)j * Ncd];
          for (int c = 0; c < Ncd; ++c) br[c] += f * xj[c];
        }
      }
    }
    auto times_minv = [&](const std::vector<double> &T, std::vector<double> &R) {
      R.assign((size_t)Nb * Ncd, 0.0);
      for (int r = 0; r < Nb; ++r) {
        const double *tr = &T[(size_t)r * Ncd];
        double *rr = &R[(size_t)r * Ncd];
        for (int k = 0; k < Ncd; ++k) {
          const double f = tr[k];
          if (f == 0.0) continue;
          const double *mk = &Minv[(size_t)k * Ncd];
          for (int c = 0; c < Ncd; ++c) rr[c] += f * mk[c];
        }
      }
    };
    std::vector<double> BD, T2;
    times_minv(Bfull, BD);
    times_minv(PTb, T2);
    for (size_t i = 0; i < BD.size(); ++i) BD[i] -= T2[i];
    // Gram matrix of all columns once: G[o,o] and g = G[o,d] are sub-blocks of it with the same row-sum order
    // (source/LOD.cc:656-662 forms them per component)
    std::vector<double> Gf((size_t)Ncd * Ncd, 0.0);
    for (int r = 0; r < Nb; ++r) {
      const double *br = &BD[(size_t)r * Ncd];
      for (int a = 0; a < Ncd; ++a) {
        const double f = br[a];
        double *ga = &Gf[(size_t)a * Ncd];
        for (int c = 0; c <= a; ++c) ga[c] += f * br[c];
      }
    }
    for (int a = 0; a < Ncd; ++a)
      for (int c = a + 1; c < Ncd; ++c) Gf[(size_t)a * Ncd + c] = Gf[(size_t)c * Ncd + a];
    if (stg && stg->G) *stg->G = Gf;
    const int nc = Ncd - 1;
    for (int d = 0; d < s; ++d) {
      std::vector<int> other;
      for (int k = 0; k < Ncd; ++k)
        if (k != d) other.push_back(k);   // source/LOD.cc:637-640
      std::vector<double> V((size_t)nc * nc), g(nc), lam;
      for (int a = 0; a < nc; ++a) {
        g[a] = Gf[(size_t)other[a] * Ncd + d];
        for (int c = 0; c < nc; ++c) V[(size_t)a * nc + c] = Gf[(size_t)other[a] * Ncd + other[c]];
      }
      if (!sym_eig(V, nc, lam)) out.status |= 4;
      // SVD of a symmetric matrix: sigma = |lambda| descending, v = eigenvector, u = sign(lambda) v
      std::vector<int> ord(nc);
      for (int i = 0; i < nc; ++i) ord[i] = i;
      std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return std::fabs(lam[a]) > std::fabs(lam[b]); });
      const double sig0 = std::fabs(lam[ord[0]]);
      std::vector<double> coef(nc);   // (u_i^T g) / sigma_i, zero below the threshold 1e-15 sigma_0 (source/LOD.cc:667)
      double kept_min = sig0;
      for (int i = 0; i < nc; ++i) {
        const int col = ord[i];
        const double sg = std::fabs(lam[col]);
        double dot = 0.0;
        for (int a = 0; a < nc; ++a) dot += V[(size_t)a * nc + col] * g[a];
        if (sg > 1e-15 * sig0) {
          coef[i] = dot / lam[col];   // sign(lambda) (v^T g) / |lambda|
          kept_min = std::min(kept_min, sg);
        } else {
          coef[i] = 0.0;
        }
      }
      std::vector<double> di(nc, 0.0);
      for (int i = 0; i < nc; ++i)
        for (int a = 0; a < nc; ++a) di[a] -= V[(size_t)a * nc + ord[i]] * coef[i];   // d = -G^+ g  (source/LOD.cc:669-671)
      auto dinf = [&]() {
        double mx = 0.0;
        for (double v : di) mx = std::max(mx, std::fabs(v));
        return mx;
      };
      out.diag[d][0] = dinf();
      int steps = 0;
      for (int i = nc - 1; i >= 0; --i) {   // source/LOD.cc:703-725
        if (dinf() < 0.5) break;
        for (int a = 0; a < nc; ++a) di[a] += V[(size_t)a * nc + ord[i]] * coef[i];
        ++steps;
      }
      out.diag[d][1] = steps;
      out.diag[d][2] = sig0;
      out.diag[d][3] = kept_min;
      out.diag[d][5] = 10;
      double *cv = &cvec[(size_t)d * Ncd];
      for (int r = 0; r < Ncd; ++r) cv[r] = Minv[(size_t)r * Ncd + d];   // source/LOD.cc:727-743
      for (int idx = 0; idx < nc; ++idx) {
        const int k = other[idx];
        for (int r = 0; r < Ncd; ++r) cv[r] += di[idx] * Minv[(size_t)r * Ncd + k];
      }
    }
  }
  // ---- phi = X c, zero extension, normalisation (source/LOD.cc:745-754), A phi with the domain-boundary rows of
  //      semi_constrained replaced by identity rows (source/LOD.cc:537-541, 758-765) ----
  out.phi.assign((size_t)s * Nf, 0.0);
  out.aphi.assign((size_t)s * Nf, 0.0);
  for (int d = 0; d < s; ++d) {
    double *phi = &out.phi[(size_t)d * Nf];
    const double *cv = &cvec[(size_t)d * Ncd];
    double nrm = 0.0;
    for (int i = 0; i < Ni; ++i) {
      const double *xi = &X[(size_t)i * Ncd];
      double acc = 0.0;
      for (int c = 0; c < Ncd; ++c) acc += xi[c] * cv[c];
      phi[internal[i]] = acc;
      nrm += acc * acc;
    }
    nrm = std::sqrt(nrm);
    for (int i = 0; i < Nf; ++i) phi[i] /= nrm;
    double *aphi = &out.aphi[(size_t)d * Nf];
    for (int nd = 0; nd < P.nnodes; ++nd)
      for (int ca = 0; ca < s; ++ca) {
        if (cls[nd] & 2) { aphi[nd * s + ca] = phi[nd * s + ca]; continue; }
        double acc = 0.0;
        for (int e = 0; e < nst; ++e) {
          int ok;
          const int nb = nbr_node(nd, e, ok);
          if (!ok) continue;
          for (int cb = 0; cb < s; ++cb) acc += A.v[((size_t)nd * nst + e) * s * s + ca * s + cb] * phi[nb * s + cb];
        }
        aphi[nd * s + ca] = acc;
      }
    out.diag[d][7] = out.status;
  }
  out.done = true;
}

template <typename F>
void parallel_for(int nthreads, size_t n, F fn) {
  std::atomic<size_t> next{0};
  std::vector<std::thread> th;
  const int nt = (int)std::max<size_t>(1, std::min<size_t>(nthreads, n));
  for (int t = 0; t < nt; ++t)
    th.emplace_back([&]() {
      for (;;) {
        const size_t i = next.fetch_add(1);
        if (i >= n) break;
        fn(i);
      }
    });
  for (auto &x : th) x.join();
}

uint32_t morton(const int c[3], int dim, int ref) {
  uint32_t code = 0;
  for (int b = 0; b < ref; ++b)
    for (int a = 0; a < dim; ++a) code |= (uint32_t)((c[a] >> b) & 1) << (dim * b + a);
  return code;
}

// neighbours of patch pid whose node boxes intersect its own, ascending patch id: the structural pattern of
// C^T (A C) (source/LOD.cc:970-971)
void neighbours(const slod_ctx *C, uint32_t pid, std::vector<uint32_t> &nb) {
  nb.clear();
  const Patch P(C->dim, C->s, C->n, C->ell, C->ref, 0, pid);
  const int w = C->w;
  int lo[3], hi[3];
  for (int a = 0; a < 3; ++a) {
    lo[a] = (a < C->dim) ? std::max(0, P.c[a] - w) : 0;
    hi[a] = (a < C->dim) ? std::min(C->N - 1, P.c[a] + w) : 0;
  }
  for (int z = lo[2]; z <= hi[2]; ++z)
    for (int y = lo[1]; y <= hi[1]; ++y)
      for (int x = lo[0]; x <= hi[0]; ++x) {
        const int qc[3] = {x, y, z};
        bool ok = true;
        for (int a = 0; a < C->dim; ++a) {
          const int qlo = std::max(qc[a] - C->ell, 0), qhi = std::min(qc[a] + C->ell, C->N - 1);
          ok = ok && std::max(qlo, P.lo[a]) <= std::min(qhi, P.hi[a]) + 1;   // node boxes [lo n, (hi + 1) n] intersect
        }
        if (ok) nb.push_back(morton(qc, C->dim, C->ref));
      }
  std::sort(nb.begin(), nb.end());
}

// K[(p,d),(q,e)] = phi_{p,d} . (A phi)_{q,e} over the fine dofs both patches hold
double overlap_dot(const slod_ctx *C, const Patch &P, const double *phi, const Patch &Q, const double *aphi, int d, int e) {
  const int n = C->n, s = C->s;
  int b0[3], b1[3];
  for (int a = 0; a < 3; ++a) {
    if (a < C->dim) {
      b0[a] = std::max(P.lo[a], Q.lo[a]) * n;
      b1[a] = (std::min(P.hi[a], Q.hi[a]) + 1) * n;
      if (b1[a] < b0[a]) return 0.0;
    } else {
      b0[a] = b1[a] = 0;
    }
  }
  const double *pp = phi + (size_t)d * P.Nf, *qq = aphi + (size_t)e * Q.Nf;
  double acc = 0.0;
  for (int z = b0[2]; z <= b1[2]; ++z)
    for (int y = b0[1]; y <= b1[1]; ++y) {
      const int pz = (C->dim == 3) ? z - P.lo[2] * n : 0, qz = (C->dim == 3) ? z - Q.lo[2] * n : 0;
      const int pb = P.node(b0[0] - P.lo[0] * n, y - P.lo[1] * n, pz), qb = Q.node(b0[0] - Q.lo[0] * n, y - Q.lo[1] * n, qz);
      const int len = (b1[0] - b0[0] + 1) * s;
      const double *p1 = pp + (size_t)pb * s, *q1 = qq + (size_t)qb * s;
      for (int i = 0; i < len; ++i) acc += p1[i] * q1[i];
    }
  return acc;
}

int check_patch(const slod_ctx *ctx, int64_t patch) {
  if (!ctx) return SLOD_ERR_INVALID;
  if (patch < 0 || patch >= ctx->n_patches) return fail(ctx, SLOD_ERR_INVALID, "patch id out of range");
  return SLOD_OK;
}

double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int compute_list(slod_ctx *ctx, const int64_t *ids, int64_t n) {
  for (int f = 0; f < ctx->n_fields; ++f)
    if (!ctx->table_set[f]) return fail(ctx, SLOD_ERR_STATE, "coefficient field not set");
  const double t0 = now_ms();
  if (ctx->par.quirk_presaved && !ctx->has_presaved) {
    for (int64_t pid = 0; pid < ctx->n_patches; ++pid) {   // the first full-size patch in patch-id order donates its matrix
      const Patch P(ctx->dim, ctx->s, ctx->n, ctx->ell, ctx->ref, 0, (uint32_t)pid);
      bool full = true;
      for (int a = 0; a < ctx->dim; ++a) full = full && (P.m[a] == 2 * ctx->ell + 1);
      if (!full) continue;
      Stencil A;
      assemble(ctx, P, P.lo, A);
      ctx->presaved = A.v;
      ctx->has_presaved = true;
      break;
    }
  }
  parallel_for(ctx->nthreads, (size_t)n, [&](size_t i) {
    const int64_t pid = ids ? ids[i] : (int64_t)i;
    compute_patch(ctx, (uint32_t)pid, ctx->out[(size_t)pid]);
  });
  ctx->ms[0] = now_ms() - t0;
  for (int64_t i = 0; i < n; ++i) {
    const int64_t pid = ids ? ids[i] : i;
    if (ctx->out[(size_t)pid].status) {
      char buf[160];
      std::snprintf(buf, sizeof buf, "patch %lld: numerical status bits 0x%x (1 A_ii not SPD, 2 M singular, 4 QL not converged)",
                    (long long)pid, ctx->out[(size_t)pid].status);
      return fail(ctx, SLOD_ERR_NUMERIC, buf);
    }
  }
  return SLOD_OK;
}

}  // namespace

// =====================================================================================================================
extern "C" {

const char *slod_last_create_error(void) { return g_create_error.c_str(); }
const char *slod_last_error(const slod_ctx *ctx) { return ctx ? ctx->err.c_str() : "null handle"; }

int slod_create(const slod_params *par, slod_ctx **out) {
  if (!par || !out) { g_create_error = "null argument"; return SLOD_ERR_INVALID; }
  *out = nullptr;
  auto bad = [&](int code, const char *m) { g_create_error = m; return code; };
  if (par->dim != 2 && par->dim != 3) return bad(SLOD_ERR_INVALID, "dim must be 2 or 3");
  if (par->problem == SLOD_PROBLEM_DIFFUSION && par->spacedim != 1) return bad(SLOD_ERR_INVALID, "diffusion needs spacedim 1");
  if (par->problem == SLOD_PROBLEM_ELASTICITY && (par->spacedim != par->dim || par->dim != 2))
    return bad(SLOD_ERR_UNSUPPORTED, "elasticity is implemented for dim = spacedim = 2 (as in the reference)");
  if (par->problem != SLOD_PROBLEM_DIFFUSION && par->problem != SLOD_PROBLEM_ELASTICITY) return bad(SLOD_ERR_INVALID, "unknown problem");
  if (par->n_global_refinements < 0 || par->n_global_refinements * par->dim > 30)
    return bad(SLOD_ERR_INVALID, "n_global_refinements out of range");
  if (par->n_subdivisions < 1 || (par->n_subdivisions & (par->n_subdivisions - 1)))
    return bad(SLOD_ERR_INVALID, "n_subdivisions must be a power of two (include/Diffusion.h:76-80)");
  if (par->oversampling < 0) return bad(SLOD_ERR_INVALID, "oversampling < 0");
  auto *C = new slod_ctx();
  C->par = *par;
  C->dim = par->dim; C->s = par->spacedim; C->ref = par->n_global_refinements; C->n = par->n_subdivisions;
  C->ell = par->oversampling; C->N = 1 << C->ref;
  C->H = std::ldexp(1.0, -C->ref);
  C->h = C->H / C->n;
  C->n_fields = (par->problem == SLOD_PROBLEM_DIFFUSION) ? 1 : 2;
  C->n_patches = 1;
  for (int a = 0; a < C->dim; ++a) C->n_patches *= C->N;
  const int mfull = std::min(2 * C->ell + 1, C->N);
  C->NfMax = C->s;
  for (int a = 0; a < C->dim; ++a) C->NfMax *= C->n * mfull + 1;
  C->w = 2 * C->ell + 1;
  C->ell_width = C->s;
  for (int a = 0; a < C->dim; ++a) C->ell_width *= 2 * C->w + 1;
  const char *env = std::getenv("SLOD_CPU_THREADS");
  C->nthreads = env ? std::max(1, std::atoi(env)) : (int)std::max(1u, std::thread::hardware_concurrency());
  local_matrices(C);
  C->out.resize((size_t)C->n_patches);
  *out = C;
  return SLOD_OK;
}

void slod_destroy(slod_ctx *ctx) { delete ctx; }

int slod_set_coefficient(slod_ctx *ctx, int field, int eta_refinement, const double *cellwise, size_t n) {
  if (!ctx || !cellwise) return SLOD_ERR_INVALID;
  if (field < 0 || field >= ctx->n_fields) return fail(ctx, SLOD_ERR_INVALID, "coefficient field index out of range");
  if (eta_refinement < 0 || eta_refinement * ctx->dim > 30) return fail(ctx, SLOD_ERR_INVALID, "eta_refinement out of range");
  size_t want = 1;
  for (int a = 0; a < ctx->dim; ++a) want *= (size_t)1 << eta_refinement;
  if (want != n) return fail(ctx, SLOD_ERR_INVALID, "coefficient table size != (2^r)^dim");
  ctx->table[field].assign(cellwise, cellwise + n);
  ctx->table_r[field] = eta_refinement;
  ctx->table_set[field] = true;
  for (auto &o : ctx->out) o.done = false;
  ctx->basis_all = ctx->coarse_done = false;
  ctx->has_presaved = false;
  return SLOD_OK;
}

int slod_patch_count(const slod_ctx *ctx, int64_t *n) {
  if (!ctx || !n) return SLOD_ERR_INVALID;
  *n = ctx->n_patches;
  return SLOD_OK;
}

int slod_get_patch_info(const slod_ctx *ctx, int64_t patch, int32_t *n_cells, int32_t *n_fine, int32_t *n_internal,
                        int32_t *n_boundary, int32_t *n_domain_boundary, int32_t *n_coarse, int32_t lo[3], int32_t m[3]) {
  int rc = check_patch(ctx, patch);
  if (rc) return rc;
  const Patch P(ctx->dim, ctx->s, ctx->n, ctx->ell, ctx->ref, ctx->par.stabilize, (uint32_t)patch);
  int nb = 0, ndb = 0;
  for (int nd = 0; nd < P.nnodes; ++nd) {
    int a[3];
    P.coords(nd, a);
    const int c = P.node_class(a);
    nb += (c & 1) ? P.s : 0;
    ndb += (c & 2) ? P.s : 0;
  }
  if (n_cells) *n_cells = P.Nc;
  if (n_fine) *n_fine = P.Nf;
  if (n_internal) *n_internal = P.Ni;
  if (n_boundary) *n_boundary = nb;
  if (n_domain_boundary) *n_domain_boundary = ndb;
  if (n_coarse) *n_coarse = P.Ncd;
  for (int a = 0; a < 3; ++a) {
    if (lo) lo[a] = P.lo[a];
    if (m) m[a] = P.m[a];
  }
  return SLOD_OK;
}

int slod_get_patch_cells(const slod_ctx *ctx, int64_t patch, uint32_t *cells, int32_t *n) {
  int rc = check_patch(ctx, patch);
  if (rc) return rc;
  const Patch P(ctx->dim, ctx->s, ctx->n, ctx->ell, ctx->ref, 0, (uint32_t)patch);
  if (n) *n = P.Nc;
  if (cells)
    for (int k = 0; k < P.Nc; ++k) cells[k] = P.cell_id(k, ctx->ref);
  return SLOD_OK;
}

int slod_get_patch_dof_class(const slod_ctx *ctx, int64_t patch, int which, uint32_t *dofs, int32_t *n) {
  int rc = check_patch(ctx, patch);
  if (rc) return rc;
  if (which < 0 || which > 2) return fail(ctx, SLOD_ERR_INVALID, "which must be 0, 1 or 2");
  const Patch P(ctx->dim, ctx->s, ctx->n, ctx->ell, ctx->ref, 0, (uint32_t)patch);
  int cnt = 0;
  for (int nd = 0; nd < P.nnodes; ++nd) {
    int a[3];
    P.coords(nd, a);
    const int c = P.node_class(a);
    const bool in = (which == 0) ? (c == 0) : (which == 1 ? (c & 1) != 0 : (c & 2) != 0);
    if (!in) continue;
    for (int k = 0; k < P.s; ++k) {
      if (dofs) dofs[cnt] = (uint32_t)(nd * P.s + k);
      ++cnt;
    }
  }
  if (n) *n = cnt;
  return SLOD_OK;
}

int slod_basis_stride(const slod_ctx *ctx, int64_t *stride) {
  if (!ctx || !stride) return SLOD_ERR_INVALID;
  *stride = ctx->NfMax;
  return SLOD_OK;
}
int slod_ell_width(const slod_ctx *ctx, int64_t *width) {
  if (!ctx || !width) return SLOD_ERR_INVALID;
  *width = ctx->ell_width;
  return SLOD_OK;
}

int slod_compute_basis(slod_ctx *ctx) {
  if (!ctx) return SLOD_ERR_INVALID;
  ctx->coarse_done = false;
  int rc = compute_list(ctx, nullptr, ctx->n_patches);
  ctx->basis_all = (rc == SLOD_OK);
  return rc;
}

/* extension of the CPU port (no counterpart in slod.h): compute only the listed patches -- sampled parity checks at
 * BASELINE size and the bounded samples of bench.py's reference arm. */
int slod_cpu_compute_patches(slod_ctx *ctx, const int64_t *ids, int64_t n) {
  if (!ctx || (!ids && n > 0)) return SLOD_ERR_INVALID;
  for (int64_t i = 0; i < n; ++i)
    if (ids[i] < 0 || ids[i] >= ctx->n_patches) return fail(ctx, SLOD_ERR_INVALID, "patch id out of range");
  return compute_list(ctx, ids, n);
}

int slod_cpu_set_threads(slod_ctx *ctx, int n) {
  if (!ctx || n < 1) return SLOD_ERR_INVALID;
  ctx->nthreads = n;
  return SLOD_OK;
}
int slod_cpu_get_threads(const slod_ctx *ctx) { return ctx ? ctx->nthreads : 0; }

int slod_get_basis(const slod_ctx *ctx, int64_t patch, int comp, double *phi, double *aphi) {
  int rc = check_patch(ctx, patch);
  if (rc) return rc;
  const PatchOut &o = ctx->out[(size_t)patch];
  if (!o.done) return fail(ctx, SLOD_ERR_STATE, "patch has not been computed");
  if (comp < 0 || comp >= ctx->s) return fail(ctx, SLOD_ERR_INVALID, "component out of range");
  if (phi) std::memcpy(phi, &o.phi[(size_t)comp * o.Nf], sizeof(double) * o.Nf);
  if (aphi) std::memcpy(aphi, &o.aphi[(size_t)comp * o.Nf], sizeof(double) * o.Nf);
  return SLOD_OK;
}

int slod_get_all_basis(const slod_ctx *ctx, double *phi, double *aphi) {
  if (!ctx) return SLOD_ERR_INVALID;
  if (!ctx->basis_all) return fail(ctx, SLOD_ERR_STATE, "slod_compute_basis has not run");
  const size_t stride = (size_t)ctx->NfMax;
  for (int64_t p = 0; p < ctx->n_patches; ++p) {
    const PatchOut &o = ctx->out[(size_t)p];
    for (int d = 0; d < ctx->s; ++d) {
      const size_t off = ((size_t)p * ctx->s + d) * stride;
      if (phi) {
        std::memset(phi + off, 0, sizeof(double) * stride);
        std::memcpy(phi + off, &o.phi[(size_t)d * o.Nf], sizeof(double) * o.Nf);
      }
      if (aphi) {
        std::memset(aphi + off, 0, sizeof(double) * stride);
        std::memcpy(aphi + off, &o.aphi[(size_t)d * o.Nf], sizeof(double) * o.Nf);
      }
    }
  }
  return SLOD_OK;
}

/* one row of K = C^T (A C) (columns ascending, structural zeros kept); needs the patch of the row and every
 * neighbour computed.  Extension of the CPU port for sampled checks at BASELINE size. */
int slod_cpu_get_coarse_row(const slod_ctx *ctx, int64_t row, int64_t *col, double *val, int32_t *n) {
  if (!ctx || !n) return SLOD_ERR_INVALID;
  const int s = ctx->s;
  const int64_t pid = row / s;
  const int d = (int)(row % s);
  int rc = check_patch(ctx, pid);
  if (rc) return rc;
  std::vector<uint32_t> nb;
  neighbours(ctx, (uint32_t)pid, nb);
  *n = (int32_t)nb.size() * s;
  if (!col || !val) return SLOD_OK;
  const PatchOut &op = ctx->out[(size_t)pid];
  if (!op.done) return fail(ctx, SLOD_ERR_STATE, "row patch has not been computed");
  const Patch P(ctx->dim, s, ctx->n, ctx->ell, ctx->ref, 0, (uint32_t)pid);
  int o = 0;
  for (uint32_t q : nb) {
    const PatchOut &oq = ctx->out[q];
    if (!oq.done) return fail(ctx, SLOD_ERR_STATE, "a neighbour patch has not been computed");
    const Patch Q(ctx->dim, s, ctx->n, ctx->ell, ctx->ref, 0, q);
    for (int e = 0; e < s; ++e) {
      col[o] = (int64_t)q * s + e;
      val[o] = overlap_dot(ctx, P, op.phi.data(), Q, oq.aphi.data(), d, e);
      ++o;
    }
  }
  return SLOD_OK;
}

// K = C^T (A C).  subset == false: every patch must be computed (slod_assemble_coarse).  subset == true (extension for
// bounded samples): rows of the computed patches only, restricted to the columns of computed neighbours.
static int assemble_rows(slod_ctx *ctx, bool subset) {
  const double t0 = now_ms();
  const int s = ctx->s;
  const int64_t np = ctx->n_patches;
  std::vector<int64_t> cnt((size_t)np + 1, 0);
  auto usable = [&](uint32_t q) { return !subset || ctx->out[q].done; };
  parallel_for(ctx->nthreads, (size_t)np, [&](size_t p) {
    if (!usable((uint32_t)p)) return;
    std::vector<uint32_t> nb;
    neighbours(ctx, (uint32_t)p, nb);
    int64_t c = 0;
    for (uint32_t q : nb) c += usable(q) ? 1 : 0;
    cnt[p + 1] = c;
  });
  ctx->csr_rowptr.assign((size_t)np * s + 1, 0);
  std::vector<int64_t> start((size_t)np + 1, 0);
  for (int64_t p = 0; p < np; ++p) start[p + 1] = start[p] + cnt[p + 1] * s * s;
  ctx->csr_col.resize((size_t)start[np]);
  ctx->csr_val.resize((size_t)start[np]);
  parallel_for(ctx->nthreads, (size_t)np, [&](size_t p) {
    const int64_t per_row = cnt[p + 1] * s;
    for (int d = 0; d < s; ++d) ctx->csr_rowptr[p * s + d] = start[p] + d * per_row;
    if (!usable((uint32_t)p)) return;
    std::vector<uint32_t> nb;
    neighbours(ctx, (uint32_t)p, nb);
    const Patch P(ctx->dim, s, ctx->n, ctx->ell, ctx->ref, 0, (uint32_t)p);
    const PatchOut &op = ctx->out[p];
    int k = 0;
    for (uint32_t q : nb) {
      if (!usable(q)) continue;
      const Patch Q(ctx->dim, s, ctx->n, ctx->ell, ctx->ref, 0, q);
      const PatchOut &oq = ctx->out[q];
      for (int d = 0; d < s; ++d)
        for (int e = 0; e < s; ++e) {
          const int64_t o = start[p] + d * per_row + (int64_t)k * s + e;
          ctx->csr_col[(size_t)o] = (int64_t)q * s + e;
          ctx->csr_val[(size_t)o] = overlap_dot(ctx, P, op.phi.data(), Q, oq.aphi.data(), d, e);
        }
      ++k;
    }
  });
  ctx->csr_rowptr[(size_t)np * s] = start[np];
  ctx->coarse_done = true;
  ctx->ms[4] = now_ms() - t0;
  return SLOD_OK;
}

int slod_assemble_coarse(slod_ctx *ctx) {
  if (!ctx) return SLOD_ERR_INVALID;
  if (!ctx->basis_all) return fail(ctx, SLOD_ERR_STATE, "slod_compute_basis has not run");
  return assemble_rows(ctx, false);
}

/* extension: coarse-matrix rows of the patches computed so far, columns restricted to computed neighbours */
int slod_cpu_assemble_coarse_subset(slod_ctx *ctx) {
  if (!ctx) return SLOD_ERR_INVALID;
  return assemble_rows(ctx, true);
}

int slod_get_coarse_csr(const slod_ctx *ctx, int64_t *rowptr, int64_t *col, double *val, int64_t *n_rows, int64_t *nnz) {
  if (!ctx) return SLOD_ERR_INVALID;
  if (!ctx->coarse_done) return fail(ctx, SLOD_ERR_STATE, "slod_assemble_coarse has not run");
  if (n_rows) *n_rows = (int64_t)ctx->csr_rowptr.size() - 1;
  if (nnz) *nnz = (int64_t)ctx->csr_col.size();
  if (!rowptr) return SLOD_OK;
  if (!col || !val) return fail(ctx, SLOD_ERR_INVALID, "null output buffer");
  std::memcpy(rowptr, ctx->csr_rowptr.data(), sizeof(int64_t) * ctx->csr_rowptr.size());
  std::memcpy(col, ctx->csr_col.data(), sizeof(int64_t) * ctx->csr_col.size());
  std::memcpy(val, ctx->csr_val.data(), sizeof(double) * ctx->csr_val.size());
  return SLOD_OK;
}

int slod_get_patch_diagnostics(const slod_ctx *ctx, int64_t patch, int comp, double out[8]) {
  int rc = check_patch(ctx, patch);
  if (rc) return rc;
  if (!out || comp < 0 || comp >= ctx->s) return fail(ctx, SLOD_ERR_INVALID, "bad argument");
  const PatchOut &o = ctx->out[(size_t)patch];
  if (!o.done) return fail(ctx, SLOD_ERR_STATE, "patch has not been computed");
  std::memcpy(out, o.diag[comp], sizeof(double) * 8);
  return SLOD_OK;
}

int slod_debug_patch_stages(slod_ctx *ctx, int64_t patch, double *X, double *Minv, double *G) {
  int rc = check_patch(ctx, patch);
  if (rc) return rc;
  for (int f = 0; f < ctx->n_fields; ++f)
    if (!ctx->table_set[f]) return fail(ctx, SLOD_ERR_STATE, "coefficient field not set");
  std::vector<double> x, mi, g;
  Stages st;
  st.X = &x; st.Minv = &mi; st.G = &g;
  PatchOut tmp;
  compute_patch(ctx, (uint32_t)patch, tmp, &st);
  if (X) std::memcpy(X, x.data(), sizeof(double) * x.size());
  if (Minv) std::memcpy(Minv, mi.data(), sizeof(double) * mi.size());
  if (G) {
    if (g.empty()) return fail(ctx, SLOD_ERR_STATE, "patch takes the LOD branch: no Gram matrix");
    std::memcpy(G, g.data(), sizeof(double) * g.size());
  }
  return SLOD_OK;
}

int slod_get_timings(const slod_ctx *ctx, double *ms, int n) {
  if (!ctx || !ms) return SLOD_ERR_INVALID;
  for (int i = 0; i < n && i < 8; ++i) ms[i] = ctx->ms[i];
  return SLOD_OK;
}

int slod_launch_count(const slod_ctx *ctx, int64_t *n) {
  if (!ctx || !n) return SLOD_ERR_INVALID;
  *n = 0;   // no GPU kernels: this is the CPU baseline
  return SLOD_OK;
}

int slod_alloc_host(size_t bytes, void **out) {
  if (!out) return SLOD_ERR_INVALID;
  *out = std::malloc(std::max<size_t>(bytes, 1));
  return *out ? SLOD_OK : SLOD_ERR_INVALID;
}
int slod_free_host(void *p) {
  std::free(p);
  return SLOD_OK;
}

}  // extern "C"
