"""ctypes binding of ``libslod_b200.so`` (include/slod.h).

This is plumbing for the Python tests and ``bench.py``; the product is the shared library.  There is
no fallback: if the library is missing or no CUDA device is usable every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SLOD_LIB") or os.path.join(_HERE, "libslod_b200.so")   # SLOD_LIB: instrumented build (tools/)

SLOD_OK = 0
ERRORS = {1: "SLOD_ERR_INVALID", 2: "SLOD_ERR_UNSUPPORTED", 3: "SLOD_ERR_CUDA", 4: "SLOD_ERR_STATE",
          5: "SLOD_ERR_NUMERIC"}
PROBLEM_DIFFUSION, PROBLEM_ELASTICITY = 0, 1

# every symbol include/slod.h declares (tests check that the library exports all of them)
EXPORTS = [
    "slod_create", "slod_destroy", "slod_last_error", "slod_last_create_error", "slod_set_coefficient",
    "slod_patch_count", "slod_get_patch_info", "slod_get_patch_cells", "slod_get_patch_fine_dofs",
    "slod_get_patch_local_dofs", "slod_get_patch_dof_class", "slod_compute_basis", "slod_get_basis",
    "slod_basis_stride", "slod_get_all_basis", "slod_assemble_coarse", "slod_get_coarse_csr",
    "slod_get_patch_diagnostics", "slod_debug_patch_stages", "slod_get_timings", "slod_compute_basis_device",
    "slod_assemble_coarse_device", "slod_ell_width", "slod_ell_to_csr", "slod_launch_count", "slod_alloc_host",
    "slod_free_host", "slod_fine_size", "slod_coarse_rhs", "slod_coarse_solve", "slod_prolongate",
    "slod_fem_solve", "slod_fine_norms",
]


class SlodParams(C.Structure):
    _fields_ = [("dim", C.c_int), ("spacedim", C.c_int), ("n_global_refinements", C.c_int),
                ("n_subdivisions", C.c_int), ("oversampling", C.c_int), ("stabilize", C.c_int),
                ("problem", C.c_int), ("quirk_presaved", C.c_int), ("device", C.c_int)]


class SlodError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


_lib = None


def load_library():
    """Load libslod_b200.so (built in-tree by ``__graft_entry__.build()``); fails loudly when absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(the CUDA library is the product; there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    P = C.POINTER
    vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int32, C.c_double
    lib.slod_create.argtypes = [P(SlodParams), P(vp)]
    lib.slod_destroy.argtypes = [vp]
    lib.slod_destroy.restype = None
    lib.slod_last_error.argtypes = [vp]
    lib.slod_last_error.restype = C.c_char_p
    lib.slod_last_create_error.restype = C.c_char_p
    lib.slod_set_coefficient.argtypes = [vp, C.c_int, C.c_int, P(dbl), C.c_size_t]
    lib.slod_patch_count.argtypes = [vp, P(i64)]
    lib.slod_fine_size.argtypes = [vp, P(i64)]
    lib.slod_coarse_rhs.argtypes = [vp, P(dbl), P(dbl)]
    lib.slod_coarse_solve.argtypes = [vp, P(dbl), P(dbl), i32, dbl, dbl, P(i32), P(dbl)]
    lib.slod_prolongate.argtypes = [vp, P(dbl), P(dbl)]
    lib.slod_fem_solve.argtypes = [vp, P(dbl), P(dbl), i32, dbl, dbl, P(i32), P(dbl)]
    lib.slod_fine_norms.argtypes = [vp, P(dbl), P(dbl), P(dbl), P(dbl)]
    lib.slod_get_patch_info.argtypes = [vp, i64] + [P(i32)] * 6 + [P(i32), P(i32)]
    lib.slod_get_patch_cells.argtypes = [vp, i64, P(C.c_uint32), P(i32)]
    lib.slod_get_patch_fine_dofs.argtypes = [vp, i64, P(C.c_uint64), P(i32)]
    lib.slod_get_patch_local_dofs.argtypes = [vp, i64, P(C.c_uint32), P(i32)]
    lib.slod_get_patch_dof_class.argtypes = [vp, i64, C.c_int, P(C.c_uint32), P(i32)]
    lib.slod_compute_basis.argtypes = [vp]
    lib.slod_get_basis.argtypes = [vp, i64, C.c_int, P(dbl), P(dbl)]
    lib.slod_basis_stride.argtypes = [vp, P(i64)]
    lib.slod_get_all_basis.argtypes = [vp, P(dbl), P(dbl)]
    lib.slod_assemble_coarse.argtypes = [vp]
    lib.slod_get_coarse_csr.argtypes = [vp, P(i64), P(i64), P(dbl), P(i64), P(i64)]
    lib.slod_get_patch_diagnostics.argtypes = [vp, i64, C.c_int, P(dbl)]
    lib.slod_debug_patch_stages.argtypes = [vp, i64, P(dbl), P(dbl), P(dbl)]
    lib.slod_get_timings.argtypes = [vp, P(dbl), C.c_int]
    lib.slod_compute_basis_device.argtypes = [vp, i64, i64, vp, vp, vp]
    lib.slod_assemble_coarse_device.argtypes = [vp, i64, i64, vp, vp, vp, vp]
    lib.slod_ell_width.argtypes = [vp, P(i64)]
    lib.slod_ell_to_csr.argtypes = [vp, P(dbl), P(i64), P(i64), P(dbl), P(i64), P(i64)]
    lib.slod_launch_count.argtypes = [vp, P(i64)]
    lib.slod_alloc_host.argtypes = [C.c_size_t, P(vp)]
    lib.slod_free_host.argtypes = [vp]
    _lib = lib
    return lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class SlodContext:
    """Thin object wrapper over the opaque ``slod_ctx`` handle."""

    def __init__(self, dim=2, spacedim=1, n_global_refinements=2, n_subdivisions=2, oversampling=1,
                 stabilize=False, problem=PROBLEM_DIFFUSION, quirk_presaved=False, device=-1):
        self.lib = load_library()
        self.par = SlodParams(dim, spacedim, n_global_refinements, n_subdivisions, oversampling,
                              int(bool(stabilize)), problem, int(bool(quirk_presaved)), device)
        self.h = C.c_void_p()
        rc = self.lib.slod_create(C.byref(self.par), C.byref(self.h))
        if rc != SLOD_OK:
            raise SlodError(rc, self.lib.slod_last_create_error().decode())
        self.dim, self.s = dim, spacedim

    def close(self):
        for ptr in getattr(self, "_pinned_ptrs", []):
            self.lib.slod_free_host(ptr)
        self._pinned_ptrs, self._pinned = [], {}
        if getattr(self, "h", None) and self.h.value:
            self.lib.slod_destroy(self.h)
            self.h = C.c_void_p()

    def _out(self, key, shape, dtype=np.float64):
        """Output array in page-locked memory (slod_alloc_host), allocated once per key and reused by later calls."""
        if not hasattr(self, "_pinned"):
            self._pinned, self._pinned_ptrs = {}, []
        shape = tuple(int(x) for x in np.atleast_1d(shape))
        arr = self._pinned.get(key)
        if arr is not None and arr.shape == shape and arr.dtype == dtype:
            return arr
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        ptr = C.c_void_p()
        if self.par.device == -2 or self.lib.slod_alloc_host(nbytes, C.byref(ptr)) != SLOD_OK:
            arr = np.empty(shape, dtype=dtype)                       # maps-only handle / allocation refused
        else:
            self._pinned_ptrs.append(ptr)
            buf = (C.c_char * max(nbytes, 1)).from_address(ptr.value)
            arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        self._pinned[key] = arr
        return arr

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != SLOD_OK:
            raise SlodError(rc, self.lib.slod_last_error(self.h).decode())

    # -- inputs ----------------------------------------------------------------------------------------
    def set_coefficient(self, field, eta_refinement, values):
        v = np.ascontiguousarray(values, dtype=np.float64).ravel()
        self._ck(self.lib.slod_set_coefficient(self.h, field, eta_refinement, _dp(v), v.size))

    # -- integer maps ----------------------------------------------------------------------------------
    @property
    def n_patches(self):
        n = C.c_int64()
        self._ck(self.lib.slod_patch_count(self.h, C.byref(n)))
        return n.value

    def patch_info(self, patch):
        v = [C.c_int32() for _ in range(6)]
        lo = (C.c_int32 * 3)()
        m = (C.c_int32 * 3)()
        self._ck(self.lib.slod_get_patch_info(self.h, patch, *[C.byref(x) for x in v], lo, m))
        keys = ["n_cells", "n_fine", "n_internal", "n_boundary", "n_domain_boundary", "n_coarse"]
        out = {k: x.value for k, x in zip(keys, v)}
        out["lo"] = tuple(lo)[: self.dim]
        out["m"] = tuple(m)[: self.dim]
        return out

    def patch_cells(self, patch):
        n = C.c_int32()
        self._ck(self.lib.slod_get_patch_cells(self.h, patch, None, C.byref(n)))
        a = np.empty(n.value, dtype=np.uint32)
        self._ck(self.lib.slod_get_patch_cells(self.h, patch, a.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(n)))
        return a

    def patch_fine_dofs(self, patch):
        n = C.c_int32()
        self._ck(self.lib.slod_get_patch_fine_dofs(self.h, patch, None, C.byref(n)))
        a = np.empty(n.value, dtype=np.uint64)
        self._ck(self.lib.slod_get_patch_fine_dofs(self.h, patch, a.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(n)))
        return a

    def patch_local_dofs(self, patch):
        n = C.c_int32()
        self._ck(self.lib.slod_get_patch_local_dofs(self.h, patch, None, C.byref(n)))
        a = np.empty(n.value, dtype=np.uint32)
        self._ck(self.lib.slod_get_patch_local_dofs(self.h, patch, a.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(n)))
        return a

    def patch_dof_class(self, patch, which):
        n = C.c_int32()
        self._ck(self.lib.slod_get_patch_dof_class(self.h, patch, which, None, C.byref(n)))
        a = np.empty(n.value, dtype=np.uint32)
        self._ck(self.lib.slod_get_patch_dof_class(self.h, patch, which, a.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                   C.byref(n)))
        return a

    # -- hot path, host buffers --------------------------------------------------------------------------
    def compute_basis(self):
        self._ck(self.lib.slod_compute_basis(self.h))

    def basis(self, patch, comp=0):
        nf = self.patch_info(patch)["n_fine"]
        phi = np.empty(nf)
        aphi = np.empty(nf)
        self._ck(self.lib.slod_get_basis(self.h, patch, comp, _dp(phi), _dp(aphi)))
        return phi, aphi

    @property
    def basis_stride(self):
        n = C.c_int64()
        self._ck(self.lib.slod_basis_stride(self.h, C.byref(n)))
        return n.value

    def all_basis(self):
        shape = (self.n_patches, self.s, self.basis_stride)
        phi = self._out("phi", shape)
        aphi = self._out("aphi", shape)
        self._ck(self.lib.slod_get_all_basis(self.h, _dp(phi), _dp(aphi)))
        return phi, aphi

    def assemble_coarse(self):
        self._ck(self.lib.slod_assemble_coarse(self.h))

    def coarse_csr(self):
        nr, nnz = C.c_int64(), C.c_int64()
        self._ck(self.lib.slod_get_coarse_csr(self.h, None, None, None, C.byref(nr), C.byref(nnz)))
        rowptr = self._out("rowptr", nr.value + 1, np.int64)
        col = self._out("col", nnz.value, np.int64)
        val = self._out("val", nnz.value)
        i64p = C.POINTER(C.c_int64)
        self._ck(self.lib.slod_get_coarse_csr(self.h, rowptr.ctypes.data_as(i64p), col.ctypes.data_as(i64p),
                                              _dp(val), C.byref(nr), C.byref(nnz)))
        return rowptr, col, val

    # -- online phase (LOD::solve, source/LOD.cc:975-1001; prolongation :1251) --
    @property
    def n_fine(self):
        n = C.c_int64()
        self._ck(self.lib.slod_fine_size(self.h, C.byref(n)))
        return n.value

    def coarse_rhs(self, f_fine):
        f = np.ascontiguousarray(f_fine, dtype=np.float64).ravel()
        if f.size != self.n_fine:
            raise ValueError(f"fine vector has {f.size} entries, expected {self.n_fine}")
        b = np.empty(self.n_patches * self.s)
        self._ck(self.lib.slod_coarse_rhs(self.h, _dp(f), _dp(b)))
        return b

    def coarse_solve(self, rhs, max_steps=10000, tolerance=1e-10, reduction=1e-10):
        """Returns (u, steps, residual); raises SlodError (SLOD_ERR_NUMERIC) when CG does not converge."""
        b = np.ascontiguousarray(rhs, dtype=np.float64).ravel()
        if b.size != self.n_patches * self.s:
            raise ValueError("coarse vector has the wrong size")
        u = np.empty_like(b)
        steps, res = C.c_int32(), C.c_double()
        self._ck(self.lib.slod_coarse_solve(self.h, _dp(b), _dp(u), max_steps, tolerance, reduction,
                                            C.byref(steps), C.byref(res)))
        return u, steps.value, res.value

    def prolongate(self, u_coarse):
        u = np.ascontiguousarray(u_coarse, dtype=np.float64).ravel()
        if u.size != self.n_patches * self.s:
            raise ValueError("coarse vector has the wrong size")
        out = np.empty(self.n_fine)
        self._ck(self.lib.slod_prolongate(self.h, _dp(u), _dp(out)))
        return out

    # -- fine-scale reference problem (assemble_and_solve_fem_problem, source/LOD.cc:1004-1094) and norms (:1252) --
    def fem_solve(self, f_fine, max_steps=200000, tolerance=1e-12, reduction=1e-10):
        """Returns (u_fine, steps, residual)."""
        f = np.ascontiguousarray(f_fine, dtype=np.float64).ravel()
        if f.size != self.n_fine:
            raise ValueError(f"fine vector has {f.size} entries, expected {self.n_fine}")
        u = np.empty_like(f)
        steps, res = C.c_int32(), C.c_double()
        self._ck(self.lib.slod_fem_solve(self.h, _dp(f), _dp(u), max_steps, tolerance, reduction,
                                         C.byref(steps), C.byref(res)))
        return u, steps.value, res.value

    def fine_norms(self, v_fine):
        """(L2 norm, H1 seminorm, energy norm) of a fine vector."""
        v = np.ascontiguousarray(v_fine, dtype=np.float64).ravel()
        if v.size != self.n_fine:
            raise ValueError(f"fine vector has {v.size} entries, expected {self.n_fine}")
        l2, h1, en = C.c_double(), C.c_double(), C.c_double()
        self._ck(self.lib.slod_fine_norms(self.h, _dp(v), C.byref(l2), C.byref(h1), C.byref(en)))
        return l2.value, h1.value, en.value

    def diagnostics(self, patch, comp=0):
        out = np.empty(8)
        self._ck(self.lib.slod_get_patch_diagnostics(self.h, patch, comp, _dp(out)))
        return out

    def debug_stages(self, patch, want_G=True):
        info = self.patch_info(patch)
        ni, ncd = info["n_internal"], info["n_coarse"]
        X = np.empty((ni, ncd))
        Minv = np.empty((ncd, ncd))
        G = np.empty((ncd, ncd)) if want_G else None
        self._ck(self.lib.slod_debug_patch_stages(self.h, patch, _dp(X), _dp(Minv), _dp(G) if want_G else None))
        return X, Minv, G

    def timings(self):
        out = np.zeros(8)
        self._ck(self.lib.slod_get_timings(self.h, _dp(out), 8))
        return out

    @property
    def launch_count(self):
        n = C.c_int64()
        self._ck(self.lib.slod_launch_count(self.h, C.byref(n)))
        return n.value

    @property
    def ell_width(self):
        n = C.c_int64()
        self._ck(self.lib.slod_ell_width(self.h, C.byref(n)))
        return n.value

    # -- hot path, device buffers (pointers are plain integers, e.g. torch.Tensor.data_ptr()) -----------
    def compute_basis_device(self, p0, p1, d_phi, d_aphi, stream=0):
        self._ck(self.lib.slod_compute_basis_device(self.h, p0, p1, C.c_void_p(d_phi), C.c_void_p(d_aphi),
                                                    C.c_void_p(stream)))

    def assemble_coarse_device(self, p0, p1, d_phi, d_aphi, d_K, stream=0):
        self._ck(self.lib.slod_assemble_coarse_device(self.h, p0, p1, C.c_void_p(d_phi), C.c_void_p(d_aphi),
                                                      C.c_void_p(d_K), C.c_void_p(stream)))

    def ell_to_csr(self, h_K):
        h_K = np.ascontiguousarray(h_K, dtype=np.float64)
        nr, nnz = C.c_int64(), C.c_int64()
        self._ck(self.lib.slod_ell_to_csr(self.h, None, None, None, None, C.byref(nr), C.byref(nnz)))
        rowptr = np.empty(nr.value + 1, dtype=np.int64)
        col = np.empty(nnz.value, dtype=np.int64)
        val = np.empty(nnz.value)
        i64p = C.POINTER(C.c_int64)
        self._ck(self.lib.slod_ell_to_csr(self.h, _dp(h_K), rowptr.ctypes.data_as(i64p), col.ctypes.data_as(i64p),
                                          _dp(val), C.byref(nr), C.byref(nnz)))
        return rowptr, col, val
