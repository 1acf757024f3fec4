"""Loader of the C++ CPU restatement (oracle/cpu/slod_cpu.cc -> oracle/_build/libslod_cpu.so) for tests/ and bench.py.
Test / baseline infrastructure: the product package never imports this."""
import ctypes as C
import importlib
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPU_LIB = os.path.join(ROOT, "oracle", "_build", "libslod_cpu.so")


def build_cpu_port(force=False):
    src = os.path.join(ROOT, "oracle", "cpu")
    if force:
        subprocess.run(["make", "-C", src, "clean"], capture_output=True)
    res = subprocess.run(["make", "-C", src], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("oracle/cpu build failed:\n" + res.stdout + res.stderr)
    return CPU_LIB


class CpuSlod:
    """SlodContext of the product binding running on libslod_cpu.so, plus the port's two extensions
    (compute a list of patches, fetch one row of the coarse matrix)."""

    def __init__(self, threads=None, **kw):
        pkg = importlib.import_module("dealii-slod_b200")
        if not os.path.exists(CPU_LIB):
            build_cpu_port()
        kw.pop("device", None)
        self.ctx = pkg.SlodContext(lib=CPU_LIB, device=-2, **kw)
        cdll = self.ctx.lib._cdll
        cdll.slod_cpu_compute_patches.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_int64]
        cdll.slod_cpu_get_coarse_row.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_double),
                                                 C.POINTER(C.c_int32)]
        cdll.slod_cpu_set_threads.argtypes = [C.c_void_p, C.c_int]
        cdll.slod_cpu_get_threads.argtypes = [C.c_void_p]
        self._cdll = cdll
        if threads:
            self.ctx._ck(cdll.slod_cpu_set_threads(self.ctx.h, int(threads)))

    def __getattr__(self, name):
        return getattr(self.ctx, name)

    @property
    def threads(self):
        return self._cdll.slod_cpu_get_threads(self.ctx.h)

    def compute_patches(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        self.ctx._ck(self._cdll.slod_cpu_compute_patches(self.ctx.h, ids.ctypes.data_as(C.POINTER(C.c_int64)), ids.size))

    def assemble_coarse_subset(self):
        """Rows of the patches computed so far, columns restricted to computed neighbours (bounded samples)."""
        self._cdll.slod_cpu_assemble_coarse_subset.argtypes = [C.c_void_p]
        self.ctx._ck(self._cdll.slod_cpu_assemble_coarse_subset(self.ctx.h))

    def coarse_row(self, row):
        n = C.c_int32()
        self.ctx._ck(self._cdll.slod_cpu_get_coarse_row(self.ctx.h, row, None, None, C.byref(n)))
        col = np.empty(n.value, dtype=np.int64)
        val = np.empty(n.value)
        self.ctx._ck(self._cdll.slod_cpu_get_coarse_row(self.ctx.h, row, col.ctypes.data_as(C.POINTER(C.c_int64)),
                                                       val.ctypes.data_as(C.POINTER(C.c_double)), C.byref(n)))
        return col, val
