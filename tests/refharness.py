"""ctypes view of oracle/_ref/libldm_ref*.so -- the reference compiled from its own sources.

TEST INFRASTRUCTURE ONLY (used by tests/, tests/golden/make_golden.py and bench.py's reference arm).
"""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DATA = "/root/reference/Euler Solver ver.5.3/Data"

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def _load(name):
    path = os.path.join(ROOT, "oracle", "_ref", name)
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    lib.ref_props_from_file.restype = C.c_void_p
    lib.ref_props_from_file.argtypes = [C.c_char_p, C.c_int, C.c_double]
    lib.ref_props_from_numbers.restype = C.c_void_p
    lib.ref_props_from_numbers.argtypes = [C.c_double] * 5 + [C.c_int]
    lib.ref_props_free.argtypes = [C.c_void_p]
    lib.ref_props_scalars.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)]
    lib.ref_props_grids.argtypes = [C.c_void_p, _dp, _dp]
    lib.ref_props_tables.argtypes = [C.c_void_p, _dp, _dp]
    lib.ref_props_multi.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, _dp, _dp]
    lib.ref_rhs.argtypes = [C.c_void_p, _dp, _dp]
    lib.ref_integrals.argtypes = [C.c_void_p, _dp, C.POINTER(C.c_double), C.c_void_p]
    lib.ref_kernels_at_nodes.argtypes = [C.c_void_p, _dp, _dp]
    if hasattr(lib, "ref_mc_rows"):
        lib.ref_mc_rows.restype = C.c_int
        lib.ref_mc_rows.argtypes = [C.c_void_p, _dp, C.c_int, C.c_int, _dp]
    lib.ref_solve.restype = C.c_int
    lib.ref_solve.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p]
    lib.ref_solve_rk4.restype = C.c_int
    lib.ref_solve_rk4.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p]
    lib.ref_set_extrapolation.argtypes = [C.c_int]
    lib.ref_result_times.argtypes = [C.c_void_p, _dp]
    lib.ref_result_point.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    return lib


class RefProblem:
    """One Properties object of the reference (Solver.cpp:40-242) plus the calls made on it."""

    def __init__(self, lib, handle):
        self.lib, self.h = lib, handle
        ints = (C.c_int * 5)()
        reals = (C.c_double * 7)()
        lib.ref_props_scalars(handle, ints, reals)
        self.species, self.G, self.K, self.mM, self.mW = list(ints)
        self.h_step, self.cutoff, self.sigmaM, self.sigmaW, self.b, self.d, self.dd = list(reals)
        self.P = self.G * self.G
        self.width = self.species + self.species ** 2 * self.P          # state vector length (Solver.cpp:260)

    def grids(self):
        g, k = np.zeros(self.G), np.zeros(self.K)
        self.lib.ref_props_grids(self.h, g, k)
        return g, k

    def tables(self):
        M, W = np.zeros((self.K, self.K)), np.zeros((self.K, self.K))
        self.lib.ref_props_tables(self.h, M, W)
        return M, W

    def multi_tables(self):
        """several species: MaxNumForM (s), MaxNumForW (s, s), M (s, K, K), W (s, s, K, K)"""
        s, K = self.species, self.K
        mM, mW = np.zeros(s, dtype=np.int32), np.zeros((s, s), dtype=np.int32)
        M, W = np.zeros((s, K, K)), np.zeros((s, s, K, K))
        self.lib.ref_props_multi(self.h, mM.ctypes.data, mW.ctypes.data, M, W)
        return mM, mW, M, W

    def rhs(self, state):
        state = np.ascontiguousarray(state, dtype=np.float64)
        out = np.zeros_like(state)
        self.lib.ref_rhs(self.h, state, out)
        return out

    def integrals(self, state, want_mc=True):
        state = np.ascontiguousarray(state, dtype=np.float64)
        wc = C.c_double()
        mc = np.zeros((self.G, self.G)) if want_mc else None
        self.lib.ref_integrals(self.h, state, C.byref(wc), mc.ctypes.data if want_mc else None)
        return wc.value, mc

    def mc_rows(self, state, first, stride):
        """the reference's per-node integral (Solver.cpp:297-316) on the rows first, first + stride, ...: [rows, G]"""
        state = np.ascontiguousarray(state, dtype=np.float64)
        out = np.zeros((len(range(first, self.G, stride)), self.G))
        n = self.lib.ref_mc_rows(self.h, state, int(first), int(stride), out)
        assert n == out.shape[0]
        return out

    def kernels_at_nodes(self):
        m, w = np.zeros((self.G, self.G)), np.zeros((self.G, self.G))
        self.lib.ref_kernels_at_nodes(self.h, m, w)
        return m, w

    def set_extrapolation(self, kind):
        """ExtrapolationType (Solver.cpp:28) of the following rhs / solve calls: 0 Zero, 1 SquaredN, 2 Alpha"""
        self.lib.ref_set_extrapolation(int(kind))

    def solve(self, t_max, dt=0.1, n0=200.0, init=None, want_derivs=False, rk4=False):
        if init is not None:
            init = np.ascontiguousarray(init, dtype=np.float64)
        fn = self.lib.ref_solve_rk4 if rk4 else self.lib.ref_solve
        n = fn(self.h, t_max, dt, n0, init.ctypes.data if init is not None else None)
        times = np.zeros(n)
        self.lib.ref_result_times(self.h, times)
        values = np.zeros((n, self.width))
        derivs = np.zeros((n, self.width)) if want_derivs else None
        for k in range(n):
            self.lib.ref_result_point(self.h, k, values[k].ctypes.data,
                                      derivs[k].ctypes.data if want_derivs else None)
        return times, values, derivs

    def close(self):
        if self.h:
            self.lib.ref_props_free(self.h)
            self.h = None

    def __del__(self):
        self.close()


class Reference:
    def __init__(self, byref=False):
        self.lib = _load("libldm_ref_byref.so" if byref else "libldm_ref.so")
        self.kind = "byref" if byref else "unmodified"

    @property
    def available(self):
        return self.lib is not None

    def from_file(self, path, route=0, sigma=0.0):
        return RefProblem(self.lib, self.lib.ref_props_from_file(path.encode(), route, sigma))

    def from_numbers(self, sigma_w, sigma_m, b=0.4, d=0.2, dd=0.001, forced_G=0):
        return RefProblem(self.lib, self.lib.ref_props_from_numbers(sigma_w, sigma_m, b, d, dd, forced_G))
