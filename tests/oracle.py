"""ctypes view of oracle/libldm_oracle.so -- the plain-C CPU oracle.  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


class _Problem(C.Structure):
    _fields_ = [("G", C.c_int), ("K", C.c_int), ("mM", C.c_int), ("mW", C.c_int),
                ("h", C.c_double), ("cutoff", C.c_double), ("sigmaM", C.c_double), ("sigmaW", C.c_double),
                ("b", C.c_double), ("d", C.c_double), ("dd", C.c_double),
                ("moments_grid", C.POINTER(C.c_double)), ("kernels_grid", C.POINTER(C.c_double)),
                ("M", C.POINTER(C.c_double)), ("W", C.POINTER(C.c_double))]


def _load():
    path = os.path.join(ROOT, "oracle", "libldm_oracle.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "libldm_oracle.so"])
    lib = C.CDLL(path)
    lib.ldmo_generate.restype = C.POINTER(_Problem)
    lib.ldmo_generate.argtypes = [C.c_double] * 5 + [C.c_int]
    lib.ldmo_free.argtypes = [C.POINTER(_Problem)]
    lib.ldmo_integrals.argtypes = [C.POINTER(_Problem), _dp, C.POINTER(C.c_double), C.c_void_p]
    lib.ldmo_kernels_at_nodes.argtypes = [C.POINTER(_Problem), _dp, _dp]
    lib.ldmo_rhs.argtypes = [C.POINTER(_Problem), _dp, _dp]
    lib.ldmo_count_points.restype = C.c_int
    lib.ldmo_count_points.argtypes = [C.c_double, C.c_double]
    lib.ldmo_solve.restype = C.c_int
    lib.ldmo_solve.argtypes = [C.POINTER(_Problem), C.c_double, C.c_double, _dp, C.c_int,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ldmo_set_threads.argtypes = [C.c_int]
    lib.ldmo_set_extrapolation.argtypes = [C.c_int]
    lib.ldmo_count_points_rk4.restype = C.c_int
    lib.ldmo_count_points_rk4.argtypes = [C.c_double, C.c_double]
    lib.ldmo_solve_rk4.restype = C.c_int
    lib.ldmo_solve_rk4.argtypes = [C.POINTER(_Problem), C.c_double, C.c_double, _dp, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ldmo_w_mass.restype = C.c_double
    lib.ldmo_w_mass.argtypes = [C.POINTER(_Problem)]
    lib.ldmo_closure_integrals.argtypes = [C.POINTER(_Problem), _dp, C.c_void_p, C.c_void_p]
    lib.ldmo_rhs_closure.argtypes = [C.POINTER(_Problem), _dp, C.c_int, _dp]
    lib.ldmo_solve_closure.restype = C.c_int
    lib.ldmo_solve_closure.argtypes = [C.POINTER(_Problem), C.c_int, C.c_double, C.c_double, _dp, C.c_void_p, C.c_void_p]
    return lib


class _Multi(C.Structure):
    _fields_ = [("species", C.c_int), ("G", C.c_int), ("K", C.c_int), ("h", C.c_double), ("cutoff", C.c_double),
                ("mM", C.c_void_p), ("mW", C.c_void_p), ("sigmaM", C.c_void_p), ("sigmaW", C.c_void_p), ("b", C.c_void_p),
                ("d", C.c_void_p), ("dd", C.c_void_p), ("moments_grid", C.c_void_p), ("kernels_grid", C.c_void_p),
                ("M", C.c_void_p), ("W", C.c_void_p)]


class OracleMulti:
    """several species (ldm_oracle.h ldmo_multi): the tables are handed in (from the product's host generator or a fixture)"""

    def __init__(self, mM, mW, sigmaM, sigmaW, b, d, dd, moments_grid, kernels_grid, M, W):
        self.lib = lib()
        self.lib.ldmo_rhs_multi.argtypes = [C.POINTER(_Multi), _dp, _dp]
        self.lib.ldmo_solve_multi.restype = C.c_int
        self.lib.ldmo_solve_multi.argtypes = [C.POINTER(_Multi), C.c_double, C.c_double, _dp, C.c_void_p, C.c_void_p]
        f64 = lambda v: np.ascontiguousarray(v, dtype=np.float64)
        i32 = lambda v: np.ascontiguousarray(v, dtype=np.int32)
        self.keep = [i32(mM), i32(mW), f64(sigmaM), f64(sigmaW), f64(b), f64(d), f64(dd), f64(moments_grid), f64(kernels_grid), f64(M), f64(W)]
        self.s, self.G, self.K = len(self.keep[0]), len(self.keep[7]), len(self.keep[8])
        kg = self.keep[8]
        self.p = _Multi(self.s, self.G, self.K, kg[2] - kg[0], self.keep[7][-1], *[a.ctypes.data for a in self.keep])
        self.width = self.s + self.s ** 2 * self.G ** 2

    def rhs(self, state):
        state = np.ascontiguousarray(state, dtype=np.float64)
        out = np.zeros_like(state)
        self.lib.ldmo_rhs_multi(C.byref(self.p), state, out)
        return out

    def solve(self, t_max, dt, init):
        n = self.lib.ldmo_count_points(t_max, dt)
        times, values = np.zeros(n), np.zeros((n, self.width))
        got = self.lib.ldmo_solve_multi(C.byref(self.p), t_max, dt, np.ascontiguousarray(init, dtype=np.float64),
                                        times.ctypes.data, values.ctypes.data)
        assert got == n
        return times, values


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        _LIB = _load()
    return _LIB


class OracleProblem:
    """ldmo_problem: Properties::Generate for one species (Solver.cpp:186-241), restated."""

    def __init__(self, sigma_w, sigma_m, b=0.4, d=0.2, dd=0.001, forced_G=0):
        self.lib = lib()
        self.p = self.lib.ldmo_generate(sigma_w, sigma_m, b, d, dd, forced_G)
        c = self.p.contents
        self.G, self.K, self.mM, self.mW = c.G, c.K, c.mM, c.mW
        self.h_step, self.cutoff = c.h, c.cutoff
        self.sigmaM, self.sigmaW, self.b, self.d, self.dd = c.sigmaM, c.sigmaW, c.b, c.d, c.dd
        self.P = self.G * self.G

    def grids(self):
        c = self.p.contents
        return (np.ctypeslib.as_array(c.moments_grid, (self.G,)).copy(),
                np.ctypeslib.as_array(c.kernels_grid, (self.K,)).copy())

    def tables(self):
        c = self.p.contents
        return (np.ctypeslib.as_array(c.M, (self.K, self.K)).copy(),
                np.ctypeslib.as_array(c.W, (self.K, self.K)).copy())

    def equilibrium(self, n0=200.0):
        """GenerateEquilibriumValuesForNAndC (Solver.cpp:416-428)"""
        st = np.full(1 + self.P, n0 * n0)
        st[0] = n0
        return st

    def rhs(self, state):
        state = np.ascontiguousarray(state, dtype=np.float64)
        out = np.zeros_like(state)
        self.lib.ldmo_rhs(self.p, state, out)
        return out

    def integrals(self, state, want_mc=True):
        state = np.ascontiguousarray(state, dtype=np.float64)
        wc = C.c_double()
        mc = np.zeros((self.G, self.G)) if want_mc else None
        self.lib.ldmo_integrals(self.p, state, C.byref(wc), mc.ctypes.data if want_mc else None)
        return wc.value, mc

    def kernels_at_nodes(self):
        m, w = np.zeros((self.G, self.G)), np.zeros((self.G, self.G))
        self.lib.ldmo_kernels_at_nodes(self.p, m, w)
        return m, w

    def count_points(self, t_max, dt=0.1):
        return self.lib.ldmo_count_points(t_max, dt)

    def solve(self, t_max, dt=0.1, init=None, full_history=True, want_derivs=False):
        init = self.equilibrium() if init is None else np.ascontiguousarray(init, dtype=np.float64)
        n = self.count_points(t_max, dt)
        times = np.zeros(n)
        rows = n if full_history else 1
        values = np.zeros((rows, 1 + self.P))
        derivs = np.zeros((rows, 1 + self.P)) if want_derivs else None
        obs = np.zeros((n, 2))
        got = self.lib.ldmo_solve(self.p, t_max, dt, init, 1 if full_history else 0, times.ctypes.data,
                                  values.ctypes.data, derivs.ctypes.data if want_derivs else None, obs.ctypes.data)
        assert got == n
        return times, values, derivs, obs

    def set_extrapolation(self, kind):
        """ExtrapolationType of the following calls (process-wide in the oracle): 0 Zero, 1 SquaredN, 2 Alpha"""
        self.lib.ldmo_set_extrapolation(int(kind))

    def solve_rk4(self, t_max, dt=0.1, init=None):
        """RungeKutta4thSolver::Solve (Solver.cpp:597-654): full history (times, values, derivatives)"""
        init = self.equilibrium() if init is None else np.ascontiguousarray(init, dtype=np.float64)
        n = self.lib.ldmo_count_points_rk4(t_max, dt)
        times, values, derivs = np.zeros(n), np.zeros((n, 1 + self.P)), np.zeros((n, 1 + self.P))
        got = self.lib.ldmo_solve_rk4(self.p, t_max, dt, init, times.ctypes.data, values.ctypes.data, derivs.ctypes.data)
        assert got == n
        return times, values, derivs

    # ---- extension, parity unpinned beyond the uniform state (oracle/ldm_oracle_closure.c) ----
    def w_mass(self):
        return self.lib.ldmo_w_mass(self.p)

    def closure_integrals(self, state):
        """(w*C)(xi) and K(xi) = integral of w(xi') C(xi') C(xi + xi') at every node"""
        state = np.ascontiguousarray(state, dtype=np.float64)
        wc, k = np.zeros((self.G, self.G)), np.zeros((self.G, self.G))
        self.lib.ldmo_closure_integrals(self.p, state, wc.ctypes.data, k.ctypes.data)
        return wc, k

    def rhs_closure(self, state, closure):
        state = np.ascontiguousarray(state, dtype=np.float64)
        out = np.zeros_like(state)
        self.lib.ldmo_rhs_closure(self.p, state, int(closure), out)
        return out

    def solve_closure(self, closure, t_max, dt=0.1, init=None):
        init = self.equilibrium() if init is None else np.ascontiguousarray(init, dtype=np.float64)
        n = self.count_points(t_max, dt)
        times, values = np.zeros(n), np.zeros((n, 1 + self.P))
        got = self.lib.ldmo_solve_closure(self.p, int(closure), t_max, dt, init, times.ctypes.data, values.ctypes.data)
        assert got == n
        return times, values

    def close(self):
        if self.p:
            self.lib.ldmo_free(self.p)
            self.p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
