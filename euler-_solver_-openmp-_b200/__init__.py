"""Python view of libldm_b200.so (include/ldm.h) that mirrors the reference's solver seam.

The reference (Kintobor1370/Euler-_Solver_-OpenMP-) is C++; its operator interface for the hot path is
`Properties::Generate` (Solver.cpp:186), `Equation(Properties)` + `Equation::operator()` (Solver.cpp:413, 431)
and `EulerSolver::Solve` (Solver.cpp:557) returning `SolverResults{Times, Values, Derivatives}`
(Solver.cpp:525).  The classes below keep those names and argument meanings on top of the C ABI so that
tests and benchmarks read like calls into the reference.  All compute happens in the CUDA library: if
it is missing, or no GPU is present, the calls raise -- there is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIBRARY_PATH = os.environ.get("LDM_LIBRARY", os.path.join(_HERE, "libldm_b200.so"))   # LDM_LIBRARY: debug builds only

ALGO_AUTO, ALGO_DIRECT, ALGO_FACTORED, ALGO_PREFIX = 0, 1, 2, 3
RECORD_LAST, RECORD_OBSERVABLES, RECORD_VALUES, RECORD_ALL = 0, 1, 2, 3

# enum ExtrapolationType (Solver.cpp:28); every experiment passes SquaredN for C; the other three are built to the letter
# of the reference (ldm.h)
Zero, SquaredN, Alpha, Exponential = 0, 1, 2, 3
# closure of the third moment (ldm.h ldm_set_closure): the reference's power-1, and the two whose integral does not factorise
CLOSURE_POWER1, CLOSURE_POWER2_SYMMETRIC, CLOSURE_POWER3 = 1, 2, 3


class LdmError(RuntimeError):
    pass


class _Problem(C.Structure):
    _fields_ = [("species", C.c_int), ("G", C.c_int), ("K", C.c_int), ("mM", C.c_int), ("mW", C.c_int),
                ("h", C.c_double), ("cutoff", C.c_double), ("sigma_m", C.c_double), ("sigma_w", C.c_double),
                ("b", C.c_double), ("d", C.c_double), ("d_prime", C.c_double),
                ("moments_grid", C.POINTER(C.c_double)), ("kernels_grid", C.POINTER(C.c_double)),
                ("M", C.POINTER(C.c_double)), ("W", C.POINTER(C.c_double)),
                ("mM_of", C.POINTER(C.c_int)), ("mW_of", C.POINTER(C.c_int)),
                ("sigma_m_of", C.POINTER(C.c_double)), ("sigma_w_of", C.POINTER(C.c_double)),
                ("b_of", C.POINTER(C.c_double)), ("d_of", C.POINTER(C.c_double)), ("d_prime_of", C.POINTER(C.c_double))]


_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_lib = None

# every symbol include/ldm.h declares: name -> (restype, argtypes)
_SIGNATURES = {
    "ldm_problem_generate": (C.c_int, [C.c_double] * 5 + [C.c_int, C.POINTER(C.POINTER(_Problem))]),
    "ldm_problem_generate_multi": (C.c_int, [C.c_int, _dp, _dp, _dp, _dp, _dp, C.c_int, C.POINTER(C.POINTER(_Problem))]),
    "ldm_problem_free": (None, [C.POINTER(_Problem)]),
    "ldm_problem_check_kernels": (C.c_int, [C.POINTER(_Problem), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "ldm_problem_kernel_structure": (C.c_int, [C.POINTER(_Problem), C.POINTER(C.c_int), C.POINTER(C.c_double)]),
    "ldm_create": (C.c_int, [C.POINTER(_Problem), C.c_int, C.POINTER(C.c_void_p)]),
    "ldm_destroy": (None, [C.c_void_p]),
    "ldm_set_algorithm": (C.c_int, [C.c_void_p, C.c_int]),
    "ldm_get_algorithm": (C.c_int, [C.c_void_p]),
    "ldm_set_extrapolation": (C.c_int, [C.c_void_p, C.c_int]),
    "ldm_set_closure": (C.c_int, [C.c_void_p, C.c_int]),
    "ldm_get_closure": (C.c_int, [C.c_void_p]),
    "ldm_closure_integrals": (C.c_int, [C.c_void_p, _dp, C.c_void_p, C.c_void_p]),
    "ldm_count_points_rk4": (C.c_int, [C.c_double, C.c_double]),
    "ldm_solve_rk4": (C.c_int, [C.c_void_p, _dp, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]),
    "ldm_rhs": (C.c_int, [C.c_void_p, _dp, _dp]),
    "ldm_integrals": (C.c_int, [C.c_void_p, _dp, C.POINTER(C.c_double), C.c_void_p]),
    "ldm_kernels_at_nodes": (C.c_int, [C.c_void_p, _dp, _dp]),
    "ldm_count_points": (C.c_int, [C.c_double, C.c_double]),
    "ldm_solve": (C.c_int, [C.c_void_p, _dp, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                            C.POINTER(C.c_int)]),
    "ldm_state_upload": (C.c_int, [C.c_void_p, _dp]),
    "ldm_state_download": (C.c_int, [C.c_void_p, _dp]),
    "ldm_advance": (C.c_int, [C.c_void_p, C.c_int, C.c_double]),
    "ldm_synchronize": (C.c_int, [C.c_void_p]),
    "ldm_last_elapsed_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "ldm_launch_count": (C.c_longlong, [C.c_void_p]),
    "ldm_device_count": (C.c_int, []),
    "ldm_solve_batch": (C.c_int, [C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_double, C.c_double,
                                  C.POINTER(C.c_void_p)]),
    "ldm_shard_configure": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "ldm_shard_rows_per_rank": (C.c_int, [C.c_void_p]),
    "ldm_shard_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]),
    "ldm_shard_set_N": (C.c_int, [C.c_void_p, C.c_double]),
    "ldm_shard_get_N": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "ldm_shard_peer_alloc": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ldm_shard_peer_connect": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ldm_shard_peer_upload": (C.c_int, [C.c_void_p, _dp, C.c_double]),
    "ldm_shard_peer_step": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p]),
    "ldm_shard_peer_advance": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_void_p]),
    "ldm_shard_peer_connect_local": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "ldm_tuning_set": (C.c_int, [C.c_char_p, C.c_int]),
    "ldm_shard_peer_download": (C.c_int, [C.c_void_p, _dp, C.POINTER(C.c_double)]),
    "ldm_measure_fp64_peak": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.c_void_p]),
    "ldm_measure_memory_pipes": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "ldm_last_error": (C.c_char_p, []),
    "ldm_version": (C.c_char_p, []),
}


def library():
    """Load libldm_b200.so (built in-tree by `make -C euler-_solver_-openmp-_b200`); raises if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIBRARY_PATH):
            raise LdmError(f"{LIBRARY_PATH} is missing: build it with __graft_entry__.build() -- there is no CPU fallback")
        lib = C.CDLL(LIBRARY_PATH)
        for name, (restype, argtypes) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = restype, argtypes
        _lib = lib
    return _lib


def _check(code):
    if code != 0:
        raise LdmError(f"ldm error {code}: {library().ldm_last_error().decode()}")


def tuning_set(name, value=0):
    """test hook (ldm.h ldm_tuning_set): force a kernel shape for the handles created from here on; tuning_set("reset")"""
    _check(library().ldm_tuning_set(name.encode(), int(value)))


def measure_fp64_peak(device=0):
    """DFMA peak of the device in TFLOP/s (live micro-benchmark; the FP64 roofline denominator)."""
    v = C.c_double()
    _check(library().ldm_measure_fp64_peak(int(device), C.byref(v), None))
    return v.value


def measure_memory_pipes(device=0):
    """(load/store data-pipe wavefronts per second over the whole device, L2 -> SM read bandwidth in GB/s), measured live"""
    wf, l2 = C.c_double(), C.c_double()
    _check(library().ldm_measure_memory_pipes(int(device), C.byref(wf), C.byref(l2)))
    return wf.value, l2.value


class Properties:
    """struct Properties (Solver.cpp:40-242), one species: grids, kernel tables and rates, host side."""

    def __init__(self, ptr):
        self._ptr = ptr
        c = ptr.contents
        self.SpeciesNumber = c.species
        self.G, self.K = c.G, c.K
        self.MaxNumForM, self.MaxNumForW = c.mM, c.mW
        self.KernelsGridStep, self.CutoffDistance = c.h, c.cutoff
        self.SigmaM, self.SigmaW, self.b, self.d, self.d_prime = c.sigma_m, c.sigma_w, c.b, c.d, c.d_prime
        self.MomentsGrid = np.ctypeslib.as_array(c.moments_grid, (c.G,)).copy()
        self.KernelsGrid = np.ctypeslib.as_array(c.kernels_grid, (c.K,)).copy()
        n = c.species
        if n == 1:
            self.M = np.ctypeslib.as_array(c.M, (c.K, c.K)).copy()
            self.W = np.ctypeslib.as_array(c.W, (c.K, c.K)).copy()
        else:       # M(i), W(i, j) and the per-species values as arrays
            self.M = np.ctypeslib.as_array(c.M, (n, c.K, c.K)).copy()
            self.W = np.ctypeslib.as_array(c.W, (n, n, c.K, c.K)).copy()
            self.MaxNumForM = np.ctypeslib.as_array(c.mM_of, (n,)).copy()
            self.MaxNumForW = np.ctypeslib.as_array(c.mW_of, (n, n)).copy()
            self.SigmaM = np.ctypeslib.as_array(c.sigma_m_of, (n,)).copy()
            self.SigmaW = np.ctypeslib.as_array(c.sigma_w_of, (n, n)).copy()
            self.b = np.ctypeslib.as_array(c.b_of, (n,)).copy()
            self.d = np.ctypeslib.as_array(c.d_of, (n,)).copy()
            self.d_prime = np.ctypeslib.as_array(c.d_prime_of, (n, n)).copy()

    @staticmethod
    def Generate(num, sW, sM, b, d, dd, forced_G=0):
        """Properties::Generate(num, sW, sM, b, d, dd) (Solver.cpp:186): scalars for one species, otherwise sW and dd
        num x num and sM, b, d of length num."""
        if num != 1:
            as_arr = lambda v, shape: np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(shape))
            ptr = C.POINTER(_Problem)()
            _check(library().ldm_problem_generate_multi(int(num), as_arr(sW, (num * num,)), as_arr(sM, (num,)), as_arr(b, (num,)),
                                                        as_arr(d, (num,)), as_arr(dd, (num * num,)), int(forced_G), C.byref(ptr)))
            return Properties(ptr)
        ptr = C.POINTER(_Problem)()
        _check(library().ldm_problem_generate(float(sW), float(sM), float(b), float(d), float(dd), int(forced_G),
                                              C.byref(ptr)))
        return Properties(ptr)

    def CheckKernels(self):
        """Properties::CheckKernels (Solver.cpp:147-183): returns (integral W, integral M), asserts 1 +- 1e-8."""
        iw, im = C.c_double(), C.c_double()
        _check(library().ldm_problem_check_kernels(self._ptr, C.byref(iw), C.byref(im)))
        assert abs(iw.value - 1) < 1e-8 and abs(im.value - 1) < 1e-8
        return iw.value, im.value

    def KernelStructure(self):
        """(separable, max relative deviation) of the m table: whether the prefix-sum path applies (host only)."""
        sep, dev = C.c_int(), C.c_double()
        _check(library().ldm_problem_kernel_structure(self._ptr, C.byref(sep), C.byref(dev)))
        return bool(sep.value), dev.value

    @property
    def state_size(self):
        n = self.SpeciesNumber
        return n + n * n * self.G * self.G

    def __del__(self):
        try:
            if self._ptr:
                library().ldm_problem_free(self._ptr)
                self._ptr = None
        except Exception:
            pass


class Equation:
    """class Equation (Solver.cpp:246-514): constructed from Properties, called as equation(t, vals, type)."""

    def __init__(self, props, device=0, algorithm=ALGO_AUTO):
        self.props = props
        self._h = C.c_void_p()
        _check(library().ldm_create(props._ptr, device, C.byref(self._h)))
        if algorithm != ALGO_AUTO:
            self.set_algorithm(algorithm)

    def set_algorithm(self, algorithm):
        _check(library().ldm_set_algorithm(self._h, algorithm))

    @property
    def algorithm(self):
        return library().ldm_get_algorithm(self._h)

    def GenerateEquilibriumValuesForNAndC(self, value):
        """Solver.cpp:416-428: [N = value | C = value^2 everywhere]"""
        st = np.full(self.props.state_size, float(value) * float(value))
        st[:self.props.SpeciesNumber] = value
        return st

    def set_extrapolation(self, chosen_type):
        """the ExtrapolationType the equation is evaluated with from here on (Solver.cpp:431: chosen_type)"""
        _check(library().ldm_set_extrapolation(self._h, int(chosen_type)))

    def set_closure(self, closure):
        """extension (ldm.h): CLOSURE_POWER1 is the reference; the other two evaluate (w*C)(xi) and K(xi) at every node"""
        _check(library().ldm_set_closure(self._h, int(closure)))

    @property
    def closure(self):
        return library().ldm_get_closure(self._h)

    def closure_integrals(self, vals):
        """(w*C)(xi) and K(xi) = integral of w(xi') C(xi') C(xi + xi') at every node, G x G each"""
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        G = self.props.G
        w_conv, k = np.empty((G, G)), np.empty((G, G))
        _check(library().ldm_closure_integrals(self._h, vals, w_conv.ctypes.data, k.ctypes.data))
        return w_conv, k

    def __call__(self, times, vals, chosen_type=SquaredN):
        self.set_extrapolation(chosen_type)
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        assert vals.size == self.props.state_size          # Solver.cpp:260 / types.h:782
        out = np.empty_like(vals)
        _check(library().ldm_rhs(self._h, vals, out))
        return out

    def integrals(self, vals, want_mc=True):
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        wc = C.c_double()
        mc = np.empty((self.props.G, self.props.G)) if want_mc else None
        _check(library().ldm_integrals(self._h, vals, C.byref(wc), mc.ctypes.data if want_mc else None))
        return wc.value, mc

    def kernels_at_nodes(self):
        G = self.props.G
        m, w = np.empty((G, G)), np.empty((G, G))
        _check(library().ldm_kernels_at_nodes(self._h, m, w))
        return m, w

    # device-resident stepping
    def upload(self, state):
        _check(library().ldm_state_upload(self._h, np.ascontiguousarray(state, dtype=np.float64)))

    def download(self):
        out = np.empty(self.props.state_size)
        _check(library().ldm_state_download(self._h, out))
        return out

    def advance(self, n_steps, dt):
        _check(library().ldm_advance(self._h, int(n_steps), float(dt)))

    def synchronize(self):
        _check(library().ldm_synchronize(self._h))

    def last_elapsed_ms(self):
        ms = C.c_float()
        _check(library().ldm_last_elapsed_ms(self._h, C.byref(ms)))
        return ms.value

    @property
    def launches(self):
        return int(library().ldm_launch_count(self._h))

    def close(self):
        if getattr(self, "_h", None):
            library().ldm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SolverResults:
    """struct SolverResults (Solver.cpp:525-537)"""

    def __init__(self, Times, Values, Derivatives, Observables=None):
        self.Times, self.Values, self.Derivatives, self.Observables = Times, Values, Derivatives, Observables


def count_points(t_max, step):
    return int(library().ldm_count_points(float(t_max), float(step)))


class EulerSolver:
    """class EulerSolver (Solver.cpp:541-593)"""

    def Solve(self, current_equation, type, time_max, step, initial_value, debug_print_enabled=False,
              record=RECORD_ALL, values_out=None, derivs_out=None):
        """values_out / derivs_out: optional caller-owned float64 arrays (e.g. pinned host memory) that receive
        SolverResults::Values / Derivatives instead of freshly allocated ones"""
        eq = current_equation
        eq.set_extrapolation(type)
        init = np.ascontiguousarray(initial_value, dtype=np.float64)
        assert init.size == eq.props.state_size
        n = count_points(time_max, step)
        width = eq.props.state_size
        times = np.empty(n)
        values = derivs = obs = None
        if record == RECORD_LAST:
            values, derivs = np.empty((1, width)), np.empty((1, width))
        elif record == RECORD_OBSERVABLES:
            obs = np.empty((n, 2)) if values_out is None else values_out.reshape(-1)[:n * 2].reshape(n, 2)
        else:
            values = np.empty((n, width)) if values_out is None else values_out.reshape(-1)[:n * width].reshape(n, width)
            derivs = None
            if record == RECORD_ALL:
                derivs = np.empty((n, width)) if derivs_out is None else derivs_out.reshape(-1)[:n * width].reshape(n, width)
        got = C.c_int()
        vals_ptr = (obs if record == RECORD_OBSERVABLES else values).ctypes.data
        _check(library().ldm_solve(eq._h, init, float(step), float(time_max), int(record), times.ctypes.data, vals_ptr,
                                   derivs.ctypes.data if derivs is not None else None, C.byref(got)))
        assert got.value == n
        if debug_print_enabled:
            col = obs[:, 0] if obs is not None else values[:, 0]
            for k in range(len(col)):
                print(f"   t = {times[k]:g};   N = {col[k]:g}")
        return SolverResults(times, values, derivs, obs)


class RungeKutta4thSolver:
    """class RungeKutta4thSolver (Solver.cpp:597-654): in the reference, instantiated by none of its experiments"""

    def Solve(self, current_equation, type, time_max, step, initial_value, debug_print_enabled=False):
        eq = current_equation
        eq.set_extrapolation(type)
        init = np.ascontiguousarray(initial_value, dtype=np.float64)
        assert init.size == eq.props.state_size
        n = int(library().ldm_count_points_rk4(float(time_max), float(step)))
        times, values, derivs = np.empty(n), np.empty((n, init.size)), np.empty((n, init.size))
        got = C.c_int()
        _check(library().ldm_solve_rk4(eq._h, init, float(step), float(time_max), times.ctypes.data, values.ctypes.data,
                                       derivs.ctypes.data, C.byref(got)))
        assert got.value == n
        if debug_print_enabled:
            for k in range(n):
                print(f"   t = {times[k]:g};   N = {values[k, 0]:g}")
        return SolverResults(times, values, derivs)


def solve_batch(equations, inits, step, time_max):
    """The sigma-sweep loop of Experiments::CalculateWithVaryingSigma (Solver.cpp:940-963): last state per entry."""
    n = len(equations)
    handles = (C.c_void_p * n)(*[e._h for e in equations])
    inits = [np.ascontiguousarray(v, dtype=np.float64) for v in inits]
    outs = [np.empty(e.props.state_size) for e in equations]
    in_ptrs = (C.c_void_p * n)(*[v.ctypes.data for v in inits])
    out_ptrs = (C.c_void_p * n)(*[v.ctypes.data for v in outs])
    _check(library().ldm_solve_batch(n, handles, in_ptrs, float(step), float(time_max), out_ptrs))
    return outs


class ShardedEquation:
    """xi-sharded stepping of one large problem: this rank owns rows rank, rank+world, ..."""

    def __init__(self, props, rank, world, device=0):
        self.eq = Equation(props, device)
        self.rank, self.world = rank, world
        _check(library().ldm_shard_configure(self.eq._h, rank, world))
        self.rows_per_rank = int(library().ldm_shard_rows_per_rank(self.eq._h))
        self.G = props.G

    def to_gathered(self, C_natural):
        """natural [G][G] -> gathered [world*rows_per_rank][G] (row x at slot (x%world)*rpr + x//world)"""
        out = np.zeros((self.world * self.rows_per_rank, self.G))
        x = np.arange(self.G)
        out[(x % self.world) * self.rows_per_rank + x // self.world] = np.asarray(C_natural).reshape(self.G, self.G)
        return out

    def from_gathered(self, C_gathered):
        x = np.arange(self.G)
        return np.asarray(C_gathered).reshape(-1, self.G)[(x % self.world) * self.rows_per_rank + x // self.world]

    def set_N(self, N):
        _check(library().ldm_shard_set_N(self.eq._h, float(N)))

    def get_N(self):
        v = C.c_double()
        _check(library().ldm_shard_get_N(self.eq._h, C.byref(v)))
        return v.value

    # ---- fused update + all-gather over NVLink peer stores (no NCCL on the data path) ----
    def peer_connect(self, all_gather_bytes):
        """all_gather_bytes(np.uint8[192]) -> np.uint8[world, 192]: the caller's collective for the IPC handles"""
        mine = np.zeros(3 * 64, dtype=np.uint8)
        _check(library().ldm_shard_peer_alloc(self.eq._h, mine.ctypes.data))
        everyone = np.ascontiguousarray(all_gather_bytes(mine), dtype=np.uint8)
        assert everyone.shape == (self.world, 3 * 64)
        _check(library().ldm_shard_peer_connect(self.eq._h, everyone.ctypes.data))

    def peer_upload(self, C_natural, N):
        _check(library().ldm_shard_peer_upload(self.eq._h, np.ascontiguousarray(self.to_gathered(C_natural)), float(N)))

    def peer_step(self, dt, stream_ptr=0):
        _check(library().ldm_shard_peer_step(self.eq._h, float(dt), C.c_void_p(stream_ptr)))

    def peer_advance(self, n_steps, dt, stream_ptr=0):
        """n Euler steps with the exchange fused in; on the prefix-sum path ONE launch per rank for all of them"""
        _check(library().ldm_shard_peer_advance(self.eq._h, int(n_steps), float(dt), C.c_void_p(stream_ptr)))

    @staticmethod
    def connect_local(ranks):
        """one process driving all ranks (several devices, or several ranks on one device in tests): replaces
        peer_connect and the exchange of IPC handles"""
        handles = (C.c_void_p * len(ranks))(*[r.eq._h for r in ranks])
        _check(library().ldm_shard_peer_connect_local(len(ranks), handles))

    def peer_download(self):
        out = np.empty((self.world * self.rows_per_rank, self.G))
        n = C.c_double()
        _check(library().ldm_shard_peer_download(self.eq._h, out, C.byref(n)))
        return self.from_gathered(out), n.value

    def step(self, d_C_gathered_ptr, d_C_shard_out_ptr, dt, stream_ptr):
        _check(library().ldm_shard_step(self.eq._h, C.c_void_p(d_C_gathered_ptr), C.c_void_p(d_C_shard_out_ptr),
                                        float(dt), C.c_void_p(stream_ptr)))
