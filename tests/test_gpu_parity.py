"""Parity of the CUDA path (through the C ABI) with the CPU oracle and the golden fixtures.

Tolerances (BASELINE.json north_star: 1e-10 relative on N(t) and C(xi,t)):
  * trajectories: |gpu - ref| <= 1e-10 * |ref| at every stored point, every component; entries the reference
    clamps to exactly 0 (Solver.cpp:578-580) must be exactly 0 on the GPU as well;
  * a single right-hand side: far from the origin dC is the cancellation residue of terms of size ~2bN^2, so
    the bound is 1e-12 times the sum of the term magnitudes (SURVEY.md 8c comparison rule).
"""
import glob
import os

import numpy as np
import pytest

from oracle import OracleProblem

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(autouse=True)
def default_shapes(ldm):
    """shape overrides (ldm_tuning_set) never outlive the test that set them"""
    yield
    ldm.tuning_set("reset")

RTOL_TRAJECTORY = 1e-10
RTOL_RHS_TERMS = 1e-12


def golden(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


def problems_from(g, name, ldm):
    forced = int(g["G"]) if "synth" in name else 0
    args = (float(g["sigmaW"]), float(g["sigmaM"]), float(g["b"]), float(g["d"]), float(g["dd"]))
    return ldm.Properties.Generate(1, *args, forced), OracleProblem(*args, forced)


def rhs_term_scale(o, state, integrals=None, nodes=None):
    """sum of |terms| of dC (Solver.cpp:464-505) and of dN (446-451), per component"""
    N = state[0]
    C = state[1:].reshape(o.G, o.G)
    wc, mc = integrals if integrals is not None else o.integrals(state)
    m_nodes, w_nodes = nodes if nodes is not None else o.kernels_at_nodes()
    scale_c = (2 * o.b * np.abs(mc) + 2 * o.b * N * np.abs(m_nodes) + 2 * o.d * np.abs(C)
               + 2 * o.dd * np.abs(wc * C / N) + 2 * o.dd * np.abs(w_nodes * C))
    scale_n = abs(N * (o.b - o.d)) + abs(o.dd * wc)
    return np.concatenate([[scale_n], scale_c.ravel()])


def assert_rhs_close(o, state, got, want):
    scale = rhs_term_scale(o, state)
    err = np.abs(got - want)
    worst = np.argmax(err / scale)
    assert np.all(err <= RTOL_RHS_TERMS * scale), (worst, got[worst], want[worst], err[worst] / scale[worst])


def assert_states_close(got, want):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape
    assert np.array_equal(got == 0, want == 0), "clamped entries differ"
    err = np.abs(got - want)
    bound = RTOL_TRAJECTORY * np.abs(want)
    worst = np.unravel_index(np.argmax(err - bound), err.shape)
    assert np.all(err <= bound), (worst, got[worst], want[worst])


def algo_id(ldm, algorithm):
    return {"direct": ldm.ALGO_DIRECT, "factored": ldm.ALGO_FACTORED, "prefix": ldm.ALGO_PREFIX}[algorithm]


ALL = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(HERE, "golden", "*.npz")) if not os.path.basename(p).startswith(("large_", "extras_", "multi_", "exp_14_6_sweep")))


@pytest.mark.parametrize("name", ALL)
def test_kernels_at_nodes_bitwise(ldm, name):
    g = golden(name)
    props, _ = problems_from(g, name, ldm)
    eq = ldm.Equation(props)
    m, w = eq.kernels_at_nodes()
    assert np.array_equal(m, g["m_nodes"]) and np.array_equal(w, g["w_nodes"])


@pytest.mark.parametrize("algorithm", ["direct", "factored", "prefix"])
@pytest.mark.parametrize("name", ALL)
def test_rhs_against_golden(ldm, name, algorithm):
    g = golden(name)
    props, o = problems_from(g, name, ldm)
    eq = ldm.Equation(props, algorithm=algo_id(ldm, algorithm))
    if "synth" in name:
        cases = [(g["state"], g["deriv"])]
    else:
        cases = [(eq.GenerateEquilibriumValuesForNAndC(200), g["deriv_k0"]), (g["state_k100"], g["deriv_k100"])]
    for state, want in cases:
        got = eq(0.0, state, ldm.SquaredN)
        assert_rhs_close(o, state, got, want)


@pytest.mark.parametrize("algorithm", ["direct", "factored", "prefix"])
@pytest.mark.parametrize("name", ["exp_14_4_b", "exp_14_4_c", "exp_14_6_a_080", "exp_14_6_b_000", "synth_g32", "synth_g64",
                                  "synth_g48_narrow_w"])
def test_integrals_against_oracle(ldm, name, algorithm):
    g = golden(name)
    props, o = problems_from(g, name, ldm)
    eq = ldm.Equation(props, algorithm=algo_id(ldm, algorithm))
    state = g["state"] if "synth" in name else g["state_k100"]
    wc, mc = eq.integrals(state)
    wc_ref, mc_ref = (float(g["wc"]), g["mc"]) if "synth" in name else o.integrals(state)
    assert abs(wc - wc_ref) <= 1e-13 * abs(wc_ref)
    assert np.all(np.abs(mc - mc_ref) <= 1e-13 * np.abs(mc_ref))


@pytest.mark.parametrize("algorithm", ["prefix", "factored"])
@pytest.mark.parametrize("name", ["exp_14_4_a", "exp_14_4_b", "exp_14_4_c"])
def test_experiment_14_4_trajectory(ldm, name, algorithm):
    """CalculateFirstMoment(100) (Solver.cpp:885-906): dt 0.1, 1001 points, N0 = 200."""
    g = golden(name)
    props, _ = problems_from(g, name, ldm)
    eq = ldm.Equation(props, algorithm=algo_id(ldm, algorithm))
    res = ldm.EulerSolver().Solve(eq, ldm.SquaredN, 100, 0.1, eq.GenerateEquilibriumValuesForNAndC(200),
                                  record=ldm.RECORD_ALL)
    assert len(res.Times) == 1001 and np.array_equal(res.Times, g["times"])
    assert_states_close(res.Values[:, [0, props.G]], g["obs"])
    for k, key in ((1, "state_k1"), (10, "state_k10"), (100, "state_k100"), (500, "state_k500"), (1000, "final_state")):
        assert_states_close(res.Values[k], g[key])
    o = OracleProblem(float(g["sigmaW"]), float(g["sigmaM"]))
    assert_rhs_close(o, g["final_state"], res.Derivatives[1000], g["final_deriv"])
    # the other record modes see the same run
    obs = ldm.EulerSolver().Solve(eq, ldm.SquaredN, 100, 0.1, eq.GenerateEquilibriumValuesForNAndC(200),
                                  record=ldm.RECORD_OBSERVABLES)
    assert np.array_equal(obs.Observables, res.Values[:, [0, props.G]])
    last = ldm.EulerSolver().Solve(eq, ldm.SquaredN, 100, 0.1, eq.GenerateEquilibriumValuesForNAndC(200),
                                   record=ldm.RECORD_LAST)
    assert np.array_equal(last.Values[0], res.Values[-1])
    assert np.array_equal(last.Derivatives[0], res.Derivatives[-1])


@pytest.mark.parametrize("algorithm", ["prefix", "factored"])
@pytest.mark.parametrize("name", [n for n in ALL if "14_6" in n])
def test_experiment_14_6_entries(ldm, name, algorithm):
    """One entry of CalculateWithVaryingSigma(200) (Solver.cpp:909-977): 2001 points; includes the entries where
    the clamp is active (14.6a[0], sigma_W = 1e-8) and where the population goes extinct (14.6b[0])."""
    g = golden(name)
    props, _ = problems_from(g, name, ldm)
    eq = ldm.Equation(props, algorithm=algo_id(ldm, algorithm))
    res = ldm.EulerSolver().Solve(eq, ldm.SquaredN, 200, 0.1, eq.GenerateEquilibriumValuesForNAndC(200),
                                  record=ldm.RECORD_VALUES)
    assert len(res.Times) == 2001 and np.array_equal(res.Times, g["times"])
    assert_states_close(res.Values[:, [0, props.G]], g["obs"])
    for k, key in ((1, "state_k1"), (2, "state_k2"), (100, "state_k100"), (2000, "final_state")):
        assert_states_close(res.Values[k], g[key])


def test_resident_stepping_equals_solve(ldm):
    g = golden("exp_14_4_b")
    props, _ = problems_from(g, "exp_14_4_b", ldm)
    eq = ldm.Equation(props)
    eq.upload(eq.GenerateEquilibriumValuesForNAndC(200))
    eq.advance(100, 0.1)
    eq.synchronize()
    assert_states_close(eq.download(), g["state_k100"])
    assert eq.last_elapsed_ms() > 0 and eq.launches > 0


@pytest.mark.parametrize("algorithm", ["prefix", "factored"])
def test_batch_of_parameter_sets(ldm, algorithm):
    """the 14.6 loop (Solver.cpp:940-963) over a few entries at once, slot-indexed results"""
    names = ["exp_14_6_a_012", "exp_14_6_b_040", "exp_14_6_a_080", "exp_14_6_b_012"]
    eqs, inits, wants = [], [], []
    for name in names:
        g = golden(name)
        props, _ = problems_from(g, name, ldm)
        eq = ldm.Equation(props, algorithm=algo_id(ldm, algorithm))
        eqs.append(eq)
        inits.append(eq.GenerateEquilibriumValuesForNAndC(200))
        wants.append(g["final_state"])
    outs = ldm.solve_batch(eqs, inits, 0.1, 200)
    for got, want in zip(outs, wants):
        assert_states_close(got, want)


@pytest.mark.parametrize("capacity", [0, 7])
def test_persistent_batch_equals_per_step_batch(ldm, capacity):
    """the batch of independent parameter sets with ALL steps in one launch (persistent_batch_kernel: a team of CTAs per
    set, per-set barriers) against one launch per kernel and step: the same bits for every set -- teams of many CTAs
    (few sets on a whole GPU), teams of one, a set whose rows outnumber... and, with the CTA budget capped at 7, several
    waves of sets that queue for the device"""
    names = ["exp_14_6_a_012", "exp_14_6_b_040", "exp_14_6_a_080", "exp_14_6_b_012", "exp_14_6_b_000", "exp_14_6_a_159", "exp_14_4_b",
             "exp_14_6_a_000", "synth_g64", "exp_14_4_c", "exp_14_6_b_159"]
    runs = []
    for persistent in (1, 0):
        ldm.tuning_set("reset")
        ldm.tuning_set("persistent", persistent)
        if capacity:
            ldm.tuning_set("batch_capacity", capacity)
        eqs, inits = [], []
        for name in names:
            g = golden(name)
            props, _ = problems_from(g, name, ldm)
            eq = ldm.Equation(props)
            eqs.append(eq)
            inits.append(g["state_k100"] if "state_k100" in g.files else g["state"])
        launches0 = sum(e.launches for e in eqs)
        runs.append(ldm.solve_batch(eqs, inits, 0.1, 3.75))           # 37 steps: an odd count, the ping-pong parity must come out right
        launches = sum(e.launches for e in eqs) - launches0
        assert (launches <= 4) if persistent else (launches >= 37 * 2)
        again = ldm.solve_batch(eqs, inits, 0.1, 3.75)                 # the handles are reusable: counters and parity start over
        assert all(np.array_equal(a, b) for a, b in zip(runs[-1], again))
    for name, one, two in zip(names, *runs):
        assert np.array_equal(one, two), name


@pytest.mark.parametrize("algorithm", ["prefix", "factored", "direct"])
@pytest.mark.parametrize("name", ["b", "c"])
def test_runge_kutta_and_extrapolation_types(ldm, name, algorithm):
    """SURVEY 8f rank 3 on the same right-hand-side kernels: RungeKutta4thSolver (Solver.cpp:597-654) and the Zero / Alpha
    Exponential extrapolation types of C (Solver.cpp:361-408), against fixtures computed by the reference itself"""
    path = os.path.join(HERE, "golden", f"extras_14_4_{name}.npz")
    if not os.path.exists(path):
        pytest.skip("fixture not generated")
    g = np.load(path)
    base = golden(f"exp_14_4_{name}")
    props, o = problems_from(base, f"exp_14_4_{name}", ldm)
    eq = ldm.Equation(props, algorithm=algo_id(ldm, algorithm))
    mid = base["state_k100"]
    res = ldm.RungeKutta4thSolver().Solve(eq, ldm.SquaredN, 4.0, 0.1, eq.GenerateEquilibriumValuesForNAndC(200))
    assert np.array_equal(res.Times, g["rk4_times"])
    assert_states_close(res.Values, g["rk4_values"])
    assert_rhs_close(o, g["rk4_values"][-1], res.Derivatives[-1], g["rk4_derivs"][-1])
    for kind, label in ((ldm.Zero, "zero"), (ldm.Alpha, "alpha"), (ldm.Exponential, "exponential")):
        if f"{label}_deriv_k0" not in g.files:
            continue
        for state, key in ((eq.GenerateEquilibriumValuesForNAndC(200), "deriv_k0"), (mid, "deriv_k100")):
            assert_rhs_close(o, state, eq(0.0, state, kind), g[f"{label}_{key}"])
        run = ldm.EulerSolver().Solve(eq, kind, 4.0, 0.1, mid, record=ldm.RECORD_VALUES)
        assert_states_close(run.Values, g[f"{label}_values"])
    res = ldm.RungeKutta4thSolver().Solve(eq, ldm.Zero, 2.0, 0.1, mid)
    assert_states_close(res.Values, g["zero_rk4_values"])
    # and back: the default type still reproduces the experiment's numbers
    run = ldm.EulerSolver().Solve(eq, ldm.SquaredN, 10.0, 0.1, eq.GenerateEquilibriumValuesForNAndC(200), record=ldm.RECORD_VALUES)
    assert_states_close(run.Values[100], base["state_k100"])


@pytest.mark.parametrize("algorithm", ["prefix", "factored"])
@pytest.mark.parametrize("name", ["2_species", "3_species"])
def test_several_species(ldm, name, algorithm):
    """Equation and EulerSolver with SpeciesNumber > 1 (Solver.cpp:265-295, 446-511) on the reference's
    own multi-species example inputs, against fixtures computed by the reference itself: one right-hand side at a state
    with distinct densities and asymmetric pair densities, the uniform state, Zero extrapolation, and 30 Euler steps"""
    g = golden(f"multi_{name}")
    s = int(g["species"])
    sw = {2: [[0.05, 0.06], [0.06, 0.05]], 3: [[0.05, 0.06, 0.07], [0.06, 0.05, 0.04], [0.07, 0.04, 0.05]]}[s]
    props = ldm.Properties.Generate(s, sw, [0.05] * s, [0.4, 0.401, 0.4001][:s], [0.2] * s, np.full((s, s), 0.001))
    eq = ldm.Equation(props, algorithm=algo_id(ldm, algorithm))
    P = props.G ** 2

    def assert_deriv_close(got, want):
        # blockwise: the far field of every dC_ij is a cancellation residue of terms as large as the block's largest entry
        for lo, hi in [(0, s)] + [(s + k * P, s + (k + 1) * P) for k in range(s * s)]:
            scale = np.max(np.abs(want[lo:hi])) + 1.0
            assert np.max(np.abs(got[lo:hi] - want[lo:hi])) <= 1e-11 * scale, (lo, np.max(np.abs(got[lo:hi] - want[lo:hi])) / scale)

    assert_deriv_close(eq(0.0, g["state"], ldm.SquaredN), g["deriv"])
    assert_deriv_close(eq(0.0, eq.GenerateEquilibriumValuesForNAndC(200), ldm.SquaredN), g["deriv_equilibrium"])
    assert_deriv_close(eq(0.0, g["state"], ldm.Zero), g["zero_deriv"])
    assert not np.array_equal(g["zero_deriv"], g["deriv"])
    res = ldm.EulerSolver().Solve(eq, ldm.SquaredN, 3.05, 0.1, g["state"], record=ldm.RECORD_ALL)
    assert np.array_equal(res.Times, g["times"])
    assert_states_close(res.Values, g["values"])
    assert_deriv_close(res.Derivatives[-1], g["derivs"][-1])
    last = ldm.EulerSolver().Solve(eq, ldm.SquaredN, 3.05, 0.1, g["state"], record=ldm.RECORD_LAST)
    assert np.array_equal(last.Values[0], res.Values[-1])
    eq.upload(g["state"])
    eq.advance(30, 0.1)
    assert np.array_equal(eq.download(), res.Values[-1])
    rk = ldm.RungeKutta4thSolver().Solve(eq, ldm.SquaredN, 1.0, 0.1, g["state"])       # S:597-654 on the same equations
    assert np.array_equal(rk.Times, g["rk4_times"])
    assert_states_close(rk.Values, g["rk4_values"])
    assert_deriv_close(rk.Derivatives[-1], g["rk4_last_deriv"])
    with pytest.raises(ldm.LdmError):
        eq.integrals(g["state"])                  # one-species entry points say so instead of misreading the state


def test_several_species_replayed_steps_and_batch(ldm):
    """the Euler loop of a multi-species handle replayed as a CUDA graph of two steps (ldm_advance, ldm_solve with record =
    last) and such handles as entries of a batch beside one-species entries: bit for bit what one launch sequence per step
    gives (record = all takes that route), for even and odd step counts and from either parity of the state ping-pong"""
    g = golden("multi_2_species")
    props = ldm.Properties.Generate(2, [[0.05, 0.06], [0.06, 0.05]], [0.05] * 2, [0.4, 0.401], [0.2] * 2, np.full((2, 2), 0.001))
    eq = ldm.Equation(props)
    init = g["state"]
    points = 41
    ref = ldm.EulerSolver().Solve(eq, ldm.SquaredN, (points - 0.5) * 0.1, 0.1, init, record=ldm.RECORD_ALL)
    before = eq.launches
    last = ldm.EulerSolver().Solve(eq, ldm.SquaredN, (points - 0.5) * 0.1, 0.1, init, record=ldm.RECORD_LAST)
    assert np.array_equal(last.Values[0], ref.Values[-1]) and np.array_equal(last.Derivatives[0], ref.Derivatives[-1])
    assert eq.launches - before >= 3 * (points - 1)            # the replayed launches are counted
    eq.upload(init)
    eq.advance(13, 0.1)                                        # odd: ends at parity 1 ...
    assert np.array_equal(eq.download(), ref.Values[13])
    eq.advance(20, 0.1)                                        # ... and goes on from there
    eq.advance(7, 0.1)
    assert np.array_equal(eq.download(), ref.Values[40])
    eq.upload(init)
    eq.advance(12, 0.05)                                       # another step size: another graph
    other = ldm.EulerSolver().Solve(eq, ldm.SquaredN, 12.5 * 0.05, 0.05, init, record=ldm.RECORD_VALUES)
    eq.upload(init); eq.advance(12, 0.05)
    assert np.array_equal(eq.download(), other.Values[12])
    # a batch of two multi-species entries and three one-species entries
    single_props = [ldm.Properties.Generate(1, sw, 0.08, 0.4, 0.2, 0.001) for sw in (0.05, 0.1, 0.15)]
    singles = [ldm.Equation(p) for p in single_props]
    single_inits = [e.GenerateEquilibriumValuesForNAndC(200) for e in singles]
    eq2 = ldm.Equation(props)
    init2 = init.copy(); init2[:2] *= 1.1
    want2 = ldm.EulerSolver().Solve(eq2, ldm.SquaredN, (points - 0.5) * 0.1, 0.1, init2, record=ldm.RECORD_VALUES).Values[-1]
    want_single = [ldm.EulerSolver().Solve(e, ldm.SquaredN, (points - 0.5) * 0.1, 0.1, v, record=ldm.RECORD_LAST).Values[0]
                   for e, v in zip(singles, single_inits)]
    outs = ldm.solve_batch([singles[0], eq, singles[1], eq2, singles[2]], [single_inits[0], init, single_inits[1], init2, single_inits[2]],
                           0.1, (points - 0.5) * 0.1)
    assert np.array_equal(outs[1], ref.Values[-1]) and np.array_equal(outs[3], want2)
    for got, want in zip((outs[0], outs[2], outs[4]), want_single):
        assert np.array_equal(got, want)


@pytest.mark.parametrize("which", ["a", "b"])
def test_whole_sweep_against_the_reference(ldm, which):
    """Experiment 14.6 a / b (S:1033-1071) entry by entry: ALL 160 entries of each sweep, 2000 Euler steps each, advanced
    as one batch (ldm_solve_batch) -- what the experiment plots per entry (N and C[0][G-1] of the last state, S:1050-1055)
    and C[0][0] within 1e-10 of the reference's own numbers (tests/golden/exp_14_6_sweep.npz, by-reference build), clamped
    zeros exactly zero"""
    g = golden("exp_14_6_sweep")
    sigmas = g[f"{which}_sigma"]
    eqs = []
    for i, sig in enumerate(sigmas):
        sw, sm = (float(sig), 0.08) if which == "a" else (0.05, float(sig))       # Data/14/6/a.csv: sigma_M = 0.08; b.csv: sigma_W = 0.05
        props = ldm.Properties.Generate(1, sw, sm, 0.4, 0.2, 0.001)
        assert (props.G, props.K, props.MaxNumForM, props.MaxNumForW) == (g[f"{which}_G"][i], g[f"{which}_K"][i], g[f"{which}_mM"][i], g[f"{which}_mW"][i])
        eqs.append(ldm.Equation(props))
    inits = [eq.GenerateEquilibriumValuesForNAndC(200) for eq in eqs]
    outs = ldm.solve_batch(eqs, inits, 0.1, 200)
    got = np.array([[o[0], o[e.props.G], o[1]] for o, e in zip(outs, eqs)])
    want = np.stack([g[f"{which}_N_end"], g[f"{which}_C_end"], g[f"{which}_C00_end"]], 1)
    assert_states_close(got, want)


def test_batch_spread_over_two_devices(ldm):
    """the sweep with its entries living on two GPUs (ldm_create's device argument): every device advances its slice
    concurrently; same bits as the same entries on one device"""
    if ldm.library().ldm_device_count() < 2:
        pytest.skip("needs two GPUs")
    names = ["exp_14_6_a_012", "exp_14_6_b_040", "exp_14_6_a_080", "exp_14_6_b_012", "exp_14_6_a_159"]
    runs = []
    for devices in ([0] * len(names), [0, 1, 1, 0, 1]):
        eqs, inits = [], []
        for name, device in zip(names, devices):
            g = golden(name)
            props, _ = problems_from(g, name, ldm)
            eq = ldm.Equation(props, device=device)
            eqs.append(eq)
            inits.append(eq.GenerateEquilibriumValuesForNAndC(200))
        runs.append(ldm.solve_batch(eqs, inits, 0.1, 200))
    for name, one, two in zip(names, *runs):
        assert np.array_equal(one, two)
        assert_states_close(two, golden(name)["final_state"])


@pytest.mark.parametrize("world", [2, 3, 5, 7])
def test_sharded_step_emulated_on_one_device(ldm, world):
    """xi-sharding with all ranks emulated on one GPU: each rank updates its rows of C from the
    gathered replica; the 'all-gather' is a concatenation.  Must equal the unsharded run bit for bit."""
    import torch
    g = golden("exp_14_4_c")
    props, _ = problems_from(g, "exp_14_4_c", ldm)
    G = props.G
    ranks = [ldm.ShardedEquation(props, r, world) for r in range(world)]
    rpr = ranks[0].rows_per_rank
    state = g["state_k100"]
    bufs = [torch.from_numpy(ranks[0].to_gathered(state[1:])).cuda(), torch.zeros(world * rpr, G, dtype=torch.float64, device="cuda")]
    for r in ranks:
        r.set_N(state[0])
    stream = torch.cuda.current_stream().cuda_stream
    steps = 5
    for k in range(steps):
        src, dst = bufs[k % 2], bufs[(k + 1) % 2]
        for r in ranks:
            r.step(src.data_ptr(), dst[r.rank * rpr:(r.rank + 1) * rpr].data_ptr(), 0.1, stream)
    torch.cuda.synchronize()
    single = ldm.Equation(props)
    single.upload(state)
    single.advance(steps, 0.1)
    want = single.download()
    got_C = ranks[0].from_gathered(bufs[steps % 2].cpu().numpy())
    assert np.array_equal(got_C.ravel(), want[1:])
    for r in ranks:
        assert r.get_N() == want[0]


@pytest.mark.parametrize("how", ["one step per launch", "all steps in one launch", "dense contraction"])
@pytest.mark.parametrize("G", [0, 200])
def test_peer_exchange_step_on_one_rank(ldm, how, G):
    """the sharded step with the exchange fused in (ldm_shard_peer_*), world = 1 so that one GPU can run it: the rank
    stores its rows into its own next replica and waits on its own flag.  Prefix-sum path: the persistent work-list
    kernel, one launch per step or one launch for all steps; dense contraction: three launches per step with the stores
    in the closing kernel.  All must equal the unsharded run bit for bit."""
    import torch
    if G:
        props = ldm.Properties.Generate(1, 0.12, 0.09, 0.4, 0.2, 0.001, G)       # two column chunks
        state = ldm.Equation(props).GenerateEquilibriumValuesForNAndC(200)
        g = props.MomentsGrid
        state[1:] *= (1 + 0.5 * np.exp(-(g[:, None] ** 2 + g[None, :] ** 2) / (2 * 0.1 ** 2))).ravel()
    else:
        base = golden("exp_14_4_c")
        props, _ = problems_from(base, "exp_14_4_c", ldm)
        state = base["state_k100"]
    sh = ldm.ShardedEquation(props, 0, 1)
    if how == "dense contraction":
        sh.eq.set_algorithm(ldm.ALGO_FACTORED)
    sh.peer_connect(lambda mine: mine.reshape(1, -1))
    sh.peer_upload(state[1:], state[0])
    steps = 7
    stream = torch.cuda.current_stream().cuda_stream
    launches0 = sh.eq.launches
    if how == "all steps in one launch":
        sh.peer_advance(4, 0.1, stream)
        sh.peer_advance(3, 0.1, stream)            # odd + even: the replica parity must come out right
        assert sh.eq.launches - launches0 == 2
    else:
        for _ in range(steps):
            sh.peer_step(0.1, stream)
    got_C, got_N = sh.peer_download()
    single = ldm.Equation(props, algorithm=sh.eq.algorithm)
    single.upload(state)
    single.advance(steps, 0.1)
    want = single.download()
    assert np.array_equal(got_C.ravel(), want[1:]) and got_N == want[0]


@pytest.mark.parametrize("world,G,order", [(2, 200, 0), (3, 200, 0), (2, 0, 0), (4, 333, 0), (2, 200, 1), (3, 200, 2)])
def test_ranks_on_one_device_elementwise(ldm, world, G, order):
    """SURVEY 8e at world > 1 on ONE GPU: every rank is a handle of its own on device 0 with its own stream and its own
    launch of the persistent work-list kernel (grids capped so that all of them are resident together); rows travel
    between the ranks' replicas through the same peer stores and epoch flags a multi-GPU run uses (wired up by
    ldm_shard_peer_connect_local instead of CUDA IPC).  Every rank's replica must equal the unsharded run ELEMENT BY
    ELEMENT -- a permutation bug in the gathered layout would keep a checksum, not this.  The steps of a launch are ordered
    by data dependencies alone (per-row arrival flags; items that start before their state is complete run behind row
    gates), so the hand-out order of the items must not matter either: rows ascending (0, the default), chunk by chunk (1),
    rows descending (2)."""
    import torch
    ldm.tuning_set("item_order", order)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    if G:
        props = ldm.Properties.Generate(1, 0.12, 0.09, 0.4, 0.2, 0.001, G)       # several column chunks, uneven row counts
        state = ldm.Equation(props).GenerateEquilibriumValuesForNAndC(200)
        g = props.MomentsGrid
        state[1:] *= (1 + 0.5 * np.exp(-(g[:, None] ** 2 + g[None, :] ** 2) / (2 * 0.1 ** 2)) + 0.1 * np.sin(7 * g[:, None] + 3 * g[None, :])).ravel()
    else:
        base = golden("exp_14_4_c")                                             # one chunk per row, 256-thread CTAs
        props, _ = problems_from(base, "exp_14_4_c", ldm)
        state = base["state_k100"]
    ldm.tuning_set("chunk_grid", max(1, (sms - 8) // world))
    ranks = [ldm.ShardedEquation(props, r, world) for r in range(world)]
    ldm.ShardedEquation.connect_local(ranks)
    for r in ranks:
        r.peer_upload(state[1:], state[0])
    streams = [torch.cuda.Stream() for _ in ranks]
    steps = 9
    for n in (4, 5):                                # two launches per rank; all ranks of a launch run side by side
        for r, st in zip(ranks, streams):
            r.peer_advance(n, 0.1, st.cuda_stream)
    torch.cuda.synchronize()
    single = ldm.Equation(props)
    single.upload(state)
    single.advance(steps, 0.1)
    want = single.download()
    for r in ranks:
        got_C, got_N = r.peer_download()
        assert np.array_equal(got_C.ravel(), want[1:]), f"rank {r.rank}: max |diff| {np.max(np.abs(got_C.ravel() - want[1:]))}"
        assert got_N == want[0]


def test_two_processes_two_gpus_elementwise(ldm):
    """the same check across two processes and two GPUs (CUDA IPC mappings, NVLink peer stores): tests/mp_shard_check.py
    under torchrun; every rank compares its replica element-wise with its own one-rank run"""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29731", os.path.join(HERE, "mp_shard_check.py"), "--G", "256", "--steps", "12"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("max_abs_diff 0.0") == 2, out.stdout[-2000:]


def test_lost_rank_is_reported_not_hung(ldm):
    """a peer group of two in which only rank 0 ever steps: its kernel gives up after the bounded wait (~9 s) and the
    next download reports the error instead of handing back an incomplete replica"""
    import torch
    base = golden("exp_14_4_b")
    props, _ = problems_from(base, "exp_14_4_b", ldm)
    ranks = [ldm.ShardedEquation(props, r, 2) for r in range(2)]
    ldm.ShardedEquation.connect_local(ranks)
    for r in ranks:
        r.peer_upload(base["state_k100"][1:], base["state_k100"][0])
    ranks[0].peer_advance(3, 0.1, torch.cuda.current_stream().cuda_stream)
    with pytest.raises(ldm.LdmError, match="never published|timed out"):
        ranks[0].peer_download()


def test_natural_layout_entry_points_refuse_a_sharded_handle(ldm):
    """after ldm_shard_configure(world > 1) the kernels index C through the gathered layout and compute this rank's rows
    only: the natural-layout entry points must say so instead of returning partial results"""
    base = golden("exp_14_4_b")
    props, _ = problems_from(base, "exp_14_4_b", ldm)
    sh = ldm.ShardedEquation(props, 1, 2)
    eq, state = sh.eq, base["state_k100"]
    for call in (lambda: eq(0.0, state), lambda: eq.integrals(state), lambda: eq.upload(state), lambda: eq.advance(1, 0.1),
                 lambda: eq.download(), lambda: ldm.EulerSolver().Solve(eq, ldm.SquaredN, 1.0, 0.1, state, record=ldm.RECORD_LAST),
                 lambda: ldm.RungeKutta4thSolver().Solve(eq, ldm.SquaredN, 1.0, 0.1, state), lambda: ldm.solve_batch([eq], [state], 0.1, 1.0)):
        with pytest.raises(ldm.LdmError):
            call()
    ldm.library().ldm_shard_configure(eq._h, 0, 1)          # the way back
    single = ldm.Equation(props)
    assert np.array_equal(eq(0.0, state), single(0.0, state))


def test_degenerate_solves(ldm):
    """a solve that stores the initial point only (time_max below one step, S:574-575), in every record mode and for the
    Runge-Kutta solver; zero steps of resident stepping; the derivative asked for at the only point"""
    g = golden("exp_14_4_b")
    props, o = problems_from(g, "exp_14_4_b", ldm)
    eq = ldm.Equation(props)
    init = g["state_k100"]
    for record in (ldm.RECORD_LAST, ldm.RECORD_OBSERVABLES, ldm.RECORD_VALUES, ldm.RECORD_ALL):
        res = ldm.EulerSolver().Solve(eq, ldm.SquaredN, 0.05, 0.1, init, record=record)
        assert len(res.Times) == 1 and res.Times[0] == 0.0
        if res.Values is not None:
            assert np.array_equal(res.Values[0], init)
        if res.Observables is not None:
            assert np.array_equal(res.Observables[0], init[[0, props.G]])
        if res.Derivatives is not None:
            assert_rhs_close(o, init, res.Derivatives[0], g["deriv_k100"])
    rk = ldm.RungeKutta4thSolver().Solve(eq, ldm.SquaredN, 0.0, 0.1, init)
    assert len(rk.Times) == 1 and np.array_equal(rk.Values[0], init)
    assert_rhs_close(o, init, rk.Derivatives[0], g["deriv_k100"])
    eq.upload(init)
    eq.advance(0, 0.1)
    assert np.array_equal(eq.download(), init)


def test_direct_path_trajectory(ldm):
    """the sample-by-sample algorithm through a whole experiment (the default is the factored one)"""
    g = golden("exp_14_4_b")
    props, _ = problems_from(g, "exp_14_4_b", ldm)
    eq = ldm.Equation(props, algorithm=ldm.ALGO_DIRECT)
    res = ldm.EulerSolver().Solve(eq, ldm.SquaredN, 100, 0.1, eq.GenerateEquilibriumValuesForNAndC(200),
                                  record=ldm.RECORD_VALUES)
    assert_states_close(res.Values[:, [0, props.G]], g["obs"])
    assert_states_close(res.Values[1000], g["final_state"])


@pytest.mark.parametrize("algorithm", ["prefix", "factored"])
def test_factored_equals_direct_on_the_16k_grid(ldm, algorithm):
    """BASELINE config 3 (G = 128, 16 384 xi-points): the CPU oracle needs ~10 s per evaluation there, so the
    GPU algorithms -- each pinned to the oracle at the smaller sizes above -- are compared with each other."""
    props = ldm.Properties.Generate(1, 0.15, 0.15, 0.4, 0.2, 0.001, 128)
    direct = ldm.Equation(props, algorithm=ldm.ALGO_DIRECT)
    factored = ldm.Equation(props, algorithm=algo_id(ldm, algorithm))
    g = props.MomentsGrid
    st = direct.GenerateEquilibriumValuesForNAndC(200)
    st[1:] *= (1 + 0.5 * np.exp(-(g[:, None] ** 2 + g[None, :] ** 2) / (2 * 0.15 ** 2))).ravel()
    (w1, m1), (w2, m2) = direct.integrals(st), factored.integrals(st)
    assert abs(w1 - w2) <= 1e-14 * abs(w1)
    assert np.all(np.abs(m1 - m2) <= 1e-13 * np.abs(m1))
    for eq in (direct, factored):
        eq.upload(st)
        eq.advance(25, 0.1)
    assert_states_close(factored.download(), direct.download())


@pytest.mark.parametrize("G,bm,bn", [(128, 40, 64), (200, 80, 128), (200, 40, 64), (96, 80, 64), (200, 128, 128), (100, 128, 128)])
def test_factored_tilings_equal_direct(ldm, G, bm, bn):
    """every tiling of the factored kernel (several beta tiles, several column chunks with their halo column, ragged
    edges) against the sample-by-sample path on grids the CPU oracle cannot reach in a test"""
    props = ldm.Properties.Generate(1, 0.12, 0.09, 0.4, 0.2, 0.001, G)
    direct = ldm.Equation(props, algorithm=ldm.ALGO_DIRECT)
    ldm.tuning_set("factored_bm", bm)
    ldm.tuning_set("factored_bn", bn)
    factored = ldm.Equation(props, algorithm=ldm.ALGO_FACTORED)
    g = props.MomentsGrid
    st = direct.GenerateEquilibriumValuesForNAndC(200)
    st[1:] *= (1 + 0.5 * np.exp(-(g[:, None] ** 2 + g[None, :] ** 2) / (2 * 0.1 ** 2)) + 0.1 * np.sin(7 * g[:, None] + 3 * g[None, :])).ravel()
    (w1, m1), (w2, m2) = direct.integrals(st), factored.integrals(st)
    assert abs(w1 - w2) <= 1e-14 * abs(w1)
    assert np.all(np.abs(m1 - m2) <= 1e-13 * np.abs(m1))


@pytest.mark.parametrize("G,bn", [(128, 0), (128, 64), (128, 32), (200, 0), (200, 96), (200, 32), (96, 0), (100, 64), (41, 32), (333, 0)])
def test_prefix_chunkings_equal_direct(ldm, G, bn):
    """every column chunking of the prefix-sum kernel (one chunk, several chunks with pairs straddling the chunk
    boundary, ragged last chunk, chunk wider than the grid) against the sample-by-sample path"""
    props = ldm.Properties.Generate(1, 0.12, 0.09, 0.4, 0.2, 0.001, G)
    direct = ldm.Equation(props, algorithm=ldm.ALGO_DIRECT)
    if bn:
        ldm.tuning_set("scan_bn", bn)
    prefix = ldm.Equation(props, algorithm=ldm.ALGO_PREFIX)
    g = props.MomentsGrid
    st = direct.GenerateEquilibriumValuesForNAndC(200)
    st[1:] *= (1 + 0.5 * np.exp(-(g[:, None] ** 2 + g[None, :] ** 2) / (2 * 0.1 ** 2)) + 0.1 * np.sin(7 * g[:, None] + 3 * g[None, :])).ravel()
    (w1, m1), (w2, m2) = direct.integrals(st), prefix.integrals(st)
    assert abs(w1 - w2) <= 1e-14 * abs(w1)
    assert np.all(np.abs(m1 - m2) <= 1e-13 * np.abs(m1))
    for eq in (direct, prefix):
        eq.upload(st)
        eq.advance(5, 0.1)
    assert_states_close(prefix.download(), direct.download())


@pytest.mark.parametrize("G,sw,sm", [(4, 0.1, 0.1), (7, 0.05, 0.12), (33, 0.2, 0.02), (31, 0.012, 0.1), (12, 0.05, 1e-8)])
def test_tiny_and_degenerate_grids_equal_direct(ldm, G, sw, sm):
    """edges of the table shapes: the smallest grids forced_G admits, one y block that is not full, a refined kernel grid
    (sigma < 0.015, S:80-84), and the degenerate m-kernel of experiment 14.6 b entry 0 (sigma_M -> 0: one quadrature node,
    MaxNumForM == 2, S:131-132) -- every algorithm against the sample-by-sample sum, through a few Euler steps"""
    props = ldm.Properties.Generate(1, sw, sm, 0.4, 0.2, 0.001, G)
    direct = ldm.Equation(props, algorithm=ldm.ALGO_DIRECT)
    g = props.MomentsGrid
    st = direct.GenerateEquilibriumValuesForNAndC(200)
    st[1:] *= (1 + 0.4 * np.exp(-(g[:, None] ** 2 + g[None, :] ** 2) / (2 * (sw + sm) ** 2)) + 0.05 * np.cos(5 * g[:, None] - 2 * g[None, :])).ravel()
    w1, m1 = direct.integrals(st)
    direct.upload(st)
    direct.advance(6, 0.1)
    want = direct.download()
    for algorithm in ("prefix", "factored"):
        eq = ldm.Equation(props, algorithm=algo_id(ldm, algorithm))
        w2, m2 = eq.integrals(st)
        assert abs(w1 - w2) <= 1e-14 * abs(w1) and np.all(np.abs(m1 - m2) <= 1e-13 * np.abs(m1)), algorithm
        eq.upload(st)
        eq.advance(6, 0.1)
        assert_states_close(eq.download(), want)


def test_table_that_is_not_separable_takes_the_dense_contraction(ldm):
    """ldm_create accepts any K x K table; the prefix-sum path is only taken when the table passes the structure check
    (a truncated Gaussian does).  A table with one perturbed entry must fall back to the dense contraction -- and still
    equal the sample-by-sample sum over that very table."""
    props = ldm.Properties.Generate(1, 0.1, 0.1, 0.45, 0.25, 0.001, 64)
    table = np.ctypeslib.as_array(props._ptr.contents.M, (props.K, props.K))
    table[3, 5] *= 1.01
    table[5, 3] *= 1.01
    eq = ldm.Equation(props)
    with pytest.raises(ldm.LdmError):
        eq.set_algorithm(ldm.ALGO_PREFIX)
    direct = ldm.Equation(props, algorithm=ldm.ALGO_DIRECT)
    g = props.MomentsGrid
    st = eq.GenerateEquilibriumValuesForNAndC(200)
    st[1:] *= (1 + 0.5 * np.exp(-(g[:, None] ** 2 + g[None, :] ** 2) / (2 * 0.1 ** 2))).ravel()
    (w1, m1), (w2, m2) = direct.integrals(st), eq.integrals(st)
    assert np.all(np.abs(m1 - m2) <= 1e-13 * np.abs(m1))
    table[3, 5] /= 1.01
    table[5, 3] /= 1.01
    clean = ldm.Equation(props)
    clean.set_algorithm(ldm.ALGO_PREFIX)
    (_, m3) = clean.integrals(st)
    assert np.max(np.abs(m3 - m2) / np.abs(m2)) > 1e-9      # the perturbation was visible, so the check above means something


@pytest.mark.parametrize("name", ["large_g128", "large_g512"])
def test_headline_grids_against_the_reference(ldm, name):
    """BASELINE configs 3 and 4 (16 384 and 262 144 xi-points): one right-hand side computed by the reference itself
    (by-reference build, 35 s / ~25 min of CPU; tests/golden/make_golden.py --large) against the GPU."""
    path = os.path.join(HERE, "golden", name + ".npz")
    if not os.path.exists(path):
        pytest.skip("fixture not generated")
    g = np.load(path)
    G = int(g["G"])
    props = ldm.Properties.Generate(1, float(g["sigmaW"]), float(g["sigmaM"]), float(g["b"]), float(g["d"]), float(g["dd"]), G)
    assert (props.K, props.MaxNumForM, props.MaxNumForW) == (int(g["K"]), int(g["mM"]), int(g["mW"]))
    assert np.array_equal(props.MomentsGrid, g["moments_grid"]) and np.array_equal(props.KernelsGrid, g["kernels_grid"])
    eq = ldm.Equation(props)
    grid = props.MomentsGrid
    st = eq.GenerateEquilibriumValuesForNAndC(200)
    st[1:] = (200.0 ** 2 * (1 + 0.5 * np.exp(-(grid[:, None] ** 2 + grid[None, :] ** 2) / (2 * float(g["sigmaM"]) ** 2)))).ravel()
    got = eq(0.0, st, ldm.SquaredN)

    class Rates:       # the scale only needs the rates; the integrals come from the GPU itself
        pass
    o = Rates()
    o.G, o.b, o.d, o.dd = G, float(g["b"]), float(g["d"]), float(g["dd"])
    scale = rhs_term_scale(o, st, integrals=eq.integrals(st), nodes=eq.kernels_at_nodes())
    err = np.abs(got - g["deriv"])
    worst = np.argmax(err / scale)
    assert np.all(err <= RTOL_RHS_TERMS * scale), (worst, got[worst], g["deriv"][worst], err[worst] / scale[worst])


def test_eight_warp_contraction_tiles(ldm):
    """the 8-warp tiles of the dense contraction (the default runs 16 warps per tile) compute the same integrals"""
    g = golden("synth_g64")
    props, o = problems_from(g, "synth_g64", ldm)
    ldm.tuning_set("factored_tn", 8)
    eq = ldm.Equation(props, algorithm=ldm.ALGO_FACTORED)
    wc, mc = eq.integrals(g["state"])
    assert abs(wc - float(g["wc"])) <= 1e-13 * abs(float(g["wc"]))
    assert np.all(np.abs(mc - g["mc"]) <= 1e-13 * np.abs(g["mc"]))
    big = ldm.Properties.Generate(1, 0.12, 0.09, 0.4, 0.2, 0.001, 200)      # two column chunks, two beta tiles
    a = ldm.Equation(big, algorithm=ldm.ALGO_FACTORED)
    ldm.tuning_set("reset")
    b = ldm.Equation(big, algorithm=ldm.ALGO_FACTORED)
    st = b.GenerateEquilibriumValuesForNAndC(200)
    gg = big.MomentsGrid
    st[1:] *= (1 + 0.3 * np.exp(-(gg[:, None] ** 2 + gg[None, :] ** 2) / 0.02)).ravel()
    (_, m1), (_, m2) = a.integrals(st), b.integrals(st)
    assert np.all(np.abs(m1 - m2) <= 1e-13 * np.abs(m2))


@pytest.mark.parametrize("algorithm", ["prefix", "factored"])
def test_persistent_kernel_equals_per_step_launches(ldm, algorithm):
    """the cooperative multi-step kernel (all Euler steps in one launch, the default where it fits) against the
    three-launches-per-step path: same arithmetic, so the same bits; and every record mode of ldm_solve agrees"""
    g = golden("exp_14_4_c")
    props, _ = problems_from(g, "exp_14_4_c", ldm)
    persistent = ldm.Equation(props, algorithm=algo_id(ldm, algorithm))
    ldm.tuning_set("persistent", 0)
    stepwise = ldm.Equation(props, algorithm=algo_id(ldm, algorithm))
    init = persistent.GenerateEquilibriumValuesForNAndC(200)
    runs = [ldm.EulerSolver().Solve(eq, ldm.SquaredN, 30, 0.1, init, record=ldm.RECORD_ALL) for eq in (persistent, stepwise)]
    assert persistent.launches < stepwise.launches / 50
    assert np.array_equal(runs[0].Values, runs[1].Values)
    assert np.array_equal(runs[0].Derivatives, runs[1].Derivatives)
    obs = ldm.EulerSolver().Solve(persistent, ldm.SquaredN, 30, 0.1, init, record=ldm.RECORD_OBSERVABLES)
    assert np.array_equal(obs.Observables, runs[0].Values[:, [0, props.G]])
    for eq in (persistent, stepwise):
        eq.upload(init)
        eq.advance(7, 0.1)      # odd step count: the ping-pong parity must come out right
        eq.advance(4, 0.1)
    assert np.array_equal(persistent.download(), stepwise.download())
    assert np.array_equal(persistent.download(), runs[0].Values[11])


@pytest.mark.parametrize("G,bn", [(200, 0), (200, 32), (333, 0), (41, 32)])
def test_work_list_kernel_equals_per_step_launches(ldm, G, bn):
    """grids whose rows span several chunks: the persistent work-list kernel (all steps of ldm_advance / of a history
    segment of ldm_solve in one launch) against one launch per kernel and step -- same arithmetic, same bits, in every
    record mode; several history segments per solve at the larger sizes"""
    props = ldm.Properties.Generate(1, 0.12, 0.09, 0.4, 0.2, 0.001, G)
    if bn:
        ldm.tuning_set("scan_bn", bn)
    ldm.tuning_set("persistent", 1)       # the work-list kernel on one rank (the default there is one launch per kernel and step)
    persistent = ldm.Equation(props)
    ldm.tuning_set("persistent", 0)
    stepwise = ldm.Equation(props)
    g = props.MomentsGrid
    init = persistent.GenerateEquilibriumValuesForNAndC(200)
    init[1:] *= (1 + 0.5 * np.exp(-(g[:, None] ** 2 + g[None, :] ** 2) / (2 * 0.1 ** 2)) + 0.1 * np.sin(7 * g[:, None] + 3 * g[None, :])).ravel()
    points = 13
    runs = [ldm.EulerSolver().Solve(eq, ldm.SquaredN, (points - 0.5) * 0.1, 0.1, init, record=ldm.RECORD_ALL) for eq in (persistent, stepwise)]
    assert len(runs[0].Times) == points
    assert persistent.launches < stepwise.launches / 4
    assert np.array_equal(runs[0].Values, runs[1].Values)
    assert np.array_equal(runs[0].Derivatives, runs[1].Derivatives)
    obs = ldm.EulerSolver().Solve(persistent, ldm.SquaredN, (points - 0.5) * 0.1, 0.1, init, record=ldm.RECORD_OBSERVABLES)
    assert np.array_equal(obs.Observables, runs[0].Values[:, [0, props.G]])
    last = ldm.EulerSolver().Solve(persistent, ldm.SquaredN, (points - 0.5) * 0.1, 0.1, init, record=ldm.RECORD_LAST)
    assert np.array_equal(last.Values[0], runs[0].Values[-1]) and np.array_equal(last.Derivatives[0], runs[0].Derivatives[-1])
    for eq in (persistent, stepwise):
        eq.upload(init)
        eq.advance(7, 0.1)      # odd step count: the ping-pong parity must come out right
        eq.advance(4, 0.1)
    assert np.array_equal(persistent.download(), stepwise.download())
    assert np.array_equal(persistent.download(), runs[0].Values[11])


def test_history_leaves_in_segments(ldm):
    """a solve whose history spans many device segments (16 MiB each): Values and Derivatives of every point arrive,
    identical to a solve that stores the observables only"""
    props = ldm.Properties.Generate(1, 0.15, 0.15, 0.4, 0.2, 0.001, 128)
    eq = ldm.Equation(props)
    init = eq.GenerateEquilibriumValuesForNAndC(200)
    points = 700                                   # 700 x 16385 x 8 B x 2 = 183 MB of history
    full = ldm.EulerSolver().Solve(eq, ldm.SquaredN, (points - 0.5) * 0.1, 0.1, init, record=ldm.RECORD_ALL)
    obs = ldm.EulerSolver().Solve(eq, ldm.SquaredN, (points - 0.5) * 0.1, 0.1, init, record=ldm.RECORD_OBSERVABLES)
    assert np.array_equal(full.Values[:, [0, props.G]], obs.Observables)
    assert np.array_equal(full.Values[0], init)
    eq.upload(full.Values[400])
    eq.advance(1, 0.1)
    assert np.array_equal(eq.download(), full.Values[401])
    assert np.array_equal(eq(0.0, full.Values[650]), full.Derivatives[650])


@pytest.mark.parametrize("G", [0, 200])
def test_history_straight_into_pinned_host_buffers(ldm, G):
    """Values / Derivatives / observables buffers in pinned, mapped host memory: the persistent kernel stores the history
    straight into them (one launch for the whole solve); identical to the route through device-side segments that pageable
    buffers take, in every record mode, for a one-chunk grid and a chunked one"""
    import torch
    props = ldm.Properties.Generate(1, 0.12, 0.09, 0.4, 0.2, 0.001, G)
    ldm.tuning_set("persistent", 1)       # chunked grid: the work-list kernel (on one rank it runs only where it is asked for)
    eq = ldm.Equation(props)
    ldm.tuning_set("reset")
    g = props.MomentsGrid
    init = eq.GenerateEquilibriumValuesForNAndC(200)
    init[1:] *= (1 + 0.5 * np.exp(-(g[:, None] ** 2 + g[None, :] ** 2) / (2 * 0.1 ** 2))).ravel()
    points, width = 23, 1 + props.G ** 2
    t_max = (points - 0.5) * 0.1
    want = ldm.EulerSolver().Solve(eq, ldm.SquaredN, t_max, 0.1, init, record=ldm.RECORD_ALL)
    values = torch.full((points * width,), float("nan"), dtype=torch.float64).pin_memory().numpy()
    derivs = torch.full((points * width,), float("nan"), dtype=torch.float64).pin_memory().numpy()
    before = eq.launches
    got = ldm.EulerSolver().Solve(eq, ldm.SquaredN, t_max, 0.1, init, record=ldm.RECORD_ALL, values_out=values, derivs_out=derivs)
    assert eq.launches - before <= 6                            # the stepping is one launch (+ the last derivative's three)
    assert np.array_equal(got.Values, want.Values) and np.array_equal(got.Derivatives, want.Derivatives)
    values[:] = np.nan
    got = ldm.EulerSolver().Solve(eq, ldm.SquaredN, t_max, 0.1, init, record=ldm.RECORD_VALUES, values_out=values)
    assert np.array_equal(got.Values, want.Values)


def test_large_grid_properties(ldm):
    """BASELINE config 3 (G = 128): too slow for the CPU oracle in a test, so check what the domain offers:
    the spatially uniform state C = N^2 must give dC ~ 2bN m(xi) - 2d'N^2 w(xi)... exactly the structure of the
    reference at t=0 -- MC = N^2 and WC = N^2 because both kernels integrate to one (Solver.cpp:147-183)."""
    props = ldm.Properties.Generate(1, 0.15, 0.15, 0.4, 0.2, 0.001, 128)
    eq = ldm.Equation(props)
    st = eq.GenerateEquilibriumValuesForNAndC(200)
    wc, mc = eq.integrals(st)
    assert abs(wc - 40000.0) <= 1e-9 * 40000 and np.all(np.abs(mc - 40000.0) <= 1e-9 * 40000)
    # linearity of the integrals in C at fixed N
    rng = np.random.default_rng(7)
    a = st.copy(); a[1:] *= rng.uniform(0.5, 1.5, a.size - 1)
    c = st.copy(); c[1:] *= rng.uniform(0.5, 1.5, c.size - 1)
    s = a.copy(); s[1:] = 0.25 * a[1:] + 0.75 * c[1:]
    (wa, ma), (wb, mb), (ws, ms) = eq.integrals(a), eq.integrals(c), eq.integrals(s)
    assert abs(ws - (0.25 * wa + 0.75 * wb)) <= 1e-12 * abs(ws)
    assert np.all(np.abs(ms - (0.25 * ma + 0.75 * mb)) <= 1e-12 * np.abs(ms))
