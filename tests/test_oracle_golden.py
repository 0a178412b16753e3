"""The CPU oracle (oracle/ldm_oracle.c) against the golden fixtures generated from the reference itself
(tests/golden/make_golden.py): bitwise on grids and kernel tables, bitwise on the right-hand side and on
whole trajectories -- the oracle is a restatement, not an approximation."""
import glob
import os

import numpy as np
import pytest

from oracle import OracleMulti, OracleProblem

GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")) if not os.path.basename(p).startswith(("large_", "extras_", "multi_", "exp_14_6_sweep")))


def make_oracle(g, name):
    forced = int(g["G"]) if "synth" in name else 0
    return OracleProblem(float(g["sigmaW"]), float(g["sigmaM"]), float(g["b"]), float(g["d"]), float(g["dd"]), forced)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_problem_tables_bitwise(path):
    g = np.load(path)
    o = make_oracle(g, path)
    assert (o.G, o.K, o.mM, o.mW) == (int(g["G"]), int(g["K"]), int(g["mM"]), int(g["mW"]))
    assert o.h_step == float(g["h"]) and o.cutoff == float(g["cutoff"])
    mg, kg = o.grids()
    M, W = o.tables()
    assert np.array_equal(mg, g["moments_grid"]) and np.array_equal(kg, g["kernels_grid"])
    assert np.array_equal(M, g["M"]) and np.array_equal(W, g["W"])
    m_nodes, w_nodes = o.kernels_at_nodes()
    assert np.array_equal(m_nodes, g["m_nodes"]) and np.array_equal(w_nodes, g["w_nodes"])


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_rhs_bitwise(path):
    g = np.load(path)
    o = make_oracle(g, path)
    if o.P * (o.mM - 1) ** 2 > 3e7:
        pytest.skip("more than a few seconds of CPU")
    if "synth" in path:
        assert np.array_equal(o.rhs(g["state"]), g["deriv"])
        wc, mc = o.integrals(g["state"])
        assert wc == float(g["wc"]) and np.array_equal(mc, g["mc"])
    else:
        assert np.array_equal(o.rhs(o.equilibrium()), g["deriv_k0"])
        assert np.array_equal(o.rhs(g["state_k100"]), g["deriv_k100"])


@pytest.mark.parametrize("name", ["exp_14_4_b", "exp_14_4_c", "exp_14_6_a_012", "exp_14_6_b_000", "exp_14_6_b_012",
                                  "exp_14_6_b_040"])
def test_trajectory_bitwise(name):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    o = make_oracle(g, name)
    t_max = 100.0 if "14_4" in name else 200.0
    times, values, _, obs = o.solve(t_max, full_history=False)
    assert len(times) == len(g["times"]) and np.array_equal(times, g["times"])
    assert np.array_equal(obs, g["obs"])
    assert np.array_equal(values[0], g["final_state"])


def test_known_answers_of_the_survey():
    """SURVEY.md 8c / BASELINE.md 3: numbers printed by the unmodified reference."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "exp_14_4_b.npz"))
    assert repr(float(g["obs"][1000, 0])) == "89.466775880443"
    assert float(g["final_state"][1]) == 20572.31205740929
    assert float(g["times"][-1]) == 99.99999999999859
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "exp_14_6_a_000.npz"))
    assert float(g["obs"][2000, 0]) == 5490.198333335871
    assert float(g["obs"][100, 0]) == 1134.0086006319391


def test_point_count_of_the_accumulated_clock():
    o = OracleProblem(0.05, 0.05)
    assert o.count_points(100.0) == 1001 and o.count_points(200.0) == 2001


@pytest.mark.parametrize("name", ["b", "c"])
def test_runge_kutta_and_extrapolation_types_bitwise(name):
    """RungeKutta4thSolver (Solver.cpp:597-654) and the Zero / Alpha extrapolation of C
    (Solver.cpp:361-389) as restated in the oracle, against the reference itself (tests/golden/make_golden.py --extras)"""
    path = os.path.join(os.path.dirname(__file__), "golden", f"extras_14_4_{name}.npz")
    if not os.path.exists(path):
        pytest.skip("fixture not generated")
    g = np.load(path)
    base = np.load(os.path.join(os.path.dirname(__file__), "golden", f"exp_14_4_{name}.npz"))
    o = OracleProblem(float(g["sigmaW"]), float(g["sigmaM"]))
    mid = base["state_k100"]
    try:
        steps = 10 if name == "c" else 40              # 14.4c: 0.2 s per right-hand side on one core
        t, v, d = o.solve_rk4(steps * 0.1 - 0.05)
        assert np.array_equal(t, g["rk4_times"][:steps + 1]) and np.array_equal(v, g["rk4_values"][:steps + 1])
        assert np.array_equal(d[:steps], g["rk4_derivs"][:steps])
        for kind, label in ((0, "zero"), (2, "alpha"), (3, "exponential")):
            if f"{label}_deriv_k0" not in g.files:
                continue
            o.set_extrapolation(kind)
            assert np.array_equal(o.rhs(o.equilibrium()), g[f"{label}_deriv_k0"])
            assert np.array_equal(o.rhs(mid), g[f"{label}_deriv_k100"])
            t, v, d, _ = o.solve(steps * 0.1 - 0.05, init=mid, want_derivs=True)
            assert len(v) == steps and np.array_equal(v, g[f"{label}_values"][:steps])     # Euler stores one point fewer (S:574-575)
            # the types are not a no-op: away from the uniform state the right-hand side differs from SquaredN's
            assert not np.array_equal(g[f"{label}_deriv_k100"], base["deriv_k100"])
    finally:
        o.set_extrapolation(1)


@pytest.mark.parametrize("name", ["2_species", "3_species"])
def test_several_species_bitwise(name):
    """Equation::operator() and EulerSolver with SpeciesNumber > 1 (Solver.cpp:265-295, 446-511) as
    restated in the oracle, against the reference itself on its own multi-species example inputs"""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"multi_{name}.npz"))
    s = int(g["species"])
    sw = {2: [[0.05, 0.06], [0.06, 0.05]], 3: [[0.05, 0.06, 0.07], [0.06, 0.05, 0.04], [0.07, 0.04, 0.05]]}[s]
    o = OracleMulti(g["mM"], g["mW"], [0.05] * s, sw, [0.4, 0.401, 0.4001][:s], [0.2] * s, np.full((s, s), 0.001),
                    g["moments_grid"], g["kernels_grid"], g["M"], g["W"])
    assert np.array_equal(o.rhs(g["state"]), g["deriv"])
    assert np.array_equal(o.rhs(np.concatenate([np.full(s, 200.0), np.full(s * s * o.G ** 2, 40000.0)])), g["deriv_equilibrium"])
    steps = 10
    t, v = o.solve(steps * 0.1 + 0.05, 0.1, g["state"])
    assert np.array_equal(t, g["times"][:steps + 1]) and np.array_equal(v, g["values"][:steps + 1])
