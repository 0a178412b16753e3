"""The closures whose triple integral does not factorise (ldm.h ldm_set_closure; csrc/ldm_closure.cu), through the C ABI,
against the CPU oracle of the extension (oracle/ldm_oracle_closure.c) -- and, where the reference has an answer (the uniform
state, power-1), against the reference's own fixtures.  Tolerances as in test_gpu_parity.py: 1e-13 on an integral, 1e-12 of
the sum of the term magnitudes on a right-hand side, 1e-10 on every component of a trajectory."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import OracleProblem
from test_closure_oracle import bumped_state
from test_gpu_parity import algo_id, assert_states_close, golden, problems_from, rhs_term_scale

pytestmark = pytest.mark.gpu

CASES = ["exp_14_4_b", "exp_14_4_c", "exp_14_6_a_080", "synth_g32"]


def closure_term_scale(o, state, closure, w_conv, k):
    """rhs_term_scale plus the magnitudes of the terms that stand in for S:486-487"""
    N, C = state[0], state[1:].reshape(o.G, o.G)
    scale = rhs_term_scale(o, state)
    wc, _ = o.integrals(state, want_mc=False)
    if closure == 3:
        extra = 2 * o.dd * np.abs(C * k) / N ** 3
    else:
        extra = o.dd * (np.abs(wc * C / N) + np.abs(C * w_conv / N) + np.abs(k / N) + N ** 3 * o.w_mass())
    scale[1:] += extra.ravel()
    return scale


def a_state(g, name, o):
    return g["state"] if "synth" in name else bumped_state(o)


@pytest.mark.parametrize("name", CASES)
def test_closure_integrals_against_oracle(ldm, name):
    g = golden(name)
    props, o = problems_from(g, name, ldm)
    eq = ldm.Equation(props)
    for state in (eq.GenerateEquilibriumValuesForNAndC(200), a_state(g, name, o)):
        w_conv, k = eq.closure_integrals(state)
        w_conv_ref, k_ref = o.closure_integrals(state)
        assert np.all(np.abs(w_conv - w_conv_ref) <= 1e-13 * np.abs(w_conv_ref))
        assert np.all(np.abs(k - k_ref) <= 1e-13 * np.abs(k_ref))
    assert eq.closure == ldm.CLOSURE_POWER1                    # asking for the integrals selects nothing


@pytest.mark.parametrize("algorithm", ["direct", "factored", "prefix"])
@pytest.mark.parametrize("name", CASES)
def test_rhs_under_the_closures(ldm, name, algorithm):
    g = golden(name)
    props, o = problems_from(g, name, ldm)
    eq = ldm.Equation(props, algorithm=algo_id(ldm, algorithm))
    state = a_state(g, name, o)
    reference_closure = eq(0.0, state, ldm.SquaredN)
    w_conv, k = o.closure_integrals(state)
    for closure in (ldm.CLOSURE_POWER2_SYMMETRIC, ldm.CLOSURE_POWER3):
        eq.set_closure(closure)
        got = eq(0.0, state, ldm.SquaredN)
        want = o.rhs_closure(state, closure)
        scale = closure_term_scale(o, state, closure, w_conv, k)
        err = np.abs(got - want)
        worst = np.argmax(err / scale)
        assert np.all(err <= 1e-12 * scale), (closure, worst, got[worst], want[worst], err[worst] / scale[worst])
        assert got[0] == reference_closure[0]                  # the first-moment equation needs no closure
        assert np.max(np.abs(got[1:] - reference_closure[1:])) > 1e-6 * np.max(np.abs(reference_closure[1:]))   # and it is not a no-op
    eq.set_closure(ldm.CLOSURE_POWER1)
    assert np.array_equal(eq(0.0, state, ldm.SquaredN), reference_closure)      # the way back: the same bits as before


@pytest.mark.parametrize("name", ["exp_14_4_a", "exp_14_4_b", "exp_14_4_c", "exp_14_6_a_080"])
def test_every_closure_meets_the_reference_at_the_uniform_state(ldm, name):
    """C == N^2: T = N^3 under every closure, so the right-hand side the REFERENCE computed there is the known answer"""
    g = golden(name)
    props, o = problems_from(g, name, ldm)
    eq = ldm.Equation(props)
    state = eq.GenerateEquilibriumValuesForNAndC(200)
    N = 200.0
    m_nodes, w_nodes = o.kernels_at_nodes()
    scale = (2 * o.b * N * N + 2 * o.b * N * np.abs(m_nodes) + 2 * o.d * N * N + 4 * o.dd * N ** 3 + 2 * o.dd * np.abs(w_nodes) * N * N).ravel()
    for closure in (ldm.CLOSURE_POWER2_SYMMETRIC, ldm.CLOSURE_POWER3):
        eq.set_closure(closure)
        got = eq(0.0, state, ldm.SquaredN)
        assert abs(got[0] - g["deriv_k0"][0]) <= 1e-12 * (abs(N * (o.b - o.d)) + o.dd * N * N)
        assert np.all(np.abs(got[1:] - g["deriv_k0"][1:]) <= 1e-12 * scale)


@pytest.mark.parametrize("closure", [2, 3])
def test_trajectory_under_the_closures(ldm, closure):
    """EulerSolver::Solve on the exchanged right-hand side, every record mode and the resident stepping; 14.4b, 40 steps
    from a state away from the uniform one"""
    props = ldm.Properties.Generate(1, 0.05, 0.05, 0.4, 0.2, 0.001)
    o = OracleProblem(0.05, 0.05)
    init = bumped_state(o)
    eq = ldm.Equation(props)
    eq.set_closure(closure)
    t_max = 4.05
    res = ldm.EulerSolver().Solve(eq, ldm.SquaredN, t_max, 0.1, init, record=ldm.RECORD_ALL)
    times, want = o.solve_closure(closure, t_max, 0.1, init)
    assert np.array_equal(res.Times, times)
    assert_states_close(res.Values, want)
    last = o.rhs_closure(want[-1], closure)
    assert np.all(np.abs(res.Derivatives[-1] - last) <= 1e-9 * np.max(np.abs(last)))
    # the closure moves the trajectory: it is not the reference's run
    _, power1, _, _ = o.solve(t_max, 0.1, init)
    assert abs(want[-1, 0] - power1[-1, 0]) > 1e-3
    for record in (ldm.RECORD_VALUES, ldm.RECORD_OBSERVABLES, ldm.RECORD_LAST):
        other = ldm.EulerSolver().Solve(eq, ldm.SquaredN, t_max, 0.1, init, record=record)
        if record == ldm.RECORD_VALUES:
            assert np.array_equal(other.Values, res.Values)
        elif record == ldm.RECORD_OBSERVABLES:
            assert np.array_equal(other.Observables, res.Values[:, [0, props.G]])
        else:
            assert np.array_equal(np.ravel(other.Values), res.Values[-1])
    eq.upload(init)
    eq.advance(len(times) - 1, 0.1)
    assert np.array_equal(eq.download(), res.Values[-1])
    # the Runge-Kutta solver runs on the same right-hand side: its first stage is this equation at the initial state
    rk = ldm.RungeKutta4thSolver().Solve(eq, ldm.SquaredN, 0.25, 0.1, init)
    assert np.array_equal(rk.Derivatives[0], res.Derivatives[0])


def test_batch_entries_with_closures(ldm):
    """ldm_solve_batch takes such handles: they advance with one launch per kernel and step beside the others"""
    props = ldm.Properties.Generate(1, 0.05, 0.05, 0.4, 0.2, 0.001)
    o = OracleProblem(0.05, 0.05)
    init = bumped_state(o)
    eqs = [ldm.Equation(props) for _ in range(3)]
    eqs[1].set_closure(ldm.CLOSURE_POWER2_SYMMETRIC)
    eqs[2].set_closure(ldm.CLOSURE_POWER3)
    outs = ldm.solve_batch(eqs, [init] * 3, 0.1, 2.05)
    for eq, out in zip(eqs, outs):
        alone = ldm.EulerSolver().Solve(eq, ldm.SquaredN, 2.05, 0.1, init, record=ldm.RECORD_LAST)
        assert np.array_equal(np.ravel(out), np.ravel(alone.Values))
    assert not np.array_equal(outs[0], outs[1]) and not np.array_equal(outs[1], outs[2])


def test_what_the_closures_are_not_built_for_is_refused(ldm):
    props = ldm.Properties.Generate(1, 0.05, 0.05, 0.4, 0.2, 0.001)
    eq = ldm.Equation(props)
    with pytest.raises(ldm.LdmError):
        eq.set_closure(7)
    eq.set_closure(ldm.CLOSURE_POWER3)
    with pytest.raises(ldm.LdmError, match="SquaredN"):
        eq.set_extrapolation(ldm.Zero)
    assert ldm.library().ldm_shard_configure(eq._h, 0, 2) == 0           # xi-sharding: ldm_shard_step carries the closures ...
    assert ldm.library().ldm_shard_peer_connect_local(1, (C.c_void_p * 1)(eq._h)) == -3      # ... the fused exchange does not
    assert ldm.library().ldm_shard_configure(eq._h, 0, 1) == 0
    eq.set_closure(ldm.CLOSURE_POWER1)
    eq.set_extrapolation(ldm.Zero)
    with pytest.raises(ldm.LdmError, match="SquaredN"):
        eq.set_closure(ldm.CLOSURE_POWER3)
    two = ldm.Properties.Generate(2, [[0.05, 0.06], [0.06, 0.05]], [0.05] * 2, [0.4, 0.401], [0.2] * 2, np.full((2, 2), 0.001))
    with pytest.raises(ldm.LdmError, match="one species"):
        ldm.Equation(two).set_closure(ldm.CLOSURE_POWER3)


@pytest.mark.parametrize("name", ["exp_14_4_b", "exp_14_4_a", "synth_g32"])
def test_every_variant_of_the_sums_kernel_adds_the_same_numbers(ldm, name):
    """the sums kernel holds a lane's offsets in registers where the quadrature is at most 64 offsets wide and reads them
    per sample otherwise: the same additions in the same order, so the same bits; the width picks the variant,
    ldm_tuning_set("closure_tiles") forces one (the held variant falls back to per-sample reads on a wider quadrature)"""
    g = golden(name)
    props, o = problems_from(g, name, ldm)
    eq = ldm.Equation(props)
    state = a_state(g, name, o)
    try:
        results = []
        for tiles in (-1, 0, 2):
            ldm.tuning_set("closure_tiles", tiles)
            results.append(eq.closure_integrals(state))
            eq.set_closure(ldm.CLOSURE_POWER2_SYMMETRIC)
            results[-1] += (eq(0.0, state, ldm.SquaredN),)
            eq.set_closure(ldm.CLOSURE_POWER1)
    finally:
        ldm.tuning_set("reset")
    for other in results[1:]:
        for a, b in zip(results[0], other):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("closure", [2, 3])
@pytest.mark.parametrize("world", [2, 3, 5])
def test_sharded_step_under_the_closures_emulated_on_one_device(ldm, world, closure):
    """xi-sharding of the closure integrals, all ranks emulated on one GPU as in test_gpu_parity.py: each rank
    computes (w*C), K and the update for ITS rows from the gathered replica; the 'all-gather' is a concatenation.  The rows
    are independent given C, so the result must equal the unsharded run bit for bit."""
    import torch
    g = golden("exp_14_4_b")
    props, o = problems_from(g, "exp_14_4_b", ldm)
    G = props.G
    ranks = [ldm.ShardedEquation(props, r, world) for r in range(world)]
    for r in ranks:
        r.eq.set_closure(closure)
    rpr = ranks[0].rows_per_rank
    state = bumped_state(o)
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
    single.set_closure(closure)
    single.upload(state)
    single.advance(steps, 0.1)
    want = single.download()
    got_C = ranks[0].from_gathered(bufs[steps % 2].cpu().numpy())
    assert np.array_equal(got_C.ravel(), want[1:])
    for r in ranks:
        assert r.get_N() == want[0]
