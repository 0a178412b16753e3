#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ from the REFERENCE ITSELF.

Runs only where /root/reference is mounted and oracle/_ref has been built (`make -C oracle ref`):
the reference's Solver.cpp is compiled unmodified (libldm_ref.so) and with five by-value parameters
turned into references (libldm_ref_byref.so, ~30x faster).  The script first proves the two builds
bit-identical on every cheap configuration, then uses the fast one where the shipped code would
need hours.  The fixtures travel to the GPU box; /root/reference does not.

    python tests/golden/make_golden.py            # everything (~6 min on 8 cores)
"""
import json
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from refharness import REF_DATA, Reference  # noqa: E402

CASES_14_4 = {"a": (0.15, 0.15), "b": (0.05, 0.05), "c": (0.015, 0.15)}   # sigma_W, sigma_M (Data/14/4/*.csv)
SWEEP = {"a": [0, 12, 80, 159], "b": [0, 12, 40, 159]}                    # Experiment_14_6 entries (S:1033-1043)


def sigma_vec():
    """Vector::setLinSpaced(160, EPSILON, 0.2) (Solver.cpp:1042, types.h:120-130)"""
    num, lo, hi = 160.0, 1e-08, 0.2
    step = (hi - lo) / (num - 1)
    return np.array([lo + i * step for i in range(159)] + [hi])


def problem_arrays(p):
    g, k = p.grids()
    M, W = p.tables()
    m_nodes, w_nodes = p.kernels_at_nodes()
    return dict(G=p.G, K=p.K, mM=p.mM, mW=p.mW, h=p.h_step, cutoff=p.cutoff, sigmaM=p.sigmaM, sigmaW=p.sigmaW,
                b=p.b, d=p.d, dd=p.dd, moments_grid=g, kernels_grid=k, M=M, W=W, m_nodes=m_nodes, w_nodes=w_nodes)


def job_14_4(case):
    fast = Reference(byref=True)
    p = fast.from_file(f"{REF_DATA}/14/4/{case}.csv")
    out = problem_arrays(p)
    t0 = time.time()
    times, values, derivs = p.solve(100.0, want_derivs=True)
    out.update(times=times, obs=np.stack([values[:, 0], values[:, p.G]], 1), final_state=values[-1],
               final_deriv=derivs[-1], state_k100=values[100], deriv_k100=derivs[100], deriv_k0=derivs[0],
               state_k1=values[1], state_k10=values[10], state_k500=values[500])
    wc, mc = p.integrals(values[100])
    out.update(wc_k100=wc, mc_k100=mc)
    out["seconds_byref"] = time.time() - t0
    # the shipped build, wherever it finishes in about a minute
    slow = Reference(byref=False)
    q = slow.from_file(f"{REF_DATA}/14/4/{case}.csv")
    for key, val in problem_arrays(q).items():
        assert np.array_equal(val, out[key]), (case, key)
    assert np.array_equal(q.rhs(values[0]), derivs[0]), case
    assert np.array_equal(q.rhs(values[100]), derivs[100]), case
    if case == "b":
        t0 = time.time()
        t2, v2, d2 = q.solve(100.0, want_derivs=True)
        assert np.array_equal(t2, times) and np.array_equal(v2, values) and np.array_equal(d2, derivs)
        out["seconds_unmodified"] = time.time() - t0
        out["unmodified_checked"] = "problem tables, rhs(k=0), rhs(k=100), whole trajectory"
    else:
        out["unmodified_checked"] = "problem tables, rhs(k=0), rhs(k=100)"
    np.savez_compressed(os.path.join(HERE, f"exp_14_4_{case}.npz"), **out)
    return f"14.4{case}: G={p.G} N_end={values[-1, 0]!r}"


def job_14_6(args):
    which, i = args
    sig = sigma_vec()[i]
    fast = Reference(byref=True)
    route = 1 if which == "a" else 2          # a: GetWithoutSigmaW, b: GetWithoutSigmaM (S:1035-1041)
    p = fast.from_file(f"{REF_DATA}/14/6/{which}.csv", route, float(sig))
    out = problem_arrays(p)
    t0 = time.time()
    times, values, derivs = p.solve(200.0, want_derivs=True)
    out.update(sigma=sig, entry=i, times=times, obs=np.stack([values[:, 0], values[:, p.G]], 1),
               final_state=values[-1], state_k100=values[100], deriv_k100=derivs[100], deriv_k0=derivs[0],
               state_k1=values[1], state_k2=values[2])
    out["seconds_byref"] = time.time() - t0
    slow = Reference(byref=False)
    q = slow.from_file(f"{REF_DATA}/14/6/{which}.csv", route, float(sig))
    for key, val in problem_arrays(q).items():
        assert np.array_equal(val, out[key]), (which, i, key)
    out["unmodified_checked"] = "problem tables"
    if p.G * p.G * (p.mM - 1) ** 2 < 1.2e6:   # one shipped-build RHS in a few seconds
        assert np.array_equal(q.rhs(values[100]), derivs[100]), (which, i)
        out["unmodified_checked"] += ", rhs(k=100)"
    np.savez_compressed(os.path.join(HERE, f"exp_14_6_{which}_{i:03d}.npz"), **out)
    return f"14.6{which}[{i}]: sigma={sig!r} G={p.G} K={p.K} mM={p.mM} mW={p.mW} N_end={values[-1, 0]!r}"


def job_sweep_entry(args):
    """One entry of Experiment_14_6 (S:1033-1071) through the by-reference build: what the experiment plots for it --
    the last state's N and C[0][G-1] (S:1050-1055) -- plus the grid sizes and the sum of the last state."""
    which, i = args
    sig = sigma_vec()[i]
    p = Reference(byref=True).from_file(f"{REF_DATA}/14/6/{which}.csv", 1 if which == "a" else 2, float(sig))
    t0 = time.time()
    times, values, _ = p.solve(200.0)
    last = values[-1]
    return dict(which=which, entry=i, sigma=sig, G=p.G, K=p.K, mM=p.mM, mW=p.mW, points=len(times), N_end=last[0],
                C_end=last[p.G], state_sum=float(last.sum()), C00_end=last[1], seconds=time.time() - t0)


def make_sweep(workers):
    """ALL 160 entries of both sweeps of Experiment_14_6, pinned to the reference (by-reference build, bit-identical to the
    shipped one on every configuration both were run on): tests/golden/exp_14_6_sweep.npz"""
    jobs = [(w, i) for w in "ab" for i in range(160)]
    jobs.sort(key=lambda j: j[1])                    # the refined-grid entries (sigma < 0.015, ~100x the cost) first
    rows = []
    with ProcessPoolExecutor(max_workers=workers) as pool:
        for r in pool.map(job_sweep_entry, jobs):
            rows.append(r)
            print(f"14.6{r['which']}[{r['entry']}] G={r['G']} K={r['K']} N_end={r['N_end']!r} ({r['seconds']:.0f} s)", flush=True)
    out = {}
    for w in "ab":
        mine = sorted((r for r in rows if r["which"] == w), key=lambda r: r["entry"])
        for key in ("sigma", "G", "K", "mM", "mW", "points", "N_end", "C_end", "state_sum", "C00_end", "seconds"):
            out[f"{w}_{key}"] = np.array([r[key] for r in mine])
    np.savez_compressed(os.path.join(HERE, "exp_14_6_sweep.npz"), **out)


def run_reference_menu(exe, keys, data_dir, workdir):
    """the reference's own program (menu and all) in `workdir`: its relative paths use backslashes (Solver.cpp:983, 993),
    which on Linux are literal file names -- symlinks named `Data\\14\\4\\b.csv` make the unmodified binary run.  Returns
    the text of the `Data\\VisData.csv` it writes."""
    import subprocess
    for root, _, files in os.walk(data_dir):
        for f in files:
            if f.endswith(".csv"):
                rel = os.path.relpath(os.path.join(root, f), os.path.dirname(data_dir))
                link = os.path.join(workdir, rel.replace("/", "\\"))
                if not os.path.lexists(link):
                    os.symlink(os.path.join(root, f), link)
    subprocess.run([exe], cwd=workdir, env=dict(os.environ, LDM_REF_KEYS=keys, TERM="dumb"), stdout=subprocess.DEVNULL,
                   stderr=subprocess.DEVNULL, check=True, timeout=3600)
    return open(os.path.join(workdir, "Data\\VisData.csv")).read()


def make_visdata():
    """Experiment 14.4 (b) through the UNMODIFIED reference program, its own menu driven by scripted keys (down, enter, esc):
    the file it writes, byte for byte -> tests/golden/VisData_14_4_b_reference.csv (48 s of one core)"""
    import tempfile
    exe = os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "solver_ref")
    with tempfile.TemporaryDirectory() as tmp:
        text = run_reference_menu(exe, "80,13,27", REF_DATA, tmp)
    with open(os.path.join(HERE, "VisData_14_4_b_reference.csv"), "w") as fh:
        fh.write(text)
    print("VisData 14.4 (b):", len(text.splitlines()), "lines, last:", text.splitlines()[-1])


def job_synth(args):
    """Forced-G grids (BASELINE.json configs 3-5): problem tables + one RHS on a non-trivial state."""
    name, sw, sm, G, b, d = args
    fast = Reference(byref=True)
    p = fast.from_numbers(sw, sm, b, d, 0.001, G)
    out = problem_arrays(p)
    g, _ = p.grids()
    st = np.full(1 + p.P, 200.0 ** 2)
    st[0] = 200.0
    r2 = g[:, None] ** 2 + g[None, :] ** 2
    st[1:] = (200.0 ** 2 * (1 + 0.5 * np.exp(-r2 / (2 * sw * sw)))).ravel()   # SURVEY 8d config 3 seed, no RNG
    t0 = time.time()
    out.update(state=st, deriv=p.rhs(st))
    wc, mc = p.integrals(st)
    out.update(wc=wc, mc=mc, seconds_byref=time.time() - t0)
    np.savez_compressed(os.path.join(HERE, f"synth_{name}.npz"), **out)
    return f"synth {name}: G={p.G} K={p.K} mM={p.mM} mW={p.mW} dN={out['deriv'][0]!r}"


def job_large(args):
    """The headline grids (BASELINE configs 3 and 4) pinned to the reference itself: ONE right-hand side each with the
    by-reference build (G=128: ~35 s; G=512: ~25 min single-threaded).  Stored without the kernel tables (those are
    covered bitwise at the smaller sizes) to keep the fixtures small."""
    name, sw, sm, G = args
    fast = Reference(byref=True)
    p = fast.from_numbers(sw, sm, 0.4, 0.2, 0.001, G)
    g, k = p.grids()
    st = np.full(1 + p.P, 200.0 ** 2)
    st[0] = 200.0
    r2 = g[:, None] ** 2 + g[None, :] ** 2
    st[1:] = (200.0 ** 2 * (1 + 0.5 * np.exp(-r2 / (2 * sm * sm)))).ravel()
    t0 = time.time()
    deriv = p.rhs(st)
    out = dict(G=p.G, K=p.K, mM=p.mM, mW=p.mW, h=p.h_step, cutoff=p.cutoff, sigmaM=p.sigmaM, sigmaW=p.sigmaW, b=p.b, d=p.d,
               dd=p.dd, moments_grid=g, kernels_grid=k, deriv=deriv, seconds_byref=time.time() - t0)
    np.savez_compressed(os.path.join(HERE, f"large_{name}.npz"), **out)
    return f"large {name}: G={p.G} K={p.K} mM={p.mM} mW={p.mW} dN={deriv[0]!r} dC00={deriv[1]!r} ({out['seconds_byref']:.0f} s)"


def job_extras(args):
    """SURVEY 8f rank 3, pinned to the reference: RungeKutta4thSolver (Solver.cpp:597-654, never instantiated by the
    reference's experiments) and the Zero / Alpha extrapolation types of C (Solver.cpp:361-389; Alpha with the integer
    division of S:377).  Unmodified build; exp 14.4 b and c, 40 steps each."""
    name, sw, sm = args
    ref = Reference()
    p = ref.from_numbers(sw, sm, 0.4, 0.2, 0.001, 0)
    out = dict(G=p.G, sigmaW=sw, sigmaM=sm, b=0.4, d=0.2, dd=0.001)
    base = np.load(os.path.join(HERE, f"exp_14_4_{name}.npz"))
    mid = base["state_k100"]
    t, v, d = p.solve(4.0, 0.1, want_derivs=True, rk4=True)
    out.update(rk4_times=t, rk4_values=v, rk4_derivs=d)
    for kind, label in ((0, "zero"), (2, "alpha"), (3, "exponential")):
        p.set_extrapolation(kind)
        out[f"{label}_deriv_k0"] = p.rhs(np.concatenate([[200.0], np.full(p.P, 40000.0)]))
        out[f"{label}_deriv_k100"] = p.rhs(mid)
        t, v, d = p.solve(4.0, 0.1, init=mid, want_derivs=True)
        out[f"{label}_times"], out[f"{label}_values"], out[f"{label}_derivs"] = t, v, d
        if label == "zero":
            t, v, d = p.solve(2.0, 0.1, init=mid, want_derivs=True, rk4=True)
            out["zero_rk4_values"] = v
    p.set_extrapolation(1)
    np.savez_compressed(os.path.join(HERE, f"extras_14_4_{name}.npz"), **out)
    return f"extras {name}: rk4 points={len(out['rk4_times'])} N_end={out['rk4_values'][-1, 0]!r} zero N_end={out['zero_values'][-1, 0]!r} alpha N_end={out['alpha_values'][-1, 0]!r}"


MULTI = {
    # the reference's own "(UNSUPPORTED)" example inputs (Data/Custom/): name -> (file, species)
    "2_species": ("(UNSUPPORTED)example_2_species.csv", 2),
    "3_species": ("(UNSUPPORTED)example_3_species.csv", 3),
}


def job_multi(name):
    """SURVEY 8f rank 4, pinned to the reference: Equation and EulerSolver with SpeciesNumber > 1 (Solver.cpp:265-295,
    446-511) on the example inputs the reference ships as "(UNSUPPORTED)" -- its FileReader, Properties::Generate and
    Equation do handle them.  Unmodified build."""
    fname, s = MULTI[name]
    ref = Reference()
    p = ref.from_file(os.path.join(REF_DATA, "Custom", fname), 0)
    assert p.species == s
    g, k = p.grids()
    mM, mW, M, W = p.multi_tables()
    P = p.P
    n0 = 200.0 - 25.0 * np.arange(s)                                     # distinct densities per species
    state = np.empty(p.width)
    state[:s] = n0
    r2 = g[:, None] ** 2 + g[None, :] ** 2
    for i in range(s):
        for j in range(s):
            bump = 1 + (0.3 + 0.05 * i - 0.04 * j) * np.exp(-r2 / (2 * 0.05 ** 2)) + 0.02 * (i - j) * np.sin(9 * g[:, None] + 4 * g[None, :])
            state[s + (i * s + j) * P: s + (i * s + j + 1) * P] = (n0[i] * n0[j] * bump).ravel()
    out = dict(species=s, G=p.G, K=p.K, mM=mM, mW=mW, moments_grid=g, kernels_grid=k, M=M, W=W, state=state,
               deriv=p.rhs(state), deriv_equilibrium=p.rhs(np.concatenate([np.full(s, 200.0), np.full(s * s * P, 40000.0)])))
    t, v, d = p.solve(3.05, 0.1, init=state, want_derivs=True)
    out.update(times=t, values=v, derivs=d)
    t, v, d = p.solve(1.0, 0.1, init=state, want_derivs=True, rk4=True)      # RungeKutta4thSolver on the same equations
    out.update(rk4_times=t, rk4_values=v, rk4_last_deriv=d[-1])
    p.set_extrapolation(0)
    out["zero_deriv"] = p.rhs(state)
    p.set_extrapolation(1)
    if s == 2:
        # what Custom_Experiment (Solver.cpp:1073-1103) plots when this file is its Data/Custom/experiment.csv: Values[k][0]
        # and Values[k][G] over CalculateFirstMoment(100) from the uniform state (by-reference build: bit-identical, 30x faster)
        fast = Reference(byref=True).from_file(os.path.join(REF_DATA, "Custom", fname), 0)
        t, v, _ = fast.solve(100.0, 0.1)
        out["custom_times"], out["custom_columns"] = t, v[:, [0, p.G]]
    np.savez_compressed(os.path.join(HERE, f"multi_{name}.npz"), **out)
    return f"multi {name}: G={p.G} K={p.K} mM={list(mM)} points={len(t)} N_end={list(v[-1, :s])}"


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--visdata":
        make_visdata()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "--sweep":
        make_sweep(int(sys.argv[2]) if len(sys.argv) > 2 else 6)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "--multi":
        for name in MULTI:
            print(job_multi(name), flush=True)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "--extras":
        which = sys.argv[2:] or ["b", "c"]
        cases = {"b": ("b", 0.05, 0.05), "c": ("c", 0.015, 0.15)}
        with ProcessPoolExecutor(max_workers=2) as pool:
            for line in pool.map(job_extras, [cases[w] for w in which]):
                print(line, flush=True)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "--large":
        with ProcessPoolExecutor(max_workers=2) as pool:
            for line in pool.map(job_large, [("g128", 0.15, 0.15, 128), ("g512", 0.10062893578616353, 0.08, 512)]):
                print(line, flush=True)
        return
    assert Reference().available and Reference(True).available, "run `make -C oracle ref` first"
    jobs = [(job_14_6, (w, i)) for w, idx in SWEEP.items() for i in idx]
    jobs += [(job_14_4, c) for c in "acb"]
    jobs += [(job_synth, ("g32", 0.15, 0.15, 32, 0.4, 0.2)), (job_synth, ("g64", 0.1, 0.1, 64, 0.45, 0.25)),
             (job_synth, ("g48_narrow_w", 0.012, 0.1, 48, 0.4, 0.2))]
    summary = []
    with ProcessPoolExecutor(max_workers=os.cpu_count()) as pool:
        futures = [pool.submit(fn, arg) for fn, arg in jobs]
        for f in futures:
            line = f.result()
            print(line, flush=True)
            summary.append(line)
    with open(os.path.join(HERE, "SUMMARY.json"), "w") as fh:
        json.dump({"generated_by": "tests/golden/make_golden.py", "source": "oracle/_ref (reference compiled from /root/reference)",
                   "lines": summary}, fh, indent=1)


if __name__ == "__main__":
    main()
