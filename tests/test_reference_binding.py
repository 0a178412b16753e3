"""The drop-in, proven on the reference's own program: oracle/_ref/solver_bound is the reference's Solver.cpp -- menu,
FileReader, Experiments, FileWriter, every line of it -- with the four lines of INTEGRATION.md A added
(include/ldm_reference_binding.h, applied by oracle/patch_binding.py, compiled by oracle/Makefile), so that
Equation::operator() (Solver.cpp:431) and EulerSolver::Solve (Solver.cpp:557) run in libldm_b200.so.  Its menu is driven
with scripted keys exactly like the unmodified build's, and the Data\\VisData.csv it writes for experiment 14.4 (b) is
compared with the file the UNMODIFIED reference wrote (tests/golden/VisData_14_4_b_reference.csv, make_golden.py --visdata).
"""
import os
import subprocess

import numpy as np
import pytest

from test_frontend import read_visdata

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BOUND = os.path.join(ROOT, "oracle", "_ref", "solver_bound")
DATA = os.path.join(ROOT, "euler-_solver_-openmp-_b200", "host", "Data")          # byte-identical copies of the reference's inputs
FIXTURE = os.path.join(ROOT, "tests", "golden", "VisData_14_4_b_reference.csv")


def run_menu(exe, keys, workdir, threads=None):
    """the reference's relative paths use backslashes (Solver.cpp:983, 993): on Linux literal file names in the cwd"""
    for root, _, files in os.walk(DATA):
        for f in files:
            if f.endswith(".csv"):
                rel = os.path.relpath(os.path.join(root, f), os.path.dirname(DATA))
                os.symlink(os.path.join(root, f), os.path.join(workdir, rel.replace("/", "\\")))
    env = dict(os.environ, LDM_REF_KEYS=keys, TERM="dumb")
    if threads:
        env["OMP_NUM_THREADS"] = str(threads)
    subprocess.run([exe], cwd=workdir, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True, timeout=900)
    return os.path.join(workdir, "Data\\VisData.csv")


def test_patch_recipe_adds_four_lines():
    """host only: the recipe applies to the reference as mounted in the build container (skipped elsewhere)"""
    ref = "/root/reference/Euler Solver ver.5.3/Solver.cpp"
    if not os.path.exists(ref):
        pytest.skip("the reference is mounted only in the build container")
    patched = subprocess.check_output(["python", os.path.join(ROOT, "oracle", "patch_binding.py"), ref]).decode().split("\n")
    original = open(ref, newline="").read().split("\n")
    assert len(patched) == len(original) + 4
    added = [l for l in patched if "ldm_binding" in l or "LdmProps() {" in l or "ldm_reference_binding" in l]
    assert len(added) == 4 and sorted(set(patched) - set(original)) == sorted(added)


@pytest.mark.gpu
def test_reference_program_runs_on_the_library(tmp_path):
    if not os.path.exists(BOUND):
        pytest.skip("oracle/_ref/solver_bound is built where /root/reference is mounted")
    out = run_menu(BOUND, "80,13,27", str(tmp_path))                       # down, enter, esc: experiment 14.4 (b)
    number, title, labels, data = read_visdata(out)
    want_number, want_title, want_labels, want = read_visdata(FIXTURE)
    assert (number, title, labels) == (want_number, want_title, want_labels) and data.shape == want.shape == (1001, 3)
    assert np.allclose(data, want, rtol=1e-5, atol=0)                      # the file carries 6 significant digits (S:856-858)
    got_lines, want_lines = open(out).read().splitlines(), open(FIXTURE).read().splitlines()
    differing = sum(a != b for a, b in zip(got_lines, want_lines))
    assert differing <= 3, f"{differing} of {len(want_lines)} lines differ textually"      # a value on a 6-digit rounding boundary may flip
    assert got_lines[-1] == "100,89.4668,91.3041"


@pytest.mark.gpu
def test_reference_program_sweep_entry_through_the_library(tmp_path):
    """the sigma sweep (S:909-977) of the reference's own program on the library: 160 x EulerSolver::Solve, each on its own
    grids, results sorted and written by the reference's own code.  Experiment 14.6 (b)."""
    if not os.path.exists(BOUND):
        pytest.skip("oracle/_ref/solver_bound is built where /root/reference is mounted")
    # one OpenMP thread: the reference's loop pushes its results into a shared vector without synchronisation (S:954-956,
    # SURVEY 5); with solves that take milliseconds instead of minutes the threads would collide there
    out = run_menu(BOUND, "80,80,80,80,13,27", str(tmp_path), threads=1)   # fifth menu entry: 14.6 (b)
    number, title, labels, data = read_visdata(out)
    assert number == 14.6 and data.shape == (160, 3)
    g = np.load(os.path.join(ROOT, "tests", "golden", "exp_14_6_sweep.npz"))
    order = np.argsort(g["b_sigma"])
    want = np.stack([g["b_sigma"][order], g["b_N_end"][order], np.sqrt(g["b_C_end"][order])], 1)
    assert np.allclose(data, want, rtol=1e-5, atol=1e-30)
