"""Host side of the product (no GPU): model setup is bit-identical to the reference's Properties::Generate,
the C ABI exports every symbol of include/ldm.h, and compute calls fail loudly without a device."""
import ctypes
import glob
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = sorted(p for p in glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")) if not os.path.basename(p).startswith(("large_", "extras_", "multi_", "exp_14_6_sweep")))


def generate(ldm, g, name):
    forced = int(g["G"]) if "synth" in name else 0
    return ldm.Properties.Generate(1, float(g["sigmaW"]), float(g["sigmaM"]), float(g["b"]), float(g["d"]),
                                   float(g["dd"]), forced)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_generate_matches_reference_bitwise(ldm, path):
    g = np.load(path)
    p = generate(ldm, g, path)
    assert (p.G, p.K, p.MaxNumForM, p.MaxNumForW) == (int(g["G"]), int(g["K"]), int(g["mM"]), int(g["mW"]))
    assert p.KernelsGridStep == float(g["h"]) and p.CutoffDistance == float(g["cutoff"])
    assert np.array_equal(p.MomentsGrid, g["moments_grid"])
    assert np.array_equal(p.KernelsGrid, g["kernels_grid"])
    assert np.array_equal(p.M, g["M"]) and np.array_equal(p.W, g["W"])
    p.CheckKernels()       # Solver.cpp:147-183: both integrals within 1e-8 of 1


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_reference_kernel_tables_are_separable_inside_their_disc(ldm, path):
    """the structure the prefix-sum path rests on: every table Properties::GenerateKernel builds (Solver.cpp:98-145) is
    q q^T on a staircase region, up to the rounding of libm's exp (a few 1e-15), and zero beyond"""
    g = np.load(path)
    p = generate(ldm, g, path)
    separable, deviation = p.KernelStructure()
    assert separable and deviation < 1e-14
    # the check is on the table that is passed in, not on how it was made: a perturbed entry is structure, not rounding
    table = np.ctypeslib.as_array(p._ptr.contents.M, (p.K, p.K))
    if table[3, 3] != 0:         # at least a full 2 x 2 block inside the disc: four entries against two parameters
        table[1, 1] *= 1 + 1e-9
        separable, deviation = p.KernelStructure()
        assert not separable and deviation > 1e-10
        table[1, 1] /= 1 + 1e-9
        table[3, 1] *= 1 + 1e-9          # an asymmetric table is not separable either
        assert not p.KernelStructure()[0]


@pytest.mark.parametrize("name", ["2_species", "3_species"])
def test_generate_several_species_matches_reference_bitwise(ldm, name):
    """Properties::Generate with num > 1 (Solver.cpp:186-241) on the reference's own multi-species example inputs: grids
    from the largest sigmas, one W table per pair, one M table per species -- bit for bit"""
    g = np.load(os.path.join(ROOT, "tests", "golden", f"multi_{name}.npz"))
    s = int(g["species"])
    sw = {2: [[0.05, 0.06], [0.06, 0.05]], 3: [[0.05, 0.06, 0.07], [0.06, 0.05, 0.04], [0.07, 0.04, 0.05]]}[s]
    p = ldm.Properties.Generate(s, sw, [0.05] * s, [0.4, 0.401, 0.4001][:s], [0.2] * s, np.full((s, s), 0.001))
    assert (p.SpeciesNumber, p.G, p.K) == (s, int(g["G"]), int(g["K"]))
    assert np.array_equal(p.MaxNumForM, g["mM"]) and np.array_equal(p.MaxNumForW, g["mW"])
    assert np.array_equal(p.MomentsGrid, g["moments_grid"]) and np.array_equal(p.KernelsGrid, g["kernels_grid"])
    assert np.array_equal(p.M, g["M"]) and np.array_equal(p.W, g["W"])
    assert p.state_size == g["state"].size


def test_header_symbols_are_exported(ldm):
    header = open(os.path.join(ROOT, "include", "ldm.h")).read()
    declared = set(re.findall(r"\b(ldm_[a-z_A-Z0-9]+)\s*\(", header))
    assert len(declared) >= 20
    lib = ctypes.CDLL(ldm.LIBRARY_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/ldm.h but not exported"
    assert declared == set(ldm._SIGNATURES), "python binding and header disagree"


def test_point_count(ldm):
    assert ldm.count_points(100, 0.1) == 1001      # Experiment_14_4 (Solver.cpp:1005)
    assert ldm.count_points(200, 0.1) == 2001      # Experiment_14_6 (Solver.cpp:1043)
    assert ldm.count_points(0.05, 0.1) == 1


def test_bad_arguments_are_reported(ldm):
    with pytest.raises(ldm.LdmError):
        ldm.Properties.Generate(5, np.full((5, 5), 0.1), [0.1] * 5, [0.4] * 5, [0.2] * 5, np.full((5, 5), 0.001))   # > LDM_MAX_SPECIES
    with pytest.raises(ldm.LdmError):
        ldm.Properties.Generate(1, -1.0, 0.1, 0.4, 0.2, 0.001)


def test_no_cpu_fallback(ldm):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    p = ldm.Properties.Generate(1, 0.05, 0.05, 0.4, 0.2, 0.001)
    with pytest.raises(ldm.LdmError, match="no CUDA device"):
        ldm.Equation(p)


def test_bench_work_model_matches_the_survey(ldm):
    """bench.py's roofline numerator: S_nz (offsets with non-zero m weight) and F_step = 11 (P S_nz + nzW) + 20 P as
    SURVEY 8d counts them on the reference's own grids (1708 of 2116 at 14.4a, 284 of 400 at b, 648 of 784 at c) and on
    the headline grid (S_m = 23 104 at G = 128)."""
    import sys
    sys.path.insert(0, ROOT)
    import bench
    for (sw, sm), (s_m, s_nz) in {(0.15, 0.15): (2116, 1708), (0.05, 0.05): (400, 284), (0.015, 0.15): (784, 648)}.items():
        p = ldm.Properties.Generate(1, sw, sm, 0.4, 0.2, 0.001)
        assert (p.MaxNumForM - 1) ** 2 == s_m
        flop, got = bench.algorithmic_flop_per_step(p)
        assert got == s_nz
        P = p.G * p.G
        assert flop >= 11.0 * P * s_nz + 20.0 * P and flop <= 11.0 * (P * s_nz + (p.MaxNumForW - 1) ** 2) + 20.0 * P
    p = ldm.Properties.Generate(1, *bench.WORKLOADS["g128"][:5], 128)
    flop, s_nz = bench.algorithmic_flop_per_step(p)
    assert (p.G, p.K, p.MaxNumForM, (p.MaxNumForM - 1) ** 2) == (128, 253, 153, 23104) and s_nz == 17956
    assert abs(flop - 3.2366e9) < 1e6
    sets = bench.lattice_parameter_sets(1024)
    assert len(set(sets)) == 1024 and all(0.05 <= s <= 0.15 + 1e-12 and 0.3 <= b <= 0.5 + 1e-12 and 0.1 <= d <= 0.3 + 1e-12 for s, b, d in sets)


def test_bench_prints_exactly_one_json_line():
    """the driver parses bench.py's stdout: whatever the libraries print on the way (the reference narrates its grids, NCCL
    its version) must not land there.  The reference arm on the reference's own 14.4a grid runs without a device."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "exp14_4a",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = out.stdout.splitlines()
    assert len(lines) == 1, out.stdout[:500]
    line = json.loads(lines[0])
    assert line["impl"] == "reference"
    if "unavailable" not in line:
        assert line["config"]["xi_points"] == 41 * 41 and line["value"] > 0 and line["e2e"]["value"] == line["value"]


def _corrupted(ldm, props, **changes):
    """a copy of the ldm_problem behind `props` with fields replaced; array fields take numpy arrays (kept alive on the copy)"""
    bad = ldm._Problem.from_buffer_copy(props._ptr.contents)
    bad._keep = []
    for name, value in changes.items():
        if isinstance(value, np.ndarray):
            arr = np.ascontiguousarray(value, dtype=np.float64)
            bad._keep.append(arr)
            setattr(bad, name, arr.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        else:
            setattr(bad, name, value)
    return bad


@pytest.mark.parametrize("case", ["mM above K", "mW above K", "even mM", "even K", "grid not increasing", "grid does not start at 0",
                                  "grid does not end at the cutoff", "kernel grid not increasing", "zero step", "mM of 1"])
def test_create_rejects_tables_the_reference_asserts_on(ldm, case):
    """ldm_create accepts any caller-built table, so it checks what the reference asserts (S:77, S:88, S:125-132,
    S:303-304) BEFORE anything indexes with it: LDM_ERR_ARGUMENT (-1), on the host, with or without a GPU"""
    props = ldm.Properties.Generate(1, 0.05, 0.05, 0.4, 0.2, 0.001)
    g, k = props.MomentsGrid.copy(), props.KernelsGrid.copy()
    changes = {
        "mM above K": dict(mM=props.K + 2), "mW above K": dict(mW=props.K + 2), "even mM": dict(mM=4), "mM of 1": dict(mM=1),
        "even K": dict(K=props.K - 1),
        "grid not increasing": dict(moments_grid=np.concatenate([g[:3], g[2:3], g[4:]])),
        "grid does not start at 0": dict(moments_grid=np.concatenate([[1e-3], g[1:]])),
        "grid does not end at the cutoff": dict(moments_grid=np.concatenate([g[:-1], [g[-1] * 1.01]])),
        "kernel grid not increasing": dict(kernels_grid=k[::-1].copy()),
        "zero step": dict(h=0.0),
    }[case]
    bad = _corrupted(ldm, props, **changes)
    handle = ctypes.c_void_p()
    code = ldm.library().ldm_create(ctypes.byref(bad), 0, ctypes.byref(handle))
    assert code == -1 and not handle.value, (case, code, ldm.library().ldm_last_error())
    assert ldm.library().ldm_last_error()          # says what is wrong
    sep, dev = ctypes.c_int(), ctypes.c_double()
    assert ldm.library().ldm_problem_kernel_structure(ctypes.byref(bad), ctypes.byref(sep), ctypes.byref(dev)) == -1


def test_tuning_keys(ldm):
    for key in ("scan_bn", "scan_nt", "factored_bm", "factored_bn", "factored_tn", "persistent", "scan", "chunk_grid", "batch_capacity", "item_order"):
        ldm.tuning_set(key, 0)
    ldm.tuning_set("reset")
    with pytest.raises(ldm.LdmError):
        ldm.tuning_set("no such key", 1)


def test_bench_accepts_the_committed_counters():
    """bench.py's roofline block reads profiles/r02_counters.json only while the sources of the profiled kernels are the
    captured ones (sha256 over ldm_persist.cu and the headers it includes): an edit to any of them without a new capture
    must show up here, not as a silent `counters_refused` in the round's bench line"""
    import sys
    sys.path.insert(0, ROOT)
    import bench
    for workload, kernel in (("g128", "persistent_scan_kernel"), ("g512", "persistent_chunks_kernel")):
        entry, refused = bench.measured_counters(workload, kernel)
        assert refused is None and entry["source_sha"] == bench.source_sha(), refused
        assert entry["lsu_wavefronts_per_step"] > 0 and entry["dram_bytes_per_step"] > 0
    csrc = os.path.join(ROOT, "euler-_solver_-openmp-_b200", "csrc")
    text = open(os.path.join(csrc, "ldm_persist.cu")).read()
    for name in bench.PROFILED_SOURCES:                          # the list IS that translation unit: nothing it includes is missing
        assert os.path.exists(os.path.join(csrc, name))
    includes = {line.split('"')[1] for name in bench.PROFILED_SOURCES for line in open(os.path.join(csrc, name))
                if line.startswith('#include "') and not line.split('"')[1].startswith("..")}
    assert includes <= set(bench.PROFILED_SOURCES), includes
    assert "persistent_scan_kernel" in text and "persistent_chunks_kernel" in text
