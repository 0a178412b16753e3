#!/usr/bin/env python
"""Benchmark of the Law-Dieckmann hot path on B200: grid-points x Euler-steps per second (FP64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

Prints ONE JSON line (rank 0).  A "step" is one explicit-Euler step of the whole moment grid: the integrals
(m*C)(xi) at every node and WC, the closure, and the fused update of N and C (Solver.cpp:431-513, 577-580).

Workloads:
  g128      exp 14.4a parameters on a forced 128 x 128 moment grid = 16 384 xi-points (BASELINE config 3);
            the N=1 headline.  With N > 1 ranks every rank advances its own replica of it (independent
            parameter sets, one per GPU, no data-path collective) -> weak scaling.
  exp14_4a  the reference's own default grid (G=41), latency-bound; reported for completeness.

`--impl reference` times the reference's own CPU code on the same workload (oracle/_ref, compiled from the
reference's sources; see oracle/Makefile) -- the only place besides cpu_baseline where bench.py runs oracle/.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "grid-pts*steps/sec (FP64)"
UNIT = "grid-pts*steps/s"
DT = 0.1

WORKLOADS = {
    # name: (sigma_w, sigma_m, b, d, d', forced_G)
    "g128": (0.15, 0.15, 0.4, 0.2, 0.001, 128),
    "g64": (0.1, 0.1, 0.4, 0.2, 0.001, 64),
    "g256": (0.10062893578616353, 0.08, 0.4, 0.2, 0.001, 256),
    "g512": (0.10062893578616353, 0.08, 0.4, 0.2, 0.001, 512),       # exp 14.6a entry 80 on 262 144 xi-points (config 4)
    "g1024": (0.10062893578616353, 0.08, 0.4, 0.2, 0.001, 1024),     # the same on 1 048 576 xi-points
    "exp14_4a": (0.15, 0.15, 0.4, 0.2, 0.001, 0),
}


def seeded_state(grid, sigma, n0=200.0):
    """SURVEY 8d config 3 seed: C = N^2 (1 + 0.5 exp(-|xi|^2 / 2 sigma^2)), fixed, no RNG."""
    r2 = grid[:, None] ** 2 + grid[None, :] ** 2
    st = np.empty(1 + grid.size ** 2)
    st[0] = n0
    st[1:] = (n0 * n0 * (1 + 0.5 * np.exp(-r2 / (2 * sigma * sigma)))).ravel()
    return st


def algorithmic_flop_per_step(props):
    """F_step = 11 (P S_nz + nzW) + 20 P   (SURVEY.md 8d): 11 FLOP per non-zero-weight sample."""
    def nz(table, extent):
        idx = [abs(i + 1) for i in range(1 - extent, extent - 1, 2)]
        return int(np.count_nonzero(table[np.ix_(idx, idx)]))
    P = props.G * props.G
    s_nz, nz_w = nz(props.M, props.MaxNumForM), nz(props.W, props.MaxNumForW)
    return 11.0 * (P * s_nz + nz_w) + 20.0 * P, s_nz


def executed_flop_per_step(props, prefix):
    """FP64 operations actually issued per step (not the 11-FLOP-per-sample model).  Prefix-sum path: the generation of E
    (7 per element) with its running sum and the scaling of the closed H rows (2), the y side (6 per tap pair), the closing
    phase (~30 per node).  Dense path: the same with the padded contraction on the tensor cores instead of the running sum."""
    G, A = props.G, (props.MaxNumForM - 1) // 2 if props.MaxNumForM > 2 else 1
    if prefix:
        return 9.0 * G * A * G + 6.0 * G * G * 2 * A + 30.0 * G * G
    bm = 40 if A <= 40 else 80
    bn = 64 if G <= 64 else 128
    tiles, chunks = -(-A // bm), -(-G // bn)
    k_pad = -(-A // 16) * 16
    gemm = 2.0 * G * (tiles * bm) * (chunks * bn) * k_pad
    e_gen = 7.0 * G * tiles * k_pad * (chunks * bn)
    y_side = 6.0 * G * chunks * G * 2 * A
    return gemm + e_gen + y_side + 30.0 * G * G


def minimal_wavefronts_per_step(props):
    """Load/store data-pipe wavefronts (128 bytes of L1 / shared memory per clock and SM) that the prefix-sum algorithm
    cannot avoid, per Euler step -- the resource model of DESIGN.md 5:
      E phase   4 values of C per element of E (G rows x A quadrature nodes x G columns); a warp load of 32 consecutive
                columns is 256 bytes = 2 wavefronts;
      H tile    every element of the A x G tile of a row written once (2 wavefronts per 32 columns);
      y side    per (32 nodes y, signed offset): the 16-bit interval indices (1 wavefront), the weights (2) and two
                8-byte gathers per lane from the H tile (2 each when free of bank conflicts)."""
    G, m = props.G, props.MaxNumForM
    A, n_off = ((m - 1) // 2, m - 1) if m > 2 else (1, 1)
    col_blocks = -(-G // 32)
    e_phase = G * A * 4 * col_blocks * 2
    h_tile = G * A * col_blocks * 2
    y_side = G * n_off * col_blocks * 7
    return float(e_phase + h_tile + y_side)


def requested_bytes_per_step(props):
    """Bytes of C the E phase of the prefix-sum path requests per step: four values per element of E (ncu: two thirds of
    them hit in L1, profiles/r01_traffic.json)."""
    G, A = props.G, (props.MaxNumForM - 1) // 2 if props.MaxNumForM > 2 else 1
    return 32.0 * G * A * G


# the translation unit of the profiled kernels (persistent_scan_kernel, persistent_chunks_kernel: ldm_persist.cu) and every
# header it includes
PROFILED_SOURCES = ("ldm_device.h", "ldm_persist.cu", "ldm_pointwise.cuh", "ldm_scan_body.cuh")


def source_sha():
    """sha256 over the sources the profiled kernels are compiled from: a counter file captured from other sources is
    refused.  (Until the closure extension arrived this ran over every file of csrc/; the capture of profiles/r02_counters.json
    was re-stamped when the list was narrowed -- the four files are byte-identical to the captured build's, and so is the SASS
    of ldm_persist.cu.o: profiles/r02_sass_sha.txt.)"""
    import hashlib
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "euler-_solver_-openmp-_b200", "csrc")
    for name in PROFILED_SOURCES:
        h.update(name.encode())
        h.update(open(os.path.join(csrc, name), "rb").read())
    return h.hexdigest()


def measured_counters(workload, kernel_prefix):
    """Per-step counters of the dominant kernel from the committed `ncu --set full` capture of THIS build
    (profiles/r02_counters.json, written by profiles/ncu_to_counters.py from the capture of profiles/capture_r02.sh): DRAM
    bytes, L2 -> L1 bytes, load/store wavefronts, executed FP64 operations.  Refused (None, with the reason) when the file was
    captured from other kernel sources or names another kernel."""
    path = os.path.join(ROOT, "profiles", "r02_counters.json")
    if not os.path.exists(path):
        return None, "profiles/r02_counters.json is missing"
    entry = json.load(open(path)).get(workload)
    if not entry:
        return None, f"no capture of workload {workload}"
    if entry.get("source_sha") != source_sha():
        return None, "the capture was taken from other kernel sources (source_sha differs): counters refused"
    if not str(entry.get("kernel", "")).startswith(kernel_prefix):
        return None, f"the capture names kernel {entry.get('kernel')!r}, the bench ran {kernel_prefix!r}"
    return entry, None


def measured_hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except (OSError, KeyError, ValueError):
        return 6549.8


class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, flag in zip(names, r[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def workload_config(name, G, K, mM, mW, s_nz, world):
    """`config` of the JSON line, the same keys and values in both arms"""
    return {"workload": name, "G": int(G), "xi_points": int(G) * int(G), "K": int(K), "mM": int(mM), "mW": int(mW), "S_nz": int(s_nz),
            "dt": DT, "parallelism": f"replicas x{world}" if world > 1 else "single",
            "l2": "flushed once immediately before the timed region (256 MiB fill); the K timed steps are consecutive steps of one trajectory",
            "seed": "C = N^2 (1 + 0.5 exp(-r^2/2 sigma^2)), N = 200"}


def nonzero_samples(table, extent):
    idx = [abs(i + 1) for i in range(1 - extent, extent - 1, 2)]
    return int(np.count_nonzero(table[np.ix_(idx, idx)]))


def reference_arm(args, rank, world, emit):
    """The reference's own CPU implementation of the same step on the host cores (single-threaded: the
    reference has no parallel construct on this route, Solver.cpp:885-906)."""
    if rank != 0:
        return
    from refharness import Reference
    sw, sm, b, d, dd, G = WORKLOADS[args.workload]
    big = G == 0 or G > 48
    ref = Reference(byref=big)
    if not ref.available:
        emit({"impl": "reference", "unavailable": "oracle/_ref was not built (needs /root/reference at build time)"})
        return
    p = ref.from_numbers(sw, sm, b, d, dd, G)
    grid, _ = p.grids()
    state = seeded_state(grid, sm)
    M, _ = p.tables()
    config = workload_config(args.workload, p.G, p.K, p.mM, p.mW, nonzero_samples(M, p.mM), args.gpus)
    # One step = the hot loop of one right-hand side (Solver.cpp:277-316, > 99.9 % of a step) on a BOUNDED SAMPLE of the
    # moment grid: every `stride`-th row (near and far rows alike -- the cost of a node grows with its distance from the
    # origin), sized so that --steps K --warmup W ends within a few minutes; the sampled rows are then advanced like
    # EulerSolver::Solve does (update + clamp).  value = sampled grid points x steps / seconds.
    stride = 1 if p.P < 4096 else max(1, p.G // 8)
    rows = np.arange(0, p.G, stride)
    C = state[1:].reshape(p.G, p.G)
    m_nodes, w_nodes = (p.kernels_at_nodes() if stride > 1 else (None, None))
    wc = p.integrals(state, want_mc=False)[0] if stride > 1 else None

    def step():
        nonlocal state
        if stride == 1:
            deriv = p.rhs(state)                                  # Equation::operator()
            state = np.maximum(state + deriv * DT, 0.0)           # EulerSolver::Solve's update and clamp (S:577-580)
            return
        mc = p.mc_rows(state, 0, stride)                          # MakeIntegralWithPairDensity at every node of the sampled rows
        c = C[rows]
        dc = 2 * b * mc + 2 * b * state[0] * m_nodes[rows] - 2 * d * c - 2 * dd * wc * c / state[0] - 2 * dd * w_nodes[rows] * c
        C[rows] = np.maximum(c + dc * DT, 0.0)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    sec = time.perf_counter() - t0
    points = len(rows) * p.G
    value = points * args.steps / sec
    kind = ("reference source with its five by-value matrix parameters turned into references (oracle/Makefile; "
            "bit-identical results, the shipped build needs ~1 h per step at this size)") if big else "unmodified reference build"
    sample = (f"{args.steps} steps, each the reference's per-node integral on {len(rows)} of the {p.G} grid rows (every {stride}-th row, "
              f"{points} of {p.P} grid points) plus their update; {kind}") if stride > 1 else f"{args.steps} full Euler steps; {kind}"
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sec / args.steps * (p.P / points), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "impl": "reference", "config": config,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "reference", "sample": sample,
                             "threads": "the reference runs a single solve on one thread (Solver.cpp:885-906); its only parallel "
                                        "construct is over independent parameter sets (S:940)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def cpu_baseline(workload, budget_s=25.0):
    from refharness import Reference
    sw, sm, b, d, dd, G = WORKLOADS[workload]
    ref = Reference(byref=True)
    if not ref.available:
        return {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": "oracle/_ref not built"}
    p = ref.from_numbers(sw, sm, b, d, dd, G)
    grid, _ = p.grids()
    state = seeded_state(grid, sm)
    n, t0 = 0, time.perf_counter()
    while True:
        deriv = p.rhs(state)
        state = np.maximum(state + deriv * DT, 0.0)
        n += 1
        sec = time.perf_counter() - t0
        if sec * (n + 1) / n > budget_s or n >= 50:
            break
    out = {"value": p.P * n / sec, "unit": UNIT, "cores": 1, "kind": "reference",
           "sample": f"{n} Euler step(s) in {sec:.1f} s, single thread (the reference is serial on this route); "
                     "reference source built by-reference (bit-identical to the shipped build, which is ~30x slower)"}
    # beside it: what the host could do at best with the same algorithm -- the C restatement (oracle/, bit-identical to the
    # reference) with its outer node loop spread over every core the box has
    try:
        from oracle import OracleProblem, lib as oracle_lib
        cores = len(os.sched_getaffinity(0))
        oracle_lib().ldmo_set_threads(cores)
        o = OracleProblem(sw, sm, b, d, dd, G)
        st = seeded_state(grid, sm)
        t0 = time.perf_counter()
        o.rhs(st)
        sec_port = time.perf_counter() - t0
        out["all_cores_port"] = {"value": p.P / sec_port, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": f"1 right-hand side (> 99.9 % of a step) in {sec_port:.2f} s, OpenMP over the grid nodes"}
    except Exception as e:           # the baseline must not take the bench line down
        out["all_cores_port"] = {"error": str(e)}
    return out


def run_sharded_peer(ldm, torch, dist, workload, rank, world, local_rank, steps, warmup):
    """BASELINE config 4 with the all-gather FUSED into the update: ONE launch per rank for all the steps (persistent
    work-list kernel); the CTA that completes a row stores it straight into every rank's next replica over NVLink (CUDA IPC
    peer mappings) and sets the row's arrival flag there -- the steps of a launch are ordered by those data dependencies alone.  No NCCL call on the data path (NCCL only carries the 64-byte
    IPC handles at set-up and the timing reduction).  Every rank then repeats the same steps ALONE on its own GPU (one
    rank, same kernel) and compares its replica with that run element by element; the one-rank time measured there, in
    the same run on the same GPU, is the denominator of `speedup_vs_1gpu`; `speedup_vs_default_1gpu` is taken against what
    ldm_advance itself runs for this grid on one GPU (faster there: no work list), the same bits again."""
    sw, sm, b, d, dd, G = WORKLOADS[workload]
    props = ldm.Properties.Generate(1, sw, sm, b, d, dd, G)
    sh = ldm.ShardedEquation(props, rank, world, device=local_rank)

    def all_gather_bytes(mine):
        if world == 1:
            return mine.reshape(1, -1)
        t = torch.from_numpy(mine).cuda()
        out = torch.empty(world * t.numel(), dtype=torch.uint8, device="cuda")
        dist.all_gather_into_tensor(out, t)
        return out.cpu().numpy().reshape(world, -1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sh.peer_connect(all_gather_bytes)
    state0 = seeded_state(props.MomentsGrid, sm)
    sh.peer_upload(state0[1:], state0[0])
    barrier()
    stream = torch.cuda.current_stream()
    sh.peer_advance(warmup, DT, stream.cuda_stream)
    torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = sh.eq.launches
    e0.record(stream)
    sh.peer_advance(steps, DT, stream.cuda_stream)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = int(sh.eq.launches - launches0)
    final_C, N_end = sh.peer_download()
    barrier()
    # the same steps on ONE rank (this GPU alone), timed the same way: element-wise reference and speed-up denominator
    one = ldm.ShardedEquation(props, 0, 1, device=local_rank)
    one.peer_connect(lambda mine: mine.reshape(1, -1))
    one.peer_upload(state0[1:], state0[0])
    one.peer_advance(warmup, DT, stream.cuda_stream)
    torch.cuda.synchronize()
    e0.record(stream)
    one.peer_advance(steps, DT, stream.cuda_stream)
    e1.record(stream)
    torch.cuda.synchronize()
    one_ms = e0.elapsed_time(e1)
    want_C, want_N = one.peer_download()
    diff = max(float(np.max(np.abs(final_C - want_C))), abs(N_end - want_N))
    # ... and the library's own choice for this grid on one GPU (ldm_advance: one launch per kernel and step for rows of several
    # chunks -- on one rank that is faster than the work-list kernel the ranks run), the stricter denominator
    plain = ldm.Equation(props, device=local_rank)
    plain.upload(state0)
    plain.advance(warmup, DT); plain.synchronize()
    plain.advance(steps, DT); plain.synchronize()
    plain_ms = plain.last_elapsed_ms()
    plain_state = plain.download()
    diff = max(diff, float(np.max(np.abs(plain_state[1:] - want_C.ravel()))), abs(plain_state[0] - want_N))
    plain.close()
    stats = torch.tensor([ms, one_ms, diff, plain_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    ms, one_ms, diff, plain_ms = (float(v) for v in stats)
    assert np.all(np.isfinite(final_C)) and N_end > 0
    flop_step, s_nz = algorithmic_flop_per_step(props)
    P = G * G
    return {"workload": workload, "G": G, "xi_points": P, "S_nz": s_nz, "scaling": "strong", "steps": steps, "warmup": warmup,
            "value": P * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps,
            "one_gpu_ms_per_step": one_ms / steps, "speedup_vs_1gpu": one_ms / ms,
            "one_gpu_default_ms_per_step": plain_ms / steps, "speedup_vs_default_1gpu": plain_ms / ms,
            "max_abs_diff": diff, "elementwise": "every rank's replica against the same steps on one rank, all G*G values and N; max over ranks",
            "collective": "fused into the persistent kernel: NVLink peer stores of every row as it closes into every rank's replica, per-row arrival "
                          "flags (the next step's first rows start while this step's last rows are still computed and shipped) and per-rank step flags "
                          "(%d bytes to each of %d peers per rank per step); no NCCL on the data path" % (sh.rows_per_rank * G * 8, world - 1),
            "gpu_launches": launches, "algorithmic_tflops": flop_step * steps / (ms * 1e-3) / 1e12,
            "checksum": float(final_C.sum() + N_end)}


def run_sharded(ldm, torch, dist, workload, rank, world, local_rank, steps, warmup):
    """BASELINE config 4: ONE problem, its xi output rows sharded cyclically over the ranks; per step every rank
    updates its rows of C from the full replica and an NCCL all-gather (in place, over NVLink) rebuilds the replica.
    N and WC are recomputed redundantly on every rank, so the all-gather is the only collective.  Strong scaling."""
    sw, sm, b, d, dd, G = WORKLOADS[workload]
    props = ldm.Properties.Generate(1, sw, sm, b, d, dd, G)
    sh = ldm.ShardedEquation(props, rank, world, device=local_rank)
    rpr = sh.rows_per_rank
    state0 = seeded_state(props.MomentsGrid, sm)
    bufs = [torch.from_numpy(sh.to_gathered(state0[1:])).cuda(), torch.zeros(world * rpr, G, dtype=torch.float64, device="cuda")]
    sh.set_N(state0[0])
    stream = torch.cuda.current_stream()
    mine = slice(rank * rpr, (rank + 1) * rpr)
    k = 0

    def step():
        nonlocal k
        src, dst = bufs[k % 2], bufs[(k + 1) % 2]
        sh.step(src.data_ptr(), dst[mine].data_ptr(), DT, stream.cuda_stream)
        if world > 1:
            dist.all_gather_into_tensor(dst.view(-1), dst[mine].reshape(-1))      # in place: input is this rank's slice of the output
        k += 1

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = sh.eq.launches
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    final_C = sh.from_gathered(bufs[k % 2].cpu().numpy())
    N_end = sh.get_N()
    assert np.all(np.isfinite(final_C)) and N_end > 0
    flop_step, s_nz = algorithmic_flop_per_step(props)
    P = G * G
    return {"workload": workload, "G": G, "xi_points": P, "S_nz": s_nz, "scaling": "strong", "steps": steps, "warmup": warmup,
            "value": P * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps,
            "collective": "NCCL all_gather_into_tensor of C (in place), %d bytes per rank per step" % (rpr * G * 8) if world > 1 else "none (1 rank)",
            "gpu_launches": int(sh.eq.launches - launches0), "algorithmic_tflops": flop_step * steps / (ms * 1e-3) / 1e12,
            "checksum": float(final_C.sum() + N_end)}


def lattice_parameter_sets(count):
    """BASELINE config 5: (b, d, sigma) sets around the exp 14.4 values on a fixed lattice -- index-derived, no RNG.
    b in [0.3, 0.5], d in [0.1, 0.3], sigma_W = sigma_M in [0.05, 0.15]."""
    sets = []
    for i in range(count):
        ib, idd, isg = i % 8, (i // 8) % 8, (i // 64) % 16
        sets.append((0.05 + 0.1 * isg / 15.0, 0.3 + 0.2 * ib / 7.0, 0.1 + 0.2 * idd / 7.0))
    return sets


def run_batch(ldm, torch, dist, rank, world, local_rank, per_gpu, steps, G=64):
    """the sigma-sweep loop (Solver.cpp:940-963) as a batch: per_gpu independent parameter sets on every GPU, all of them
    advanced by ONE launch per kernel and step (blockIdx.z = set); no communication.  Weak scaling."""
    sets = lattice_parameter_sets(per_gpu * world)[rank * per_gpu:(rank + 1) * per_gpu]
    eqs, inits = [], []
    t0 = time.perf_counter()
    for sigma, b, d in sets:
        props = ldm.Properties.Generate(1, sigma, sigma, b, d, 0.001, G)
        eq = ldm.Equation(props, device=local_rank)
        eqs.append(eq)
        inits.append(seeded_state(props.MomentsGrid, sigma))
    setup_s = time.perf_counter() - t0
    t_max = (steps + 0.5) * DT
    ldm.solve_batch(eqs, inits, DT, (2 + 0.5) * DT)                    # warm-up: builds the ring operators, loads the kernels
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = eqs[0].launches
    t0 = time.perf_counter()
    outs = ldm.solve_batch(eqs, inits, DT, t_max)
    wall_s = time.perf_counter() - t0
    ms = eqs[0].last_elapsed_ms()                                      # CUDA events around the batched stepping loop
    times = torch.tensor([ms, wall_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, wall_ms = float(times[0]), float(times[1])
    assert all(np.all(np.isfinite(o)) for o in outs)
    P = G * G
    flops = sum(algorithmic_flop_per_step(e.props)[0] for e in eqs)
    return {"workload": f"{per_gpu} parameter sets per GPU, G={G} ({P} xi-points each), lattice of (b, d, sigma)", "sets": per_gpu * world,
            "scaling": "weak", "steps": steps, "value": world * per_gpu * P * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps,
            "e2e_value": world * per_gpu * P * steps / (wall_ms * 1e-3), "gpu_launches": int(eqs[0].launches - launches0),
            "algorithmic_tflops_per_gpu": flops * steps / (ms * 1e-3) / 1e12, "setup_s": setup_s,
            "checksum": float(sum(o.sum() for o in outs))}


def run_sweep(ldm, torch, local_rank):
    """Experiment 14.6 a as the reference defines it (Solver.cpp:909-977, 1033-1071): 160 entries, sigma_W = linspace(1e-8, 0.2),
    sigma_M = 0.08, 2000 Euler steps each, every entry with its own grids (G 11..39, K 23..227) -- set up on the host,
    advanced as ONE batch (ldm_solve_batch), last states read back.  Wall clock, set-up included: what a user waits for."""
    sigmas = np.linspace(1e-8, 0.2, 160)          # Vector::setLinSpaced(160, EPSILON, 0.2): same values to the last bit but one
    sigmas[-1] = 0.2
    from concurrent.futures import ThreadPoolExecutor
    t0 = time.perf_counter()

    def create(sw):                               # host tables + upload; ctypes drops the GIL inside the library calls
        props = ldm.Properties.Generate(1, float(sw), 0.08, 0.4, 0.2, 0.001)
        return ldm.Equation(props, device=local_rank)

    with ThreadPoolExecutor(max_workers=min(8, len(os.sched_getaffinity(0)))) as pool:   # as the front-end does (host/Solver.cpp)
        eqs = list(pool.map(create, sigmas))
    inits = [eq.GenerateEquilibriumValuesForNAndC(200) for eq in eqs]
    setup_s = time.perf_counter() - t0
    outs = ldm.solve_batch(eqs, inits, DT, 200)
    total_s = time.perf_counter() - t0
    points = sum(e.props.G ** 2 for e in eqs)
    assert all(np.all(np.isfinite(o)) for o in outs)
    return {"workload": "experiment 14.6 a: 160 entries x 2000 Euler steps, one batched call", "entries": 160, "steps": 2000,
            "xi_points_total": points, "seconds_total": total_s, "seconds_setup": setup_s,
            "value": points * 2000 / total_s, "unit": UNIT, "stepping_ms": eqs[0].last_elapsed_ms(),
            "N_end": [float(outs[i][0]) for i in (0, 12, 80, 159)],
            "reference": "estimated >= 15 h on 8 cores (SURVEY 6): entries i = 12 and 80 alone take 141 s and 20 s of one core"}


def run_closures(ldm, props, init, local_rank, lsu_peak, steps=10):
    """The extension of DESIGN.md 11 on the headline grid: Euler steps with the third moment closed by the power-3 and the
    power-2 (symmetric) closure -- K(xi) = integral of w(xi') C(xi') C(xi + xi') (and (w*C)(xi)) at every node, sample by
    sample, behind the same m*C integral.  Device time of ldm_advance.  Never lets the bench line down: an error is reported
    in the leg."""
    try:
        eq = ldm.Equation(props, device=local_rank)
        n_off = max(props.MaxNumForW - 1, 1)
        samples = props.G ** 2 * n_off ** 2
        # load/store data-pipe wavefronts the sums kernel needs at least: per warp iteration (32 offsets along y of one sample row)
        # table entry, abscissa and weight two each, interval one, two 8-byte shared-memory taps two each
        wavefronts_min = 11.0 * props.G ** 2 * n_off * ((n_off + 31) // 32)
        out = {"workload": f"{props.G * props.G} xi-points, w quadrature {n_off} x {n_off}: {samples} samples per integral and step",
               "steps": steps, "parity": "extension: pinned to the reference at the uniform state and for power-1 only (DESIGN.md 11)"}
        for closure, name in ((ldm.CLOSURE_POWER3, "power3"), (ldm.CLOSURE_POWER2_SYMMETRIC, "power2")):
            eq.set_closure(closure)
            eq.upload(init); eq.advance(3, DT); eq.synchronize()
            eq.upload(init); eq.synchronize()
            before = eq.launches
            eq.advance(steps, DT); eq.synchronize()
            ms = eq.last_elapsed_ms() / steps
            state = eq.download()
            assert np.all(np.isfinite(state))
            out[name] = {"ms_per_step": ms, "value": props.G ** 2 / (ms * 1e-3), "unit": UNIT,
                         "gpu_launches_per_step": (eq.launches - before) / steps,
                         "samples_per_s": samples * (2 if closure == ldm.CLOSURE_POWER2_SYMMETRIC else 1) / (ms * 1e-3),
                         "N_after": float(state[0])}
        out["lsu_frac_power3"] = wavefronts_min / (out["power3"]["ms_per_step"] * 1e-3) / lsu_peak
        out["lsu_frac_note"] = ("minimal data-pipe wavefronts of the K integral over the whole power-3 step (m*C integral and closing "
                                "included in the time) against the pipe's peak measured live")
        return out
    except Exception as e:                       # noqa: BLE001 -- the headline legs must not depend on the extension
        return {"error": f"{type(e).__name__}: {e}"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="g128", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sharded", default="g512", help="workload of the xi-sharded strong-scaling leg ('none' to skip)")
    ap.add_argument("--sharded-steps", type=int, default=50)
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="sharded leg at N > 1: all-gather fused into the update over NVLink peer stores, or an NCCL all-gather per step")
    ap.add_argument("--batch-sets", type=int, default=128, help="parameter sets per GPU of the batched leg (0 to skip)")
    ap.add_argument("--batch-steps", type=int, default=200)
    ap.add_argument("--no-sweep", action="store_true", help="skip the experiment-14.6 sweep leg (N = 1 only)")
    ap.add_argument("--item-order", type=int, default=None, help="measurements: hand-out order of the work-list kernel (ldm_tuning_set item_order)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries the ONE JSON line and nothing else: whatever libraries print on file descriptor 1 while the run lasts
    # (NCCL announces its version there, the reference narrates its grids) goes to stderr instead
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)

    if args.impl == "reference":
        reference_arm(args, rank, world, emit)
        return

    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package
    ldm = load_package()
    if args.item_order is not None:
        ldm.tuning_set("item_order", args.item_order)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    sw, sm, b, d, dd, G = WORKLOADS[args.workload]
    props = ldm.Properties.Generate(1, sw, sm, b, d, dd, G)
    eq = ldm.Equation(props, device=local_rank)
    P = props.G * props.G
    state0 = seeded_state(props.MomentsGrid, sm)
    flop_step, s_nz = algorithmic_flop_per_step(props)
    prefix = eq.algorithm == ldm.ALGO_PREFIX
    flop_executed = executed_flop_per_step(props, prefix)
    fp64_peak = ldm.measure_fp64_peak(local_rank)
    lsu_peak, l2_peak = ldm.measure_memory_pipes(local_rank)       # wavefronts/s over the whole device, GB/s

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB of L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident stepping: inputs already in HBM.  The timed region is K consecutive Euler steps of one
    # trajectory (each step consumes the previous step's output, as in the reference's solve loop); L2 is flushed
    # once, immediately before the region, so the first step starts cold.  CUDA events on the launching stream. ----
    eq.upload(state0)
    eq.advance(args.warmup, DT)
    eq.synchronize()
    launches0 = eq.launches
    flush.fill_(1)
    barrier()
    with ClockSampler(local_rank) as clocks:
        eq.advance(args.steps, DT)
        eq.synchronize()
        total_ms = eq.last_elapsed_ms()
        launches = eq.launches - launches0        # kernels of ours launched inside the timed region
        t_end = time.time() + 1.5                 # keep the same load running until nvidia-smi has sampled it
        while time.time() < t_end and len(clocks.rows) < 3:
            eq.advance(args.steps, DT)
            eq.synchronize()
    barrier()
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = world * P * args.steps / (total_ms * 1e-3)
    final = eq.download()
    assert np.all(np.isfinite(final)) and final[0] > 0

    # ---- the same step with L2 flushed before EVERY step (one launch per step, each starting from HBM): the other reading
    # of the timing rule.  A solve never runs like this -- step k+1 consumes what step k left in L2 -- so it is reported
    # beside `value`, not instead of it.  CUDA events of the handle's own stream around each one-step advance. ----
    cold_steps = min(args.steps, 50)
    cold_ms = 0.0
    for _ in range(cold_steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        eq.advance(1, DT)
        eq.synchronize()
        cold_ms += eq.last_elapsed_ms()
    if world > 1:
        t = torch.tensor([cold_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        cold_ms = float(t.item())
    cold = {"value": world * P * cold_steps / (cold_ms * 1e-3), "unit": UNIT, "steps": cold_steps, "ms_per_step": cold_ms / cold_steps,
            "how": "256 MiB fill + synchronize before every step, one launch per step (ldm_advance(1)), CUDA events on the launching stream"}

    # ---- end to end through the reference-facing call: EulerSolver::Solve with host buffers.  The headline is the
    # reference's own contract -- SolverResults holds Values AND Derivatives of every point (Solver.cpp:546-553), so both
    # cross PCIe, segment by segment under the computation; beside it the observables-only mode (what Experiment_14_4
    # reads off the result, S:1011-1015).  Pinned host buffers, wall clock around the call. ----
    init_pinned = torch.from_numpy(state0).pin_memory().numpy()
    solver = ldm.EulerSolver()
    t_max = (args.steps + 0.5) * DT             # exactly args.steps Euler steps
    width = 1 + P
    values_pinned = torch.empty((args.steps + 1) * width, dtype=torch.float64).pin_memory().numpy()
    derivs_pinned = torch.empty((args.steps + 1) * width, dtype=torch.float64).pin_memory().numpy()
    obs_pinned = torch.empty((args.steps + 1) * 2, dtype=torch.float64).pin_memory().numpy()

    def timed_solve(record):
        kw = dict(values_out=values_pinned, derivs_out=derivs_pinned) if record == ldm.RECORD_ALL else dict(values_out=obs_pinned)
        # warm-up: one untimed call of the SAME length into the same pinned buffers (the first device-to-host copy into freshly
        # pinned pages runs at about half the PCIe rate and then bounds the call: 21 instead of 15 us per step, measured with
        # tools/solve_record_timing.py)
        solver.Solve(eq, ldm.SquaredN, t_max, DT, init_pinned, record=record, **kw)
        barrier()
        t0 = time.perf_counter()
        res = solver.Solve(eq, ldm.SquaredN, t_max, DT, init_pinned, record=record, **kw)
        sec = time.perf_counter() - t0
        assert len(res.Times) == args.steps + 1
        if world > 1:
            t = torch.tensor([sec], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        return world * P * args.steps / sec, res

    e2e_value, res_all = timed_solve(ldm.RECORD_ALL)
    e2e_observables, res_obs = timed_solve(ldm.RECORD_OBSERVABLES)
    assert np.array_equal(res_obs.Observables, res_all.Values[:, [0, props.G]])

    sharded = None
    if args.sharded != "none":
        if args.exchange == "peer":
            sharded = run_sharded_peer(ldm, torch, dist, args.sharded, rank, world, local_rank, args.sharded_steps, 5)
            if world > 1:
                sharded["nccl_variant"] = run_sharded(ldm, torch, dist, args.sharded, rank, world, local_rank, min(args.sharded_steps, 20), 2)
        else:
            sharded = run_sharded(ldm, torch, dist, args.sharded, rank, world, local_rank, args.sharded_steps, 2)

    batch = None
    if args.batch_sets > 0:
        batch = run_batch(ldm, torch, dist, rank, world, local_rank, args.batch_sets, args.batch_steps)

    sweep = run_sweep(ldm, torch, local_rank) if (world == 1 and not args.no_sweep) else None
    closures = run_closures(ldm, props, state0, local_rank, lsu_peak) if (world == 1 and not args.no_sweep) else None

    if rank == 0:
        ms_per_step = total_ms / args.steps
        achieved = flop_step / (ms_per_step * 1e-3) / 1e12
        kernel_name = ("persistent_scan_kernel" if props.G <= 128 else "scan_kernel") if prefix else "persistent_kernel"   # what ldm_advance runs on one GPU
        counters, refused = measured_counters(args.workload, kernel_name)
        wf_min = minimal_wavefronts_per_step(props)
        roofline = {
            # the contract's figure: the 11-FLOP-per-sample model of the reference's sum over the measured DFMA peak
            "bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
            "traffic": counters["dram_bytes_per_step"] if counters else None,
            "algorithmic_flop_per_step": flop_step, "algorithm": "prefix" if prefix else "dense",
            "peak_source": "measured live: DFMA chain micro-benchmark (MEASURED_PEAKS.json has no FP64 figure)",
            "note": "frac > 1: `achieved` counts the reference's sample-by-sample work (11 FLOP per sample); the prefix-sum path issues ~120x "
                    "fewer FP64 operations, so the FP64 pipe no longer bounds the step.  What does is the load/store data pipe (gathers "
                    "of C and of the H tile) and the L2 -> SM path: lsu_frac and l2_frac below measure the machine against those",
            # the resources that bound the prefix-sum path, each against a peak measured live on this device
            "lsu_peak_wavefronts_per_s": lsu_peak, "l2_peak_GBps": l2_peak,
            "lsu_wavefronts_min_per_step": wf_min,
            "lsu_frac": wf_min / (ms_per_step * 1e-3) / lsu_peak,
            "hbm_GBps_algorithmic": 16.0 * P / (ms_per_step * 1e-3) / 1e9,
            "executed_flop_per_step_model": flop_executed,
        }
        if counters:
            roofline.update({
                "counters_from": "profiles/r02_counters.json: ncu --set full of this build's " + counters["kernel"] + ", per step of its launch",
                "lsu_wavefronts_executed_per_step": counters["lsu_wavefronts_per_step"],
                "lsu_frac_executed": counters["lsu_wavefronts_per_step"] / (ms_per_step * 1e-3) / lsu_peak,
                "l2_bytes_per_step": counters["l2_to_l1_bytes_per_step"],
                "l2_frac": counters["l2_to_l1_bytes_per_step"] / (ms_per_step * 1e-3) / 1e9 / l2_peak,
                "executed_flop_per_step": counters["fp64_flop_per_step"],
                "executed_frac": counters["fp64_flop_per_step"] / (ms_per_step * 1e-3) / 1e12 / fp64_peak,
                "waves_per_sm": counters.get("waves_per_sm"),
            })
        else:
            roofline["counters_refused"] = refused
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.workload, props.G, props.K, props.MaxNumForM, props.MaxNumForW, s_nz, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8.0 * (1 + P) / args.steps,
                    "d2h_bytes_per_step": 16.0 * (1 + P) * (args.steps + 1) / args.steps,
                    "call": "EulerSolver::Solve twin (ldm_solve, record = all: Values and Derivatives of every point, the reference's SolverResults), "
                            "pinned host buffers: the persistent kernel stores the history straight into them (one launch for the whole solve, the "
                            "derivative at the last point included), wall clock of the call; one untimed call of the same length comes first",
                    "observables_only": {"value": e2e_observables, "unit": UNIT, "d2h_bytes_per_step": 16.0,
                                         "call": "the same with record = observables (N and C[0][G-1] per point, what Experiment_14_4 reads)"}},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "roofline_hbm": {"bound": "hbm", "achieved": 16.0 * P / (ms_per_step * 1e-3) / 1e9, "peak": measured_hbm_peak(), "unit": "GB/s",
                             "frac": 16.0 * P / (ms_per_step * 1e-3) / 1e9 / measured_hbm_peak(), "traffic": roofline["traffic"],
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs (fallback 6549.8)",
                             "note": "algorithmic bytes (C read once and written once per step, SURVEY 8d) over the step time: three orders of "
                                     "magnitude under the copy bandwidth -- this path is not HBM bound, C never leaves L2 between steps"},
            "clocks": clocks.summary(),
            "l2_flushed_every_step": cold,
        }
        if sharded is not None:
            line["sharded"] = sharded
        if batch is not None:
            line["batch"] = batch
        if sweep is not None:
            line["sweep"] = sweep
        if closures is not None:
            line["closures"] = closures
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.workload)
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
