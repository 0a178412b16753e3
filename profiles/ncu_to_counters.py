#!/usr/bin/env python
"""Per-step counters of one kernel launch from an `ncu --set full` capture, for bench.py's roofline block.

    ncu -i gpurun_out/X.ncu-rep --page raw --csv | python profiles/ncu_to_counters.py WORKLOAD STEPS_IN_LAUNCH [KERNEL_REGEX]

merges one entry into profiles/r02_counters.json: the kernel's name, the sha256 of the kernel sources the capture was taken
from (bench.py refuses counters whose source_sha differs from the sources it runs), and per Euler step of the launch:
DRAM bytes, L2 -> L1 bytes, load/store data-pipe wavefronts (all / shared memory), executed FP64 operations.
"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def number(text):
    return float(text.replace(",", "")) if text not in ("", "n/a") else 0.0


def main():
    workload, steps = sys.argv[1], int(sys.argv[2])
    pattern = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
    rows = list(csv.reader(sys.stdin))
    head, units = rows[0], rows[1]
    body = [r for r in rows[2:] if len(r) == len(head) and (pattern is None or pattern.search(r[head.index("Kernel Name")]))]
    assert body, "no launch of that kernel in the capture"
    row = body[-1]

    def metric(name, default=None):
        if name not in head:
            assert default is not None, f"{name} is not in the capture"
            return default
        value, unit = number(row[head.index(name)]), units[head.index(name)]
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0, "nsecond": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "ns": 1e-9}.get(unit, 1.0)
        return value * scale

    from bench import source_sha
    dadd = metric("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", 0.0)
    dmul = metric("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", 0.0)
    dfma = metric("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", 0.0)
    l2_sectors = metric("lts__t_sectors_srcunit_tex_op_read.sum", 0.0)
    entry = {
        "kernel": re.sub(r"^.*::", "", row[head.index("Kernel Name")]).split("(")[0],
        "kernel_full_name": row[head.index("Kernel Name")],
        "source_sha": source_sha(),
        "steps_in_launch": steps,
        "launch_seconds": metric("gpu__time_duration.sum"),
        "grid": int(metric("launch__grid_size")), "block": int(metric("launch__block_size")),
        "registers_per_thread": int(metric("launch__registers_per_thread")),
        "waves_per_sm": metric("launch__waves_per_multiprocessor"),
        "sm_cycles_per_step": metric("sm__cycles_elapsed.max") / steps,
        "dram_bytes_per_step": (metric("dram__bytes_read.sum") + metric("dram__bytes_write.sum")) / steps,
        "dram_bytes_per_launch": metric("dram__bytes_read.sum") + metric("dram__bytes_write.sum"),
        "l2_to_l1_bytes_per_step": 32.0 * l2_sectors / steps,
        "lsu_wavefronts_per_step": metric("l1tex__data_pipe_lsu_wavefronts.sum") / steps,
        "lsu_wavefronts_shared_per_step": metric("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum") / steps,
        "global_load_sectors_per_step": metric("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", 0.0) / steps,
        "fp64_flop_per_step": (dadd + dmul + 2.0 * dfma) / steps,
        "warp_instructions_per_step": metric("smsp__inst_executed.sum") / steps,
        "lsu_pipe_pct_avg": metric("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", 0.0),
        "fp64_pipe_pct": metric("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", 0.0),
        "issue_active_pct": metric("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.0),
        "l2_hit_rate_pct": metric("lts__t_sector_hit_rate.pct", 0.0),
        "capture": "ncu --set full --clock-control none (profiles/capture_r02.sh); DRAM bytes are those of the whole launch, cold caches at its start",
    }
    path = os.path.join(ROOT, "profiles", "r02_counters.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    data[workload] = entry
    json.dump(data, open(path, "w"), indent=1)
    print(json.dumps(entry, indent=1))


if __name__ == "__main__":
    main()
