#!/bin/bash
# Round-end check on one B200: the GPU tests, smoke(), and both arms of bench.py, each timed.
mkdir -p gpurun_out
S=$(date +%s); timeout 420 python -m pytest tests -x -q -m gpu 2>&1 | tail -4; echo "tests $(( $(date +%s)-S )) s"
S=$(date +%s); python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2; echo "smoke $(( $(date +%s)-S )) s"
S=$(date +%s); python bench.py > gpurun_out/bench_r1f_n1.json 2> gpurun_out/bench_r1f_n1.err; echo "bench rc=$? $(( $(date +%s)-S )) s"
S=$(date +%s); python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1f_ref.json 2>&1; echo "ref rc=$? $(( $(date +%s)-S )) s"
cut -c1-300 gpurun_out/bench_r1f_ref.json
