# GPU check of the closure extension (run under gpurun): its tests, its step times, then the whole GPU suite and one bench line.
TAG=${1:-r02n}
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity_closure.py -q -m gpu > gpurun_out/${TAG}_pytest_closure.log 2>&1; echo "closure pytest exit $?" | tee -a gpurun_out/${TAG}_pytest_closure.log
tail -40 gpurun_out/${TAG}_pytest_closure.log | cut -c1-400
timeout 120 python euler-_solver_-openmp-_b200/tools/closure_step_timing.py > gpurun_out/${TAG}_closure_timing.txt 2>&1; echo "timing exit $?"; cat gpurun_out/${TAG}_closure_timing.txt | cut -c1-300
timeout 240 python -m pytest tests -q -m gpu > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/${TAG}_pytest_gpu.log
tail -6 gpurun_out/${TAG}_pytest_gpu.log | cut -c1-300
timeout 240 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench exit $?"; tail -3 gpurun_out/${TAG}_bench_n1.err
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench_n1.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"])
print({k: v for k, v in d["roofline"].items() if "frac" in k or "counters" in k})
PY
