# GPU check of a build (run under gpurun): the GPU tests, smoke(), one bench line.  Logs under gpurun_out/.
TAG=${1:-r02}
python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -15 gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench exit $?"; tail -5 gpurun_out/${TAG}_bench_n1.err
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench_n1.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["e2e"]["observables_only"]["value"])
print({k: v for k, v in d["roofline"].items() if "frac" in k or "peak" in k})
print({k: d["sharded"][k] for k in ("ms_per_step", "one_gpu_ms_per_step", "speedup_vs_1gpu", "max_abs_diff", "gpu_launches")})
print({k: d["batch"][k] for k in ("ms_per_step", "value", "e2e_value")}, d["sweep"]["seconds_total"], d["sweep"]["stepping_ms"])
PY
