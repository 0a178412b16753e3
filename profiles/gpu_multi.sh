# Multi-GPU check (run under gpurun --gpus N): the two-process element-wise test, then the bench line under torchrun.
N=${1:-2}
TAG=${2:-r02}
shift 2
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "two_processes or two_devices" > gpurun_out/${TAG}_pytest_n${N}.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_n${N}.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --no-cpu-baseline "$@" \
    > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err; echo "bench exit $?"; tail -3 gpurun_out/${TAG}_bench_n${N}.err
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench_n${N}.json"))
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")})
s = d["sharded"]
print({k: s[k] for k in ("ms_per_step", "one_gpu_ms_per_step", "speedup_vs_1gpu", "max_abs_diff", "gpu_launches", "checksum")})
if "nccl_variant" in s: print("nccl", s["nccl_variant"]["ms_per_step"], s["nccl_variant"]["checksum"])
if d.get("batch"): print({k: d["batch"][k] for k in ("ms_per_step", "value", "e2e_value")})
PY
