# Sharded leg only (run under gpurun --gpus N): bench under torchrun with the other legs off; extra args go to bench.py.
N=${1:-8}
TAG=${2:-r02}
shift 2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --no-cpu-baseline --batch-sets 0 --no-sweep --steps 50 --warmup 5 "$@" \
    > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err; echo "bench exit $?"; tail -2 gpurun_out/${TAG}_bench_n${N}.err
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench_n${N}.json"))
s = d["sharded"]
print({k: s[k] for k in ("ms_per_step", "one_gpu_ms_per_step", "speedup_vs_1gpu", "max_abs_diff", "gpu_launches", "checksum")})
if "nccl_variant" in s: print("nccl", s["nccl_variant"]["ms_per_step"])
PY
