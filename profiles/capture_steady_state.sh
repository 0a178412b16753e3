# Steady-state capture of the persistent kernel: the 200-step launch of the default bench (the 20-step warm-up launch is
# skipped), so the one-off table build no longer weighs on the samples as it did in the 5-step capture of r01b.
set -x
python bench.py --steps 200 --warmup 20 --no-cpu-baseline --batch-sets 0 --no-sweep --sharded none > gpurun_out/plain_r1f.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:persistent_scan_kernel -s 1 -c 1 -o gpurun_out/prof_r1f_persistent_scan_200 -f \
    python bench.py --steps 200 --warmup 20 --no-cpu-baseline --batch-sets 0 --no-sweep --sharded none > gpurun_out/ncu_r1f.log 2>&1
tail -3 gpurun_out/ncu_r1f.log
ls -la gpurun_out | grep r1f
