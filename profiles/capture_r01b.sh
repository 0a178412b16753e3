set -x
python bench.py --steps 200 --warmup 20 > gpurun_out/bench_r1b_n1.json 2> gpurun_out/bench_r1b_n1.err
tail -c 300 gpurun_out/bench_r1b_n1.err
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --batch-sets 0 --no-sweep > gpurun_out/plain_a.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_prefix_persistent.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --batch-sets 0 --no-sweep > gpurun_out/ncu_a.log 2>&1
LDM_PERSISTENT=0 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --batch-sets 0 --no-sweep --sharded none > gpurun_out/plain_b.log 2>&1 && \
LDM_PERSISTENT=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_prefix_stepwise.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --batch-sets 0 --no-sweep --sharded none > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:persistent_scan_kernel -c 1 -o gpurun_out/prof_r1b_persistent_scan -f python bench.py --steps 6 --warmup 3 --no-cpu-baseline --batch-sets 0 --no-sweep --sharded none > gpurun_out/ncu_c.log 2>&1
ncu --set full --clock-control none --cache-control none --import-source on -k regex:scan_kernel -s 8 -c 1 -o gpurun_out/prof_r1b_scan_g512 -f python bench.py --steps 6 --warmup 3 --no-cpu-baseline --batch-sets 0 --no-sweep --workload g512 --sharded none > gpurun_out/ncu_d.log 2>&1
ls -la gpurun_out | tail -12
(cd euler-_solver_-openmp-_b200 && python tools/scan_clocks.py 128 && python tools/scan_clocks.py 512 0.10062893578616353 0.08 && python tools/scan_clocks.py 0) > gpurun_out/phase_clocks_prefix.txt 2>&1
python euler-_solver_-openmp-_b200/tools/small_grids.py > gpurun_out/small_grids.txt 2>&1
