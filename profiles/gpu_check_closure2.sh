# second GPU check of the closure extension: its tests, its step times, one ncu capture of the sums kernel at 16 384 points
TAG=${1:-r02o}
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity_closure.py -q -m gpu > gpurun_out/${TAG}_pytest_closure.log 2>&1; echo "closure pytest exit $?" | tee -a gpurun_out/${TAG}_pytest_closure.log
tail -40 gpurun_out/${TAG}_pytest_closure.log | cut -c1-400
timeout 120 python euler-_solver_-openmp-_b200/tools/closure_step_timing.py > gpurun_out/${TAG}_closure_timing.txt 2>&1; echo "timing exit $?"; cat gpurun_out/${TAG}_closure_timing.txt | cut -c1-300
timeout 150 ncu --set full --clock-control none --import-source on -k regex:w_sums_kernel --launch-skip 4 --launch-count 2 -f -o gpurun_out/${TAG}_ncu_w_sums_g128 \
    python euler-_solver_-openmp-_b200/tools/closure_step_timing.py 2 128 > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu exit $?"; tail -3 gpurun_out/${TAG}_ncu.log | cut -c1-300
