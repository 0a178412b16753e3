# ncu capture of the closure sums kernel at 16 384 points (run under gpurun): profiles/r02o_ncu_w_sums_g128.csv came from this
TAG=${1:-r02o}
mkdir -p gpurun_out
timeout 150 ncu --set full --clock-control none --import-source on -k regex:w_sums_kernel --launch-skip 4 --launch-count 2 -f -o gpurun_out/${TAG}_ncu_w_sums_g128 \
    python euler-_solver_-openmp-_b200/tools/closure_step_timing.py 2 128 > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu exit $?"; tail -3 gpurun_out/${TAG}_ncu.log | cut -c1-300
