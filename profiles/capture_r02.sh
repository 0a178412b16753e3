# Round 2 captures (run under gpurun, one GPU): the steady-state launch of the dominant kernel of the bench build, with the
# counters bench.py's roofline block reads (profiles/ncu_to_counters.py turns the raw page into profiles/r02_counters.json).
#   g128: persistent_scan_kernel, the 200-step launch of the default bench (launch 0 is the 20-step warm-up)
#   g512: persistent_chunks_kernel (the kernel of the xi-sharded ranks) on one rank: the 20-step launch of the sharded leg
#         (launch 0: its 5 warm-up steps)
set -x
EXTRA=lts__t_sectors_srcunit_tex_op_read.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum
ARGS128="--steps 200 --warmup 20 --no-cpu-baseline --batch-sets 0 --no-sweep --sharded none"
ARGS512="--steps 20 --warmup 5 --no-cpu-baseline --batch-sets 0 --no-sweep --sharded g512 --sharded-steps 20"
python bench.py $ARGS128 > gpurun_out/r02_plain_g128.log 2>&1 && \
ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:persistent_scan_kernel -s 1 -c 1 -o gpurun_out/r02_persistent_scan_g128 -f \
    python bench.py $ARGS128 > gpurun_out/r02_ncu_g128.log 2>&1
python bench.py $ARGS512 > gpurun_out/r02_plain_g512.log 2>&1 && \
ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:persistent_chunks_kernel -s 1 -c 1 -o gpurun_out/r02_persistent_chunks_g512 -f \
    python bench.py $ARGS512 > gpurun_out/r02_ncu_g512.log 2>&1
tail -2 gpurun_out/r02_ncu_g128.log gpurun_out/r02_ncu_g512.log
ls -la gpurun_out | grep r02_
