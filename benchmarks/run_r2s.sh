#!/bin/bash
# ncu pipe metrics of the secondary configurations (one pass, metrics only)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python benchmarks/bench_configs.py --only cfg1e,cfg2,cfg4,cfg5 > gpurun_out/r2s_plain.log 2>&1 &&
ncu --metrics sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.per_cycle_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sass__inst_executed_local_loads,sass__inst_executed_local_stores,launch__registers_per_thread,gpu__time_duration.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__inst_executed_pipe_fp64.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum --clock-control none -k regex:"crb_rk4|crb_rk45" --csv --log-file gpurun_out/r2s_secondary_ncu.csv python benchmarks/bench_configs.py --only cfg1e,cfg2,cfg4,cfg5 > gpurun_out/r2s_ncu.log 2>&1
tail -3 gpurun_out/r2s_ncu.log; wc -l gpurun_out/r2s_secondary_ncu.csv
