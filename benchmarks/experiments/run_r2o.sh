#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python benchmarks/bench_lqr_long.py --sizes 32:148 > gpurun_out/r2o_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:crb_lqr_kernel -c 1 -o gpurun_out/r2o_lqr32 python benchmarks/bench_lqr_long.py --sizes 32:148 > gpurun_out/r2o_ncu.log 2>&1
tail -3 gpurun_out/r2o_plain.log; tail -3 gpurun_out/r2o_ncu.log
