#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python benchmarks/bench_lqr_long.py --out gpurun_out/r2n_lqr_long.json 2> gpurun_out/r2n_lqr_long.err
tail -3 gpurun_out/r2n_lqr_long.err
