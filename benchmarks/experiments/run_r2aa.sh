#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:crb_rk4_shared_kernel -c 1 -f -o gpurun_out/r2aa_shared python benchmarks/bench_configs.py --only cfg5 > gpurun_out/r2aa_ncu_shared.log 2>&1
ls -la gpurun_out/r2aa_shared.ncu-rep
