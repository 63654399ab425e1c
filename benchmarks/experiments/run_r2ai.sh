#!/bin/bash
# line-level ncu capture of the nonlinear RK4 kernel (cfg 2), exported to CSV on the box
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"crb_rk4_kernel" -c 1 -f -o gpurun_out/r2ai_cfg2 python benchmarks/bench_configs.py --only cfg2 > gpurun_out/r2ai.log 2>&1
ncu -i gpurun_out/r2ai_cfg2.ncu-rep --page raw --csv > gpurun_out/r2ai_cfg2_raw.csv 2>/dev/null
ncu -i gpurun_out/r2ai_cfg2.ncu-rep --page source --csv --print-source sass > gpurun_out/r2ai_cfg2_sass.csv 2>/dev/null
rm -f gpurun_out/r2ai_cfg2.ncu-rep; ls -la gpurun_out
