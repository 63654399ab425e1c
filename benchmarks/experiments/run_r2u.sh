#!/bin/bash
# shared-operator kernel with zero-tile skipping (cfg 5), parity tests of that path, and a line-level ncu capture of
# the RK45 kernel (cfg 4) and of the shared-operator kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_outputs.py tests/test_gpu_api.py -m gpu -q -x -k "shared or lqr or cfg5 or config5 or rollout" 2>&1 | tail -3
python benchmarks/bench_configs.py --only cfg5,cfg4 2>/dev/null | tee gpurun_out/r2u_cfg.json
timeout 600 ncu --set full --import-source on --clock-control none -k regex:crb_rk45_kernel -c 1 -f -o gpurun_out/r2u_rk45 python benchmarks/bench_configs.py --only cfg4 > gpurun_out/r2u_ncu_rk45.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:crb_rk4_shared_kernel -c 1 -f -o gpurun_out/r2u_shared python benchmarks/bench_configs.py --only cfg5 > gpurun_out/r2u_ncu_shared.log 2>&1
ls -la gpurun_out/*.ncu-rep
