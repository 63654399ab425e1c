#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
timeout 300 python benchmarks/launch_sweep.py --label r2f --nsteps 1,20,50,100 --out gpurun_out/r2f_sweep.json > gpurun_out/r2f_sweep.log 2>&1
tail -60 gpurun_out/r2f_pytest.log | cut -c1-300
cat gpurun_out/r2f_sweep.json
