#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_lqr.py -m gpu -q -x > gpurun_out/r2p_pytest_lqr.log 2>&1; echo "pytest lqr rc=$?" >> gpurun_out/r2p_pytest_lqr.log
tail -4 gpurun_out/r2p_pytest_lqr.log | cut -c1-300
timeout 900 python benchmarks/bench_lqr_long.py --sizes 6:8192,14:1024,16:592,24:296,32:296 --out gpurun_out/r2p_lqr_long.json 2> gpurun_out/r2p_lqr_long.err
