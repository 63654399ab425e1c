#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests/test_gpu_lqr.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -3
python benchmarks/bench_lqr_long.py --out gpurun_out/r2z_lqr_long.json 2>/dev/null
