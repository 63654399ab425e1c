#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_lqr.py -m gpu -q -x > gpurun_out/r2m_pytest_lqr.log 2>&1; echo "pytest lqr rc=$?" >> gpurun_out/r2m_pytest_lqr.log
tail -15 gpurun_out/r2m_pytest_lqr.log | cut -c1-300
timeout 900 python benchmarks/bench_lqr_long.py --out gpurun_out/r2m_lqr_long.json 2> gpurun_out/r2m_lqr_long.err
timeout 300 python benchmarks/bench_configs.py --only cfg5 > gpurun_out/r2m_cfg5.jsonl 2> gpurun_out/r2m_cfg5.err; cat gpurun_out/r2m_cfg5.jsonl
