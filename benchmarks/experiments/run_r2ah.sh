#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
python -m pytest tests/test_gpu_parity.py tests/test_gpu_inputs.py tests/test_gpu_outputs.py tests/test_gpu_api.py tests/test_gpu_fullsize.py -m gpu -q -x -k "rk45 or RK45 or config4 or adaptive" 2>&1 | tail -2
for i in 1 2; do python benchmarks/bench_configs.py --only cfg4,cfg4x4 2>/dev/null | cut -c1-200; done
CRB_RK45_PILOT=0 python benchmarks/bench_configs.py --only cfg4 2>/dev/null | cut -c1-200
CRB_CFG4_UNIFORM=1 python benchmarks/bench_configs.py --only cfg4 2>/dev/null | cut -c1-200
