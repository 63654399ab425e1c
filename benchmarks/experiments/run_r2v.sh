#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_outputs.py tests/test_gpu_api.py -m gpu -q -x -k "shared or lqr or cfg5 or config5 or rollout" 2>&1 | tail -3
python benchmarks/bench_configs.py --only cfg5 2>/dev/null | tee gpurun_out/r2v_cfg.json
