#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_inputs.py -m gpu -q -x -k "rk45" 2>&1 | tail -2
python benchmarks/bench_configs.py --only cfg4,cfg2 2>/dev/null | tee gpurun_out/r2y_cfg.json
echo "ums variant"
CRB_LIB=$PWD/build/variants/libcrb_rk4_ums.so python benchmarks/bench_configs.py --only cfg2 2>/dev/null | tee -a gpurun_out/r2y_cfg.json
CRB_LIB=$PWD/build/variants/libcrb_rk4_ums.so python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cfg2 or configs_1_2 or random_topologies or extreme" 2>&1 | tail -2
