#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "default"; python benchmarks/bench_configs.py --only cfg4 2>/dev/null
for v in q0_m1_u1 q1_m6_u1 q1_m5_u1; do
  echo $v; CRB_LIB=$PWD/build/libcrb_rk45_$v.so python benchmarks/bench_configs.py --only cfg4 2>/dev/null
done
CRB_LIB=$PWD/build/libcrb_rk45_q1_m6_u1.so python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "rk45" 2>&1 | tail -2
