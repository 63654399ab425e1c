#!/bin/bash
# RK45 (cfg 4): one-warp blocks, rolled stage inputs, committed state in shared memory, register caps (10 / 11 / 12 warps per SM)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
B="python benchmarks/bench_configs.py --only cfg4"
run() { echo "== $1"; shift; env "$@" $B 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('%.3f ms  att mean %.1f max %.0f  %.3g el-att/s' % (d['ms'], d['attempts_mean'], d['attempts_max'], d['element_attempts_per_s']))"; }
V=$PWD/build/variants
for U in "X=1" "CRB_CFG4_UNIFORM=1"; do
  echo "#### $U"
  for n in base A B C D E; do run $n CRB_LIB=$V/libcrb_rk45_$n.so $U; done
done 2>&1 | tee gpurun_out/r2x.log
for n in A B; do CRB_LIB=$V/libcrb_rk45_$n.so python -m pytest tests/test_gpu_parity.py tests/test_gpu_inputs.py -m gpu -q -x -k "rk45" 2>&1 | tail -2; done | tee -a gpurun_out/r2x.log
