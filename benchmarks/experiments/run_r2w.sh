#!/bin/bash
# RK45 (cfg 4): what bounds it?  occupancy sweep through padded shared memory, lock-stepped warps, rolled stage inputs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
B="python benchmarks/bench_configs.py --only cfg4"
run() { echo "== $1"; shift; env "$@" $B 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('%.3f ms  att mean %.1f max %.0f  %.3g el-att/s' % (d['ms'], d['attempts_mean'], d['attempts_max'], d['element_attempts_per_s']))"; }
V=$PWD/build/variants
for U in "" "CRB_CFG4_UNIFORM=1"; do
  echo "#### $U"
  run "base 8 warps/SM" CRB_LIB=$V/libcrb_rk45_base.so X=1 $U
  run "base 6 warps/SM (pad 30K)" CRB_LIB=$V/libcrb_rk45_base.so CRB_RK45_PAD_SMEM=56000 $U
  run "base 4 warps/SM (pad 100K)" CRB_LIB=$V/libcrb_rk45_base.so CRB_RK45_PAD_SMEM=100000 $U
  run "base 2 warps/SM (pad 200K)" CRB_LIB=$V/libcrb_rk45_base.so CRB_RK45_PAD_SMEM=200000 $U
  run "rolled" CRB_LIB=$V/libcrb_rk45_rolled.so X=1 $U
  run "wpb8 (no barrier)" CRB_LIB=$V/libcrb_rk45_wpb8.so X=1 $U
  run "lock8" CRB_LIB=$V/libcrb_rk45_lock8.so X=1 $U
  run "lock4" CRB_LIB=$V/libcrb_rk45_lock4.so X=1 $U
done
CRB_LIB=$V/libcrb_rk45_lock8.so python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "rk45" 2>&1 | tail -2
