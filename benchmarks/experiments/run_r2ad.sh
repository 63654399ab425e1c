#!/bin/bash
# RK45 (cfg 4): register caps with the maximum shared-memory carve-out (10 / 12 warps per SM)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
B="python benchmarks/bench_configs.py --only cfg4"
run() { echo "== $1"; shift; env "$@" $B 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('%.3f ms  att mean %.1f max %.0f  %.3g el-att/s' % (d['ms'], d['attempts_mean'], d['attempts_max'], d['element_attempts_per_s']))"; }
V=$PWD/build/variants
MET=launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem,launch__shared_mem_per_block_dynamic,launch__shared_mem_config_size,sm__warps_active.avg.per_cycle_active,gpu__time_duration.sum,launch__registers_per_thread,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__icc_request_hit_rate.pct,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed
{
for U in "X=1" "CRB_CFG4_UNIFORM=1"; do
  echo "#### $U"
  for n in base r200 r184 r168 r168w3 r168w4; do run $n CRB_LIB=$V/libcrb_rk45_$n.so $U; done
done
for n in r200 r168 r168w3; do
  echo "== ncu uniform $n"
  CRB_CFG4_UNIFORM=1 CRB_LIB=$V/libcrb_rk45_$n.so ncu --metrics $MET --clock-control none -k regex:crb_rk45_kernel -c 1 python benchmarks/bench_configs.py --only cfg4 2>/dev/null | grep -E "^\s+(sm__|gcc__|smsp__|gpu__|launch__|l1tex)"
done
} 2>&1 | tee gpurun_out/r2ad.log
