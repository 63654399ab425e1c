#!/bin/bash
# instruction-cache metrics of every kernel family
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
MET=sm__icc_request_hit_rate.pct,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,launch__registers_per_thread,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,sm__warps_active.avg.per_cycle_active
ncu --metrics $MET --clock-control none -k regex:"crb_rk4|crb_rk45|crb_midpoint|crb_lqr" --csv --log-file gpurun_out/r2ae_icache.csv python benchmarks/bench_configs.py --only cfg1e,cfg2,cfg2m,cfg4,cfg5,cfg3p,cfg3g,cfg3m,cfg3i,cfg3mid,lqr > gpurun_out/r2ae.log 2>&1
tail -3 gpurun_out/r2ae.log; ncu --metrics $MET --clock-control none -k regex:"crb_rk4" --csv --log-file gpurun_out/r2ae_icache3.csv python benchmarks/one_launch.py --nsteps 20 --launches 2 > /dev/null 2>&1; cat gpurun_out/r2ae_icache3.csv >> gpurun_out/r2ae_icache.csv
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2ae_icache.csv')) if len(r)>14 and r[0].isdigit()]
L={}
for r in rows: L.setdefault(int(r[0]),{'k':r[4][:70]})[r[12].split('.')[0][-28:]]=r[14]
seen={}
for i in sorted(L): seen[L[i]['k']]=L[i]
for k,v in seen.items(): print(k, {a:b for a,b in v.items() if a!='k'})
PY
