#!/bin/bash
# profiles/ <- gpurun_out/ after `bash benchmarks/run_r2_final_a.sh; bash benchmarks/run_r2_final_b.sh` on a GPU box
cd "$(dirname "$0")/.."
python profiles/secondary_pipes.py gpurun_out/r2s_secondary_ncu.csv > /dev/null
cp gpurun_out/r2s_secondary_ncu.csv profiles/r2_secondary_ncu_metrics.csv
cp gpurun_out/r2f_launches.csv profiles/r2_launches.csv
cp gpurun_out/r2f_bench_configs.jsonl profiles/r2_bench_configs.jsonl
cp gpurun_out/r2f_bench_1gpu_steps20.json profiles/r2_bench_1gpu_steps20_final.json
cp gpurun_out/r2f_bench_1gpu_default.json profiles/r2_bench_1gpu_default_final.json
cp gpurun_out/r2f_bench_reference_arm.json profiles/r2_bench_reference_arm_final.json
for r in lin2p rk45 shared; do
  { head -2 profiles/r2_$( [ $r = lin2p ] && echo rk4_lin2p || echo $r )_final_ncu_summary.txt | grep -v "^---\|^Kernel Name"; python profiles/ncu_extract.py gpurun_out/r2f_${r}_raw.csv; python - $r <<'PY'
import csv,sys
rows=list(csv.reader(open('gpurun_out/r2f_%s_raw.csv'%sys.argv[1])))
h,r=rows[0],rows[2]
keep=('sm__warps_active.avg.per_cycle_active','sm__icc_request_hit_rate.pct','gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed','sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','launch__shared_mem_per_block_dynamic','sm__pipe_tma_cycles_active.avg.pct_of_peak_sustained_active')
for a,b in zip(h,r):
    if 'issue_stalled' in a and 'per_issue_active.ratio' in a and float(b or 0)>0.05: print(a,'=',b)
    if a in keep: print(a,'=',b)
PY
  } > /tmp/sum_$r.txt
  cp /tmp/sum_$r.txt profiles/r2_$( [ $r = lin2p ] && echo rk4_lin2p || echo $r )_final_ncu_summary.txt
done
