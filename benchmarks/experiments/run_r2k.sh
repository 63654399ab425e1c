#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python benchmarks/bench_recording.py --out gpurun_out/r2k_recording.json > gpurun_out/r2k_recording.log 2>&1; echo "recording rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2k_recording.json'))
for r in d['rows']:
    print(r['method'], r['save_every'], r['recording'], round(r['ms'],3), '%.3e'%r['element_steps_per_s'], round(r['frame_write_gbs'],1), round(r['frame_write_frac_of_hbm_peak'],3))
PY
python benchmarks/one_launch.py --method midpoint --nsteps 50 --save-every 1 --launches 3 > gpurun_out/r2k_plain_mid.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:midpoint -s 2 -c 1 -o gpurun_out/r2k_mid_rec python benchmarks/one_launch.py --method midpoint --nsteps 50 --save-every 1 --launches 3 > gpurun_out/r2k_ncu_mid.log 2>&1
tail -2 gpurun_out/r2k_plain_mid.log gpurun_out/r2k_ncu_mid.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2k_bench_20.json 2> gpurun_out/r2k_bench_20.err; echo "bench20 rc=$?"
cut -c1-1500 gpurun_out/r2k_bench_20.json
