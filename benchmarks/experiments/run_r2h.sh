#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python benchmarks/bench_recording.py --out gpurun_out/r2h_recording.json > gpurun_out/r2h_recording.log 2>&1; echo "recording rc=$?"
rm -f gpurun_out/r2h_host_ceiling_1gpu.json
timeout 300 python benchmarks/host_ceiling.py --mode single --gpus 1 --out gpurun_out/r2h_host_ceiling_1gpu.json > gpurun_out/r2h_hc.log 2>&1; echo "hc rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2h_recording.json'))
for r in d['rows']:
    print(r['method'], r['save_every'], r['recording'], round(r['ms'],3), '%.3e'%r['element_steps_per_s'], round(r['frame_write_gbs'],1), round(r['frame_write_frac_of_hbm_peak'],3))
PY
cat gpurun_out/r2h_host_ceiling_1gpu.json
