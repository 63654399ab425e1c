#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
tail -5 gpurun_out/r2j_pytest.log | cut -c1-300
timeout 600 python benchmarks/bench_recording.py --out gpurun_out/r2j_recording.json > gpurun_out/r2j_recording.log 2>&1; echo "recording rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2j_recording.json'))
for r in d['rows']:
    print(r['method'], r['save_every'], r['recording'], round(r['ms'],3), '%.3e'%r['element_steps_per_s'], round(r['frame_write_gbs'],1), round(r['frame_write_frac_of_hbm_peak'],3))
PY
timeout 300 python benchmarks/launch_sweep.py --label r2j --nsteps 1,20,50,100 --out gpurun_out/r2j_sweep.json > gpurun_out/r2j_sweep.log 2>&1
cat gpurun_out/r2j_sweep.json | cut -c1-900
