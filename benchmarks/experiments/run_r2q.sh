#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log
tail -5 gpurun_out/r2q_pytest.log | cut -c1-300
timeout 900 python benchmarks/bench_lqr_long.py --out gpurun_out/r2q_lqr_long.json 2> gpurun_out/r2q_lqr_long.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2q_bench_20.json 2> gpurun_out/r2q_bench_20.err; echo "bench20 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2q_bench_20.json')); print('value', d['value'], 'frac', d['roofline']['frac'], 'fp64', d['roofline']['fp64']['frac'], 'e2e', d['e2e']['value'], d['e2e']['host_traffic_gbs'], d['e2e']['ceiling_gbs']); print({k:(v.get('value'),v.get('unit')) for k,v in d['secondary'].items()}); print(d['cpu_baseline']['kind'], d['cpu_baseline']['value'], d['cpu_baseline']['port']['value'])"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
