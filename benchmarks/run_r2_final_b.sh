#!/bin/bash
# final build, part B: the bench lines (driver's arguments, defaults, reference arm)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2>/dev/null > gpurun_out/r2f_bench_reference_arm.json; echo "ref rc=$?"
python bench.py --gpus 1 --steps 20 --warmup 5 2>gpurun_out/r2f_bench20.err > gpurun_out/r2f_bench_1gpu_steps20.json; echo "bench20 rc=$?"
python bench.py 2>gpurun_out/r2f_bench_default.err > gpurun_out/r2f_bench_1gpu_default.json; echo "default rc=$?"
python -c "
import json
for f in ('gpurun_out/r2f_bench_1gpu_steps20.json','gpurun_out/r2f_bench_1gpu_default.json'):
    d=json.load(open(f)); print(f, '%.4g'%d['value'], '%.4g'%d['e2e']['value'], round(d['roofline']['frac'],3), round(d['roofline']['fp64']['frac'],3), {k:('%.3g'%v['value']) for k,v in d['secondary'].items() if isinstance(v,dict) and 'value' in v}, d['clocks'])
d=json.load(open('gpurun_out/r2f_bench_reference_arm.json')); print('reference', d['value'], d['cpu_baseline'])
"
python __graft_entry__.py smoke 2>&1 | tail -1
