#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log
tail -5 gpurun_out/r2l_pytest.log | cut -c1-300
timeout 300 python benchmarks/bench_configs.py --only cfg4,cfg2,cfg5 > gpurun_out/r2l_cfgs.jsonl 2> gpurun_out/r2l_cfgs.err; cat gpurun_out/r2l_cfgs.jsonl
CRB_RK45_SLOTS=0 timeout 300 python benchmarks/bench_configs.py --only cfg4 > gpurun_out/r2l_cfg4_m4.jsonl 2> gpurun_out/r2l_cfg4_m4.err; cat gpurun_out/r2l_cfg4_m4.jsonl
for cm in 0 4736 2368; do
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --no-secondary --e2e-chunk-members $cm > gpurun_out/r2l_bench20_chunk$cm.json 2> gpurun_out/r2l_bench20_chunk$cm.err
python -c "
import json; d=json.load(open('gpurun_out/r2l_bench20_chunk$cm.json')); print('chunk', $cm, 'value', d['value'], 'e2e', d['e2e']['value'], 'kernel_ms', d['kernel_ms_per_launch'], 'ceiling', d['e2e'].get('ceiling_gbs'), d['e2e'].get('host_traffic_gbs'))"
done
