#!/bin/bash
# final build on 8 GPUs of one box: bench line (driver's arguments and defaults), BASELINE config 5 sharded by member
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2f_bench_8gpu_steps20.json 2> gpurun_out/r2f_bench_8gpu_steps20.err; echo "bench 8 (20) rc=$?"
timeout 300 $TR --nproc-per-node 8 --master-port 29530 bench.py --gpus 8 --steps 1000 --warmup 50 > gpurun_out/r2f_bench_8gpu.json 2> gpurun_out/r2f_bench_8gpu.err; echo "bench 8 rc=$?"
timeout 300 $TR --nproc-per-node 8 --master-port 29532 benchmarks/bench_cfg5_sharded.py > gpurun_out/r2f_cfg5_8gpu.json 2> gpurun_out/r2f_cfg5_8gpu.err; echo "cfg5 rc=$?"
timeout 300 $TR --nproc-per-node 8 --master-port 29533 benchmarks/bench_cfg5_sharded.py --designs > gpurun_out/r2f_cfg5_designs_8gpu.json 2> gpurun_out/r2f_cfg5_designs_8gpu.err; echo "cfg5 designs rc=$?"
tail -c 600 gpurun_out/r2f_cfg5_8gpu.json; tail -c 600 gpurun_out/r2f_cfg5_designs_8gpu.json
python -c "
import json
for f in ('gpurun_out/r2f_bench_8gpu.json','gpurun_out/r2f_bench_8gpu_steps20.json'):
    d=json.load(open(f)); print(f, d['value'], d['e2e']['value'], d['roofline']['frac'], d.get('final_gather_ms'))
"
