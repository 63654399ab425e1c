#!/bin/bash
# 8 GPUs of one box: host <-> device copy ceiling (no kernels) and the bench line at N = 8
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
rm -f gpurun_out/r2_host_ceiling_8gpu.jsonl
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 1 2 4 8; do
  timeout 200 $TR --nproc-per-node $n --master-port $((29500+n)) benchmarks/host_ceiling.py --mode ranks --out gpurun_out/r2_host_ceiling_8gpu.jsonl > gpurun_out/r2_hc_ranks_$n.log 2>&1; echo "hc ranks $n rc=$?"
done
timeout 200 $TR --nproc-per-node 8 --master-port 29520 benchmarks/host_ceiling.py --mode ranks --bind --out gpurun_out/r2_host_ceiling_8gpu.jsonl > gpurun_out/r2_hc_ranks_8_bind.log 2>&1; echo "hc ranks 8 bind rc=$?"
timeout 200 python benchmarks/host_ceiling.py --mode single --gpus 8 --out gpurun_out/r2_host_ceiling_8gpu.jsonl > gpurun_out/r2_hc_single_8.log 2>&1; echo "hc single 8 rc=$?"
timeout 400 $TR --nproc-per-node 8 --master-port 29530 bench.py --gpus 8 --steps 1000 --warmup 50 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err; echo "bench 8 rc=$?"
timeout 400 $TR --nproc-per-node 8 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_8gpu_steps20.json 2> gpurun_out/r2_bench_8gpu_steps20.err; echo "bench 8 (20) rc=$?"
cat gpurun_out/r2_host_ceiling_8gpu.jsonl
cut -c1-600 gpurun_out/r2_bench_8gpu.json; python -c "
import json
for f in ('gpurun_out/r2_bench_8gpu.json','gpurun_out/r2_bench_8gpu_steps20.json'):
    d=json.load(open(f)); print(f, d['value'], d['e2e']['value'], d['roofline']['frac'], d.get('final_gather_ms'))
"
