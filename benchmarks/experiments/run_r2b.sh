#!/bin/bash
# round-2 experiment B: persistent kernel with uniform control flow
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
timeout 300 python benchmarks/launch_sweep.py --label r2b_persistent --out gpurun_out/r2b_sweep.json > gpurun_out/r2b_sweep.log 2>&1
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2b_bench_20.json 2> gpurun_out/r2b_bench_20.err
timeout 300 python bench.py --gpus 1 --steps 1000 --warmup 50 --no-cpu > gpurun_out/r2b_bench_1000.json 2> gpurun_out/r2b_bench_1000.err
tail -3 gpurun_out/r2b_pytest.log
cat gpurun_out/r2b_sweep.json
cut -c1-400 gpurun_out/r2b_bench_20.json gpurun_out/r2b_bench_1000.json
