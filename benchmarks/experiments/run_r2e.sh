#!/bin/bash
# round-2 experiment E: ticket-scheduled persistent kernel + new input / plug-in tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
timeout 300 python benchmarks/launch_sweep.py --label r2e_tickets --out gpurun_out/r2e_sweep.json > gpurun_out/r2e_sweep.log 2>&1
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2e_bench_20.json 2> gpurun_out/r2e_bench_20.err
timeout 300 python bench.py --gpus 1 --steps 1000 --warmup 50 --no-cpu > gpurun_out/r2e_bench_1000.json 2> gpurun_out/r2e_bench_1000.err
tail -25 gpurun_out/r2e_pytest.log
cat gpurun_out/r2e_sweep.json
cut -c1-300 gpurun_out/r2e_bench_20.json gpurun_out/r2e_bench_1000.json
