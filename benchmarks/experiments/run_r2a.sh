#!/bin/bash
# round-2 experiment A: parity of the persistent kernel, launch-time table old vs new, driver-style bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
CRB_LIB=$PWD/build/libcrb_r1.so timeout 300 python benchmarks/launch_sweep.py --label r1 --out gpurun_out/r2a_sweep_r1.json > gpurun_out/r2a_sweep_r1.log 2>&1
timeout 300 python benchmarks/launch_sweep.py --label r2_persistent --out gpurun_out/r2a_sweep_r2.json > gpurun_out/r2a_sweep_r2.log 2>&1
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2a_bench_20.json 2> gpurun_out/r2a_bench_20.err
timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu > gpurun_out/r2a_bench_50.json 2> gpurun_out/r2a_bench_50.err
timeout 300 python bench.py --gpus 1 --steps 1000 --warmup 50 --no-cpu > gpurun_out/r2a_bench_1000.json 2> gpurun_out/r2a_bench_1000.err
tail -3 gpurun_out/r2a_pytest.log
cat gpurun_out/r2a_sweep_r1.json gpurun_out/r2a_sweep_r2.json
cat gpurun_out/r2a_bench_20.json gpurun_out/r2a_bench_50.json gpurun_out/r2a_bench_1000.json
