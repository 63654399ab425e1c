#!/bin/bash
# round-2 experiment C: one-warp persistent blocks, stagger sweep
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
for st in 0 80 165 330 700; do
CRB_STAGGER_NS=$st timeout 300 python benchmarks/launch_sweep.py --label r2c_stagger$st --nsteps 1,20,50,100 --reps 20 --out gpurun_out/r2c_sweep_$st.json > gpurun_out/r2c_sweep_$st.log 2>&1
done
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2c_bench_20.json 2> gpurun_out/r2c_bench_20.err
tail -3 gpurun_out/r2c_pytest.log
cat gpurun_out/r2c_sweep_*.json
cut -c1-300 gpurun_out/r2c_bench_20.json
