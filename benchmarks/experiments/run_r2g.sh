#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2g_bench_20.json 2> gpurun_out/r2g_bench_20.err; echo "bench20 rc=$?"
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2g_ref_20.json 2> gpurun_out/r2g_ref_20.err; echo "ref20 rc=$?"
timeout 600 python bench.py --no-cpu --no-secondary > gpurun_out/r2g_bench_default.json 2> gpurun_out/r2g_bench_default.err; echo "bench default rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1; echo "smoke rc=$?"
tail -30 gpurun_out/r2g_pytest.log | cut -c1-300
cat gpurun_out/r2g_bench_20.json; tail -5 gpurun_out/r2g_bench_20.err
cat gpurun_out/r2g_ref_20.json
cut -c1-1200 gpurun_out/r2g_bench_default.json; tail -3 gpurun_out/r2g_bench_default.err
