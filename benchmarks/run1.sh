set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 2000 --warmup 200 --no-cpu > gpurun_out/b2.json 2>gpurun_out/b2.err; cat gpurun_out/b2.json; tail -3 gpurun_out/b2.err
python benchmarks/e2e_probe.py > gpurun_out/e2e_probe2.json 2>&1
cat gpurun_out/e2e_probe2.json
