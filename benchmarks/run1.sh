python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --steps 2000 --warmup 200 --no-cpu > gpurun_out/b3.json 2>gpurun_out/b3.err; python - <<PY
import json
d=json.load(open("gpurun_out/b3.json")); print("BENCH", d["value"], d["kernel_ms_per_launch"], d["e2e"]["value"], d["roofline"]["frac"])
PY
python benchmarks/bench_configs.py --only cfg3m,cfg3g,cfg5 2>&1 | tail -8
