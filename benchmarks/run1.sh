python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python benchmarks/bench_configs.py --only cfg2,cfg4,cfg3g,cfg5 2>&1 | tail -8
