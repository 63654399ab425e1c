python benchmarks/bench_configs.py --only cfg3p 2>&1 | tail -3
