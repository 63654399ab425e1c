python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "per_member" 2>&1 | tail -15
python benchmarks/bench_configs.py --only cfg3m,cfg3g 2>&1 | tail -5
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
