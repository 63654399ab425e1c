for sl in 2 3 4 0; do echo "slots $sl"; CRB_RK45_SLOTS=$sl python benchmarks/bench_configs.py --only cfg4 2>&1 | tail -1; done
