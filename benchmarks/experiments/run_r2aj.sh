#!/bin/bash
# nonlinear element polynomials regrouped: parity + cfg 2 / cfg 4 timings
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for i in 1 2; do python benchmarks/bench_configs.py --only cfg2,cfg2m,cfg4,cfg4x4 2>/dev/null | cut -c1-210; done
