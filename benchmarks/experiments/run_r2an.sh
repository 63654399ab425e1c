#!/bin/bash
# sincos coefficients in constant memory: parity + gravity configurations
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for i in 1 2; do python benchmarks/bench_configs.py --only cfg1e,cfg4,cfg5,lqr 2>/dev/null | grep -v "LQR synthesis\|banded" | cut -c1-200; done
