#!/bin/bash
# crb_sincos with the small-angle fast path: parity + gravity configurations
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for i in 1 2; do python benchmarks/bench_configs.py --only cfg1e,cfg4,cfg5,lqr 2>/dev/null | cut -c1-200; done
