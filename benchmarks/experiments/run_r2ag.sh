#!/bin/bash
# RK45: pilot launch + longest-first member order
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_inputs.py tests/test_gpu_outputs.py tests/test_gpu_api.py tests/test_gpu_fullsize.py -m gpu -q -x -k "rk45 or RK45 or config4 or adaptive" 2>&1 | tail -3
for P in 0 4 8 12 20 32; do echo "== pilot $P"; CRB_RK45_PILOT=$P python benchmarks/bench_configs.py --only cfg4,cfg4x4 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  %s: %.3f ms  %.3g el-att/s' % (d['config'][:22], d['ms'], d['element_attempts_per_s']))"; done
echo "== auto"; python benchmarks/bench_configs.py --only cfg4,cfg4x4 2>/dev/null | cut -c1-200
