#!/bin/bash
# persistent kernel: tile tickets drawn one tile ahead (CRB_TICKET_AHEAD), A/B on one box
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
CRB_LIB=$PWD/build/variants/libcrb_rk4_fast_ahead.so python -m pytest tests/test_gpu_parity.py tests/test_gpu_outputs.py tests/test_gpu_midpoint.py tests/test_gpu_fullsize.py -m gpu -q -x -k "config3 or paired or persistent or midpoint or lean" 2>&1 | tail -2
for i in 1 2 3; do
  for lib in continuum_robot_b200/libcrb.so build/variants/libcrb_rk4_fast_ahead.so; do
    echo -n "$lib: "
    CRB_LIB=$PWD/$lib python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --no-secondary 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.4g  frac %.3f  kernel %.4f ms' % (d['value'], d['roofline']['frac'], d['kernel_ms_per_launch']))"
  done
done
for lib in continuum_robot_b200/libcrb.so build/variants/libcrb_rk4_fast_ahead.so; do echo -n "$lib default: "; CRB_LIB=$PWD/$lib python bench.py --no-cpu --no-secondary 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.4g  frac %.3f  kernel %.4f ms' % (d['value'], d['roofline']['frac'], d['kernel_ms_per_launch']))"; done
