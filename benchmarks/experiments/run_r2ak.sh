#!/bin/bash
# paired operator form with the summed rotation stiffness (CRB_FAST_CWS): parity + A/B on one box
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for i in 1 2 3; do
  for lib in "" "$PWD/build/variants/libcrb_rk4_fast_nocws.so"; do
    echo -n "lib=${lib:-in-tree (cws)}: "
    CRB_LIB=${lib:-$PWD/continuum_robot_b200/libcrb.so} python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --no-secondary 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.4g  frac %.3f  kernel %.4f ms' % (d['value'], d['roofline']['frac'], d['kernel_ms_per_launch']))"
  done
done
for lib in "" "$PWD/build/variants/libcrb_rk4_fast_nocws.so"; do echo "lib=${lib:-in-tree (cws)}"; CRB_LIB=${lib:-$PWD/continuum_robot_b200/libcrb.so} python benchmarks/launch_sweep.py 2>/dev/null | tail -3 | cut -c1-300; done
