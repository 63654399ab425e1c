#!/bin/bash
# round-2 experiment D: ncu --set full of the old (round-1) and the persistent kernel, 100 fused steps
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python benchmarks/one_launch.py --nsteps 100 --launches 3 > gpurun_out/r2d_plain_new.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lin2 -s 2 -c 1 -o gpurun_out/r2d_new python benchmarks/one_launch.py --nsteps 100 --launches 3 > gpurun_out/r2d_ncu_new.log 2>&1
CRB_LIB=$PWD/build/libcrb_r1.so python benchmarks/one_launch.py --nsteps 100 --launches 3 > gpurun_out/r2d_plain_old.log 2>&1 &&
CRB_LIB=$PWD/build/libcrb_r1.so ncu --set full --clock-control none --import-source on -k regex:lin2 -s 2 -c 1 -o gpurun_out/r2d_old python benchmarks/one_launch.py --nsteps 100 --launches 3 > gpurun_out/r2d_ncu_old.log 2>&1
tail -2 gpurun_out/r2d_plain_new.log gpurun_out/r2d_ncu_new.log gpurun_out/r2d_plain_old.log gpurun_out/r2d_ncu_old.log
ls -la gpurun_out/*.ncu-rep
