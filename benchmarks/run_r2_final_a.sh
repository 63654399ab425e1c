#!/bin/bash
# final build, part A: ncu captures (launch list of the driver's bench command, --set full of the RK45 / shared-operator
# kernels, pipe metrics of the secondary configurations), secondary benchmarks.  (gpurun brings back at most 64 MiB: the
# .ncu-rep files are exported to raw CSV pages on the box and deleted.)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python benchmarks/bench_configs.py --only cfg1e,cfg2,cfg2m,cfg4,cfg4x4,cfg5,cfg3p,cfg3g,cfg3m,cfg3i,cfg3mid,lqr 2>/dev/null > gpurun_out/r2f_bench_configs.jsonl; wc -l gpurun_out/r2f_bench_configs.jsonl
# launch list of the driver's command (only after the same command exited 0 without ncu)
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --no-secondary > gpurun_out/r2f_plain.json 2>gpurun_out/r2f_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --no-secondary > gpurun_out/r2f_launches.log 2>&1
wc -l gpurun_out/r2f_launches.csv
timeout 600 ncu --set full --clock-control none -k regex:crb_rk45_kernel -c 1 -f -o gpurun_out/r2f_rk45 python benchmarks/bench_configs.py --only cfg4 > gpurun_out/r2f_ncu_rk45.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:crb_rk4_shared_kernel -c 1 -f -o gpurun_out/r2f_shared python benchmarks/bench_configs.py --only cfg5 > gpurun_out/r2f_ncu_shared.log 2>&1
bash benchmarks/run_r2s.sh > /dev/null 2>&1
python benchmarks/one_launch.py --nsteps 20 --launches 3 > /dev/null 2>&1 &&
timeout 600 ncu --set full --clock-control none -k regex:lin2p -s 2 -c 1 -f -o gpurun_out/r2f_lin2p python benchmarks/one_launch.py --nsteps 20 --launches 3 > gpurun_out/r2f_ncu_lin2p.log 2>&1
for r in r2f_rk45 r2f_shared r2f_lin2p; do ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/${r}_raw.csv 2>/dev/null; rm -f gpurun_out/$r.ncu-rep; done
ls -la gpurun_out/; du -sh gpurun_out
