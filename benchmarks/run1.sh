python bench.py > gpurun_out/r1b_bench_1gpu.json 2> gpurun_out/r1b_bench_1gpu.err
python bench.py --impl reference > gpurun_out/r1b_bench_reference_arm.json 2> gpurun_out/r1b_ref.err
python benchmarks/bench_configs.py --only cfg2,cfg4,cfg5,cfg3g,cfg3m > gpurun_out/r1b_bench_configs.json 2>&1
python bench.py --steps 100 --warmup 50 --no-cpu > gpurun_out/plain_l.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1b_launches.csv python bench.py --steps 100 --warmup 50 --no-cpu > gpurun_out/ncu_l.log 2>&1
python benchmarks/bench_configs.py --only cfg5 > gpurun_out/plain_c5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:shared -s 1 -c 1 -o gpurun_out/prof_shared python benchmarks/bench_configs.py --only cfg5 > gpurun_out/ncu_c5.log 2>&1
ncu -i gpurun_out/prof_shared.ncu-rep --page raw --csv > gpurun_out/prof_shared_raw.csv 2>/dev/null
cat gpurun_out/r1b_bench_1gpu.json gpurun_out/r1b_bench_reference_arm.json gpurun_out/r1b_bench_configs.json
