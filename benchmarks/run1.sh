set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in "" _B _C _D; do
  CRB_LIB=$PWD/continuum_robot_b200/libcrb$v.so python bench.py --steps 2000 --warmup 200 --no-cpu > gpurun_out/var$v.json 2>gpurun_out/var$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/var$v.json")); print("VARIANT '$v'", d["value"], d["kernel_ms_per_launch"], d["e2e"]["value"])
PY
done
PROBE_BIND=0 python benchmarks/e2e_probe.py > gpurun_out/e2e_probe_unbound.json 2>&1
PROBE_BIND=1 python benchmarks/e2e_probe.py > gpurun_out/e2e_probe_bound.json 2>&1
cat gpurun_out/e2e_probe_unbound.json gpurun_out/e2e_probe_bound.json
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1; lscpu | grep -i -E "numa|model name|socket" >> gpurun_out/topo.txt
