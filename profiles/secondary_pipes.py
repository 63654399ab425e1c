"""profiles/secondary_pipes.json from the metrics-only ncu pass over the secondary configurations
(benchmarks/run_r2s.sh -> gpurun_out/r2s_secondary_ncu.csv): python profiles/secondary_pipes.py CSV [units.json]

The LAST launch of each configuration's kernel counts (earlier ones are warm-up).  `fp64_vector_flops_per_unit` =
(2 DFMA + DADD + DMUL thread instructions) / units of the launch; units come from bench_configs.py's own workload
sizes (element-steps for RK4; for cfg 4 the attempt count is data dependent and the figure is left out)."""
import csv
import datetime
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNELS = {  # config -> (kernel-name prefix, units per launch or None)
    "cfg1": ("void crb_rk4_fast_kernel<3, 2, 1, 1, 1", 131072 * 10 * 100),
    "cfg2": ("void crb_rk4_kernel<3, 78, 3", 32768 * 20 * 200),
    "cfg4": ("void crb_rk45_kernel<2, 78, 5", None),
    "cfg5": ("void crb_rk4_shared_kernel<5, 2, 1", 131072 * 200),
}
M = {
    "fp64_pipe_pct": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "dmma_pipe_pct": "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "lsu_wavefronts_pct": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "warps_active_per_sm": "sm__warps_active.avg.per_cycle_active",
    "registers": "launch__registers_per_thread",
    "ms": "gpu__time_duration.sum",
}


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
    launches = {}
    for r in rows:
        try:
            launches.setdefault(int(r[0]), {"kernel": r[4]})[r[12]] = float(r[14].replace(",", ""))
        except ValueError:  # "n/a": metric not available in this pass
            pass
    out = {}
    for cfg, (prefix, units) in KERNELS.items():
        ids = [i for i in sorted(launches) if launches[i]["kernel"].startswith(prefix)]
        if not ids:
            continue
        L = launches[ids[-1]]
        e = {"kernel": L["kernel"][:90]}
        for k, m in M.items():
            if m in L:
                e[k] = round(L[m] / 1e6, 3) if k == "ms" else (int(L[m]) if k == "registers" else round(L[m], 1))
        if cfg != "cfg5":
            e.pop("dmma_pipe_pct", None)
        fl = [L.get("smsp__sass_thread_inst_executed_op_%s_pred_on.sum" % o) for o in ("dfma", "dadd", "dmul")]
        if units and all(v is not None for v in fl):
            e["fp64_vector_flops_per_unit"] = round((2 * fl[0] + fl[1] + fl[2]) / units, 1)
        out[cfg] = e
    commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    out["captured_on"] = {
        "round": 2, "date": datetime.date.today().isoformat(), "commit": commit,
        "how": "ncu --metrics ... --clock-control none -k regex:crb_rk4|crb_rk45, python benchmarks/bench_configs.py --only cfg1e,cfg2,cfg4,cfg5 (benchmarks/run_r2s.sh); last launch of each kernel; profiles/secondary_pipes.py",
        "note": "fp64_pipe_pct = sm__pipe_fp64_cycles_active (vector DFMA/DADD/DMUL only); dmma_pipe_pct = the FP64 tensor sub-pipe; the two share one pipe on B200 (benchmarks/micro/pipes.cu), so their sum is the pipe's load.  ptxas: 0 B spills for all four kernels",
    }
    json.dump(out, open(os.path.join(ROOT, "profiles", "secondary_pipes.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
