#!/usr/bin/env python
"""bench.py -- beam-element RK4 steps/s for BASELINE config 3 (65,536 linear beams x 32 elements).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one classical RK4 step of the whole ensemble shard (4 RHS evaluations per member).  Steps are issued
in launches of ``--steps-per-launch`` fused steps (default 50 = one 1 ms output frame at h = 2e-5, the output
cadence of the reference's examples, example_utilities.py:21); the K timed steps are captured once in a CUDA graph
(before the warm-up) and replayed inside the timed region.

  value         element-steps/s with state resident in HBM (CUDA events around the K steps, max over ranks)
  e2e           same metric through the public host API with HOST buffers: every call copies the state from pinned
                host memory, runs its fused steps and copies the state back
  roofline      achieved / peak / frac: BASELINE's figure of merit -- ALGORITHMIC bytes (96 B per element-step,
                SURVEY 8d) / mean launch duration vs the measured HBM copy bandwidth of MEASURED_PEAKS.json.
                Fused launches keep the state in registers, so that is not the binding resource: `bound` names the
                real limiter (the FP64 pipe), `fp64` gives its fraction against the DFMA rate measured in this run
                (crb_probe_dfma), `dram` the measured DRAM traffic of the launch (ncu capture, profiles/traffic.json)
  cpu_baseline  the UNMODIFIED reference (oracle/_ref, installed by oracle/build_ref.py) on the host cores through its
                own CSV path and multiprocessing fan-out, bounded sample; `port` = the NumPy oracle port beside it
  secondary     BASELINE configs 1, 2, 4, 5 as ensembles (N = 1 only), each a sub-second measurement

``--impl reference`` times the reference's CPU implementation alone (oracle/_ref; the oracle port if it is absent).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_BYTES_PER_ELEMENT_STEP = 96.0  # q,v (2) x 3 DOF x 8 B x (read + write), SURVEY 8(d)
WORKLOAD = "cfg3: 65536-member linear beam ensemble (random per-element E, random IC), 32 elements, fixed-step RK4 FP64"
REF_DIR = os.path.join(ROOT, "oracle", "_ref")

_OUT = sys.stdout


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------------------------
# CPU arms (reported baselines, test infrastructure only)
# ---------------------------------------------------------------------------------------------
def _port_worker(args):
    """The NumPy oracle port (oracle/beam_oracle.py) stepping `members` of the config-3 ensemble."""
    members, n_elements, steps, seed = args
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    from continuum_robot_b200 import ensembles as ens
    from oracle import beam_oracle as bo

    e = ens.config3(max(members) + 1, n_elements, seed=seed)
    n = e.n_free
    stepping = 0.0
    for i in members:
        spec = bo.BeamSpec.uniform(n_elements)
        spec.elastic_modulus = e.E[i].copy()
        b = bo.BeamOracle(spec)  # assembly + M^-1: once per beam, outside the timed stepping
        u = np.zeros(n)
        x0 = np.concatenate([e.q0[i], e.v0[i]])
        t0 = time.perf_counter()
        bo.rk4_solve(lambda t, x: b.rhs(t, x, u), x0, 0.0, e.h, steps)
        stepping += time.perf_counter() - t0
    return stepping


def _reference_worker(args):
    """The UNMODIFIED reference (oracle/_ref): beam built through its own CSV path
    (examples/example_utilities.py:37-73 style), classical RK4 around `get_dynamic_system()` (north_star R1)."""
    members, n_elements, steps, seed = args
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    import warnings

    warnings.filterwarnings("ignore")
    sys.path.insert(0, REF_DIR)
    from continuum_robot.models.dynamic_beam_model import DynamicEulerBernoulliBeam  # the reference

    from continuum_robot_b200 import ensembles as ens

    e = ens.config3(max(members) + 1, n_elements, seed=seed)
    m = ens.material()
    n = e.n_free
    stepping = 0.0
    for i in members:
        with tempfile.NamedTemporaryFile(mode="w", delete=False, suffix=".csv") as f:
            f.write("length,elastic_modulus,moment_inertia,density,cross_area,type,boundary_condition,wetted_area,drag_coef\n")
            for k in range(n_elements):
                f.write(f"{m['length']},{e.E[i, k]},{m['I']},{m['rho']},{m['A']},linear,{'FIXED' if k == 0 else 'NONE'},"
                        f"{m['wetted_area']},{m['drag_coef']}\n")
            path = f.name
        beam = DynamicEulerBernoulliBeam(path)
        os.unlink(path)
        beam.create_system_func()
        beam.create_input_func()
        fun = beam.get_dynamic_system()
        u = np.zeros(n)
        x = np.concatenate([e.q0[i], e.v0[i]])
        h = e.h
        t0 = time.perf_counter()
        for k in range(steps):
            t = k * h
            k1 = fun(t, x, u)
            k2 = fun(t + 0.5 * h, x + 0.5 * h * k1, u)
            k3 = fun(t + 0.5 * h, x + 0.5 * h * k2, u)
            k4 = fun(t + h, x + h * k3, u)
            x = x + (h / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4)
        stepping += time.perf_counter() - t0
        if not np.isfinite(x).all():
            raise RuntimeError("reference diverged")
    return stepping


def cpu_throughput(kind: str, steps: int, members_per_core: int = 1, n_elements: int = 32, warmup: int = 2):
    """element-steps/s with one process per host core (the reference's own fan-out pattern,
    examples/beam_comparison_gravity.py:72-73)."""
    from multiprocessing import get_context

    cores = host_cores()
    worker = _reference_worker if kind == "reference" else _port_worker
    tasks = [(list(range(c * members_per_core, (c + 1) * members_per_core)), n_elements, steps, 1234) for c in range(cores)]
    ctx = get_context("spawn")
    with ctx.Pool(cores) as pool:
        pool.map(worker, [([0], n_elements, max(1, warmup), 1234)] * cores)  # imports + warm-up steps
        dt = max(pool.map(worker, tasks))  # slowest worker's time inside the RK4 loop
    total = cores * members_per_core * n_elements * steps
    return total / dt, cores, f"{cores * members_per_core} members x {steps} RK4 steps, {cores} processes"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_DIR, "continuum_robot"))


def cpu_baseline(ref_steps: int, port_steps: int, n_elements: int = 32):
    """The reference itself where it travelled (kind "reference"), with the oracle port beside it."""
    pv, cores, psample = cpu_throughput("port", port_steps, n_elements=n_elements)
    port = {"value": pv, "unit": "element-steps/s", "cores": cores, "kind": "port", "sample": psample}
    if not reference_available():
        return port
    rv, cores, rsample = cpu_throughput("reference", ref_steps, n_elements=n_elements)
    return {"value": rv, "unit": "element-steps/s", "cores": cores, "kind": "reference", "sample": rsample,
            "what": "unmodified cram9030/continuum-robot (oracle/_ref): DynamicEulerBernoulliBeam(csv) + classical RK4 around "
                    "get_dynamic_system(), one process per core", "port": port}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind = "reference" if reference_available() else "port"
    # each "step" advances a bounded sample (one member per host core) by one RK4 step; the reference needs ~3 ms per
    # member-step, so K is capped to keep the arm within a few minutes
    cap = 2000 if kind == "reference" else 4000
    steps = min(max(1, args.steps), cap)
    # ... and with few steps several members per core, so that the timed sample is about a second of stepping per core
    # rather than a few milliseconds (20 steps of one member are 30 ms)
    per_core = max(1, min(64, -(-600 // steps)))
    val, cores, sample = cpu_throughput(kind, steps, members_per_core=per_core, warmup=min(max(args.warmup, 1), 20))
    ms = 1e3 * (cores * 32) / val
    line = {
        "impl": "reference", "metric": "beam-element RK4 steps/sec", "value": val, "unit": "element-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample, "steps_run": steps},
        "cpu_baseline": {"value": val, "unit": "element-steps/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "element-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_OUT, flush=True)


# ---------------------------------------------------------------------------------------------
# clocks sampler (profiling recipe's nvidia-smi line)
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def window(self, t_begin, t_end):
        """Summary of the samples taken while the device ran the same load: from `t_begin` (spin-up) to `t_end`."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        rows = [r for (ts, r) in self.rows if t_begin <= ts <= t_end + 0.02] or [r for (_, r) in self.rows[-3:]]
        sm = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in rows for k in range(4) if len(r) > 5 + k and r[5 + k].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(rows)}

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def build_ensemble(rank: int, members: int, n_elements: int, device):
    from continuum_robot_b200 import ensembles as ens
    from continuum_robot_b200.dynamic_beam import BatchedDynamicEulerBernoulliBeam

    e = ens.config3(members, n_elements, seed=1234 + rank)  # rank 0 = the BASELINE ensemble
    m = ens.material()
    par = np.empty((members, n_elements, 7))
    par[:, :, 0], par[:, :, 2], par[:, :, 3], par[:, :, 4] = m["length"], m["I"], m["rho"], m["A"]
    par[:, :, 1] = e.E
    par[:, :, 5], par[:, :, 6] = m["wetted_area"], m["drag_coef"]
    beam = BatchedDynamicEulerBernoulliBeam({"params": par, "type": ["linear"] * n_elements}, device=device)
    beam.create_system_func()
    beam.create_input_func()
    x0 = np.concatenate([e.q0, e.v0], axis=1)
    return e, beam, x0


def probe_fp64_peak(dev, reps: int = 5):
    """DFMA rate of this device in TFLOP/s (crb_probe_dfma, CUDA events, best of `reps`)."""
    import ctypes as C

    import torch

    from continuum_robot_b200 import _lib

    lib = _lib.load()
    iters = 20000
    flops = C.c_int64()
    _lib.check(lib.crb_probe_dfma(iters, None, 0, C.byref(flops), None))
    threads = flops.value // (16 * iters)
    scratch = torch.empty(threads, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    best = 1e30
    for _ in range(reps + 1):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.check(lib.crb_probe_dfma(iters, scratch.data_ptr(), threads, C.byref(flops), stream))
        b.record()
        torch.cuda.synchronize(dev)
        best = min(best, a.elapsed_time(b))
    return flops.value / (best * 1e-3) / 1e12


def secondary_measurements():
    """BASELINE configs 1, 2, 4, 5 as ensembles: value + unit (+ the pipe shares ncu measured for the kernel,
    profiles/secondary_pipes.json, captured on the build named there)."""
    sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
    import bench_configs as bc

    pipes = {}
    try:
        pipes = json.load(open(os.path.join(ROOT, "profiles", "secondary_pipes.json")))
    except Exception:
        pass
    out = {}
    keymap = {"cfg1": "cfg1e", "cfg2": "cfg2", "cfg4": "cfg4", "cfg5": "cfg5"}
    for cfg, name in keymap.items():
        try:
            r = next(iter(bc.run([name], reps=2)))
        except Exception as ex:  # a secondary number never takes the headline line down
            out[cfg] = {"error": f"{type(ex).__name__}: {ex}"}
            continue
        if "element_attempts_per_s" in r:
            entry = {"value": r["element_attempts_per_s"], "unit": "element-attempts/s", "attempts_mean": r["attempts_mean"]}
        elif cfg == "cfg5":
            entry = {"value": r["member_steps_per_s"], "unit": "member-steps/s"}
        else:
            entry = {"value": r["element_steps_per_s"], "unit": "element-steps/s"}
        entry.update({"workload": r["config"], "ms": r["ms"]})
        if cfg in pipes:
            entry["ncu"] = pipes[cfg]
        out[cfg] = entry
    if "captured_on" in pipes:  # which build the `ncu` pipe shares were measured on
        out["ncu_captured_on"] = {k: pipes["captured_on"].get(k) for k in ("commit", "date", "round")}
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from continuum_robot_b200.integrate import HostPipeline, rk4_steps

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from continuum_robot_b200.sharding import bind_to_gpu_numa_node

    numa_bound = bind_to_gpu_numa_node(local) if world > 1 else False
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, N, S = args.members, args.elements, args.steps_per_launch
    e, beam, x0 = build_ensemble(rank, B, N, dev)
    h = e.h
    X = torch.from_numpy(x0).to(dev)
    x_host = torch.from_numpy(x0).pin_memory()
    system = beam.make_system(B)
    chunks = lambda k: [min(S, k - i) for i in range(0, k, S)]  # noqa: E731
    timed_chunks = chunks(args.steps)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- clocks: the sampler needs ~0.3 s to start; the GPU is kept under load meanwhile (same kernel on a scratch
    # copy of the state) so that the timed region does not begin on an idle, down-clocked device ----
    sampler = ClockSampler(local)
    sampler.start()
    t_spin = time.perf_counter()
    scratch = X.clone()
    # the K timed steps as ONE CUDA graph (kernel launches + ticket-counter resets).  Capturing enqueues nothing, so
    # the ensemble is not advanced by it; one eager pass on the scratch copy first, so that no lazy module loading
    # or attribute set-up happens inside the capture
    side = torch.cuda.Stream(device=dev)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        for c in sorted(set(timed_chunks)):
            rk4_steps(beam, scratch, 0.0, h, c, system=system)
    side.synchronize()
    with torch.cuda.graph(graph, stream=side):
        tk = 0
        for c in timed_chunks:
            rk4_steps(beam, X, tk * h, h, c, system=system)  # u = 0: the system is autonomous, t0 is not used
            tk += c
    while time.perf_counter() - t_spin < 0.45:
        for _ in range(20):
            rk4_steps(beam, scratch, 0.0, h, S, system=system)
        torch.cuda.synchronize(dev)
    # ---- warm-up: W steps on the ensemble itself ----
    for c in chunks(args.warmup):
        rk4_steps(beam, X, 0.0, h, c, system=system)
    # ---- timed region: EXACTLY K steps, state resident in HBM ----
    barrier()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    graph.replay()
    stop.record()
    barrier()
    elapsed_ms = start.elapsed_time(stop)
    # per-launch durations of the dominant kernel (roofline): the same launches, event-bracketed one by one
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in timed_chunks]
    for (a, b), c in zip(ev, timed_chunks):
        a.record()
        rk4_steps(beam, scratch, 0.0, h, c, system=system)
        b.record()
    torch.cuda.synchronize(dev)
    clocks = sampler.window(t_spin, time.perf_counter())
    launch_ms = [a.elapsed_time(b) for a, b in ev]
    full = [ms for ms, c in zip(launch_ms, timed_chunks) if c == timed_chunks[0]]
    full_steps = timed_chunks[0]
    finite = bool(torch.isfinite(X).all().item())
    del scratch

    # ---- e2e: host buffers through the public host-pipeline API: every call copies the state from pinned host
    # memory, runs its fused steps and copies the state back (chunked: copies overlap the kernels on three streams) ----
    pipe = HostPipeline(beam, B, chunk_members=args.e2e_chunk_members)
    for _ in range(3):
        pipe.run(x_host, 0.0, h, timed_chunks[0])
    pipe.synchronize()
    torch.cuda.synchronize(dev)
    x_host.copy_(torch.from_numpy(x0))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    te = 0
    for c in timed_chunks:
        pipe.run(x_host, te * h, h, c)
        te += c
    pipe.wait()  # the current stream waits for the last copy-out before the closing event
    e1.record()
    barrier()
    pipe.synchronize()
    e2e_ms = e0.elapsed_time(e1)
    e2e_ok = bool(np.isfinite(x_host.numpy()).all())
    e2e_calls = len(timed_chunks)

    # ---- final gather (the only collective of the job; outside the step path) ----
    gather_ms = None
    if world > 1:
        tip = X[:, beam.n_free - 2].contiguous()
        out = torch.empty(world * B, dtype=tip.dtype, device=dev) if rank == 0 else None
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.gather(tip, list(out.chunk(world)) if rank == 0 else None, dst=0)  # connection set-up, untimed
        barrier()
        g0.record()
        dist.gather(tip, list(out.chunk(world)) if rank == 0 else None, dst=0)
        g1.record()
        torch.cuda.synchronize(dev)
        gather_ms = g0.elapsed_time(g1)
        tt = torch.tensor([elapsed_ms, e2e_ms, float(np.mean(full))], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        elapsed_ms, e2e_ms, mean_launch = (float(v) for v in tt.tolist())
    else:
        mean_launch = float(np.mean(full))

    if rank == 0:
        total_units = float(world) * B * N * args.steps
        value = total_units / (elapsed_ms * 1e-3)
        e2e_val = total_units / (e2e_ms * 1e-3)
        peak, which = measured_peak()
        launch_units = float(B) * N * full_steps
        achieved = ALG_BYTES_PER_ELEMENT_STEP * launch_units / (mean_launch * 1e-3) / 1e9
        fp64_peak = probe_fp64_peak(dev)
        traffic, fp64, dram = None, None, None
        tf = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tf):
            try:
                tj = json.load(open(tf))
                if N == tj.get("elements") and B == tj.get("members"):
                    # a launch reads state + stiffness coefficients once and writes the state once whatever the number
                    # of fused steps: the ncu figure of the captured launch carries over to this one
                    traffic = tj.get("dram_bytes_per_launch")
                    dram = {"bytes_per_launch": traffic, "gbs": traffic / (mean_launch * 1e-3) / 1e9,
                            "frac_of_hbm_peak": traffic / (mean_launch * 1e-3) / 1e9 / peak, "captured_on": tj.get("captured_on")}
                    tfl = tj["fp64_flops_per_element_step"] * launch_units / (mean_launch * 1e-3) / 1e12
                    fp64 = {"achieved": tfl, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tfl / fp64_peak,
                            "flops_per_element_step": tj["fp64_flops_per_element_step"],
                            "note": "flops counted by ncu for this kernel (profiles/traffic.json); peak = DFMA rate measured "
                                    "in this run (crb_probe_dfma)"}
            except Exception:
                traffic, fp64, dram = None, None, None
        state_bytes = X.numel() * 8
        # plain-copy ceiling of the box for this many GPUs (benchmarks/host_ceiling.py, profiles/r2_host_ceiling_8gpu.json)
        ceiling = None
        try:
            hc = json.load(open(os.path.join(ROOT, "profiles", "r2_host_ceiling_8gpu.json")))
            ceiling = hc["duplex_aggregate_gbs_by_gpus"].get(str(world))
        except Exception:
            pass
        host_gbs = 2.0 * state_bytes * e2e_calls * world / (e2e_ms * 1e-3) / 1e9
        line = {
            "metric": "beam-element RK4 steps/sec", "value": value, "unit": "element-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": WORKLOAD, "members_per_gpu": B, "elements": N, "h": h, "steps_per_launch": S,
                "launches": timed_chunks, "launch": "CUDA graph of the K steps' launches, replayed once in the timed region",
                "l2": "inputs larger than L2: state 100.7 MB (read+written) + per-member stiffness coefficients 67 MB per launch",
                "finite": finite,
            },
            "roofline": {
                "bound": "fp64", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": which, "fp64": fp64, "dram": dram,
                "note": "achieved / peak / frac are BASELINE's figure of merit: ALGORITHMIC bytes (96 B per element-step) per "
                        "launch / mean launch duration vs the measured HBM copy bandwidth.  A launch fuses its steps with the "
                        "state in registers, so real DRAM traffic is `dram` (a few % of peak) and the binding resource is the "
                        "FP64 pipe: `fp64` is the honest roofline fraction (see DESIGN.md 4, profiles/)",
            },
            "e2e": {"value": e2e_val, "unit": "element-steps/s",
                    "h2d_bytes_per_step": state_bytes * e2e_calls / args.steps,
                    "d2h_bytes_per_step": state_bytes * e2e_calls / args.steps,
                    "call": f"HostPipeline.run x {e2e_calls} ({timed_chunks} fused RK4 steps per call): each call H2D state "
                            f"{state_bytes} B (pinned) + kernels + D2H state, chunks of {pipe.chunk_members} members on 3 "
                            f"streams, native pipeline crb_rk4_host",
                    "host_traffic_gbs": host_gbs, "ceiling_gbs": ceiling,
                    "ceiling_note": "aggregate H2D + D2H bytes per second of this leg vs the plain-copy duplex ceiling measured on "
                                    "an 8 x B200 box of this pool with the same number of GPUs copying at once (no kernels)",
                    "finite": e2e_ok},
            "gpu_launches": len(timed_chunks),
            "numa_bound": numa_bound,
            "clocks": clocks,
            "kernel_ms_per_launch": mean_launch,
        }
        if gather_ms is not None:
            line["final_gather_ms"] = gather_ms
        if world == 1 and not args.no_secondary:
            line["secondary"] = secondary_measurements()
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(args.cpu_ref_steps, args.cpu_steps, n_elements=N)
        sampler.stop()
        print(json.dumps(line), file=_OUT, flush=True)
    else:
        sampler.stop()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10000)
    ap.add_argument("--warmup", type=int, default=500)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--members", type=int, default=65536)
    ap.add_argument("--elements", type=int, default=32)
    ap.add_argument("--steps-per-launch", type=int, default=50)
    ap.add_argument("--cpu-steps", type=int, default=6000, help="RK4 steps per member in the oracle-port sample")
    ap.add_argument("--cpu-ref-steps", type=int, default=1000, help="RK4 steps per member in the reference sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--e2e-chunk-members", type=int, default=0, help="members per pipelined chunk (0 = two kernel waves)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    # stdout carries the ONE JSON line and nothing else: libraries that write to file descriptor 1 (NCCL prints its
    # version banner there under torchrun) are sent to stderr, the JSON line goes to the saved descriptor
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
