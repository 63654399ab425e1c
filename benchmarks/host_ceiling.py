#!/usr/bin/env python
"""Host <-> device copy ceiling of the box, all GPUs at once (no kernels): what bounds the host-buffer `e2e` figure.

    torchrun --nproc-per-node N benchmarks/host_ceiling.py --mode ranks [--bind]     one process per GPU
    python benchmarks/host_ceiling.py --mode single --gpus N                         one process driving N devices

Every GPU copies a pinned 100.7 MB buffer (the config-3 state) host->device and device->host concurrently on two
streams, `reps` times back to back; the timed region is bracketed by a barrier.  Reports per-GPU and aggregate GB/s
(H2D + D2H bytes), max over ranks."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def copy_loop(torch, dev, host_in, host_out, d_in, d_out, reps):
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    return s1, s2, a, b


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="ranks", choices=["ranks", "single"])
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--bind", action="store_true", help="bind each rank to its GPU's NUMA node before allocating")
    ap.add_argument("--reps", type=int, default=40)
    ap.add_argument("--mb", type=float, default=100.663296)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    import torch

    nbytes = int(a.mb * 1e6) // 8 * 8
    if a.mode == "ranks":
        import torch.distributed as dist

        rank = int(os.environ.get("RANK", "0"))
        local = int(os.environ.get("LOCAL_RANK", "0"))
        world = int(os.environ.get("WORLD_SIZE", "1"))
        torch.cuda.set_device(local)
        bound = False
        if a.bind:
            from continuum_robot_b200.sharding import bind_to_gpu_numa_node

            bound = bind_to_gpu_numa_node(local)
        if world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        devs = [torch.device("cuda", local)]
    else:
        rank, world, bound = 0, 1, False
        devs = [torch.device("cuda", i) for i in range(a.gpus)]
    st = []
    for dev in devs:
        with torch.cuda.device(dev):
            hi = torch.empty(nbytes // 8, dtype=torch.float64).pin_memory()
            ho = torch.empty(nbytes // 8, dtype=torch.float64).pin_memory()
            hi.fill_(1.0)  # first touch after the binding
            ho.fill_(0.0)
            di = torch.empty(nbytes // 8, dtype=torch.float64, device=dev)
            do = torch.ones(nbytes // 8, dtype=torch.float64, device=dev)
            s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
            st.append((dev, hi, ho, di, do, s1, s2))

    def run(reps, h2d=True, d2h=True):
        evs = []
        for dev, hi, ho, di, do, s1, s2 in st:
            with torch.cuda.device(dev):
                e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                e0.record(s1)
                s2.wait_event(e0)
                for _ in range(reps):
                    if h2d:
                        with torch.cuda.stream(s1):
                            di.copy_(hi, non_blocking=True)
                    if d2h:
                        with torch.cuda.stream(s2):
                            ho.copy_(do, non_blocking=True)
                e1.record(s1)
                e2.record(s2)
                evs.append((dev, e0, e1, e2))
        ms = 0.0
        for dev, e0, e1, e2 in evs:
            torch.cuda.synchronize(dev)
            ms = max(ms, e0.elapsed_time(e1), e0.elapsed_time(e2))
        return ms

    def barrier():
        if a.mode == "ranks" and world > 1:
            import torch.distributed as dist

            dist.barrier()
        for dev, *_ in st:
            torch.cuda.synchronize(dev)

    res = {}
    for name, kw in (("h2d", dict(d2h=False)), ("d2h", dict(h2d=False)), ("duplex", {})):
        run(3, **kw)
        barrier()
        ms = run(a.reps, **kw)
        barrier()
        if a.mode == "ranks" and world > 1:
            import torch.distributed as dist

            t = torch.tensor([ms], device=devs[0], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        ngpu = world if a.mode == "ranks" else len(devs)
        dirs = 2 if name == "duplex" else 1
        res[name] = {"ms_per_rep": ms / a.reps, "per_gpu_gbs_per_direction": nbytes * a.reps / (ms * 1e-3) / 1e9,
                     "aggregate_gbs": dirs * ngpu * nbytes * a.reps / (ms * 1e-3) / 1e9}
    if rank == 0:
        out = {"mode": a.mode, "gpus": world if a.mode == "ranks" else len(devs), "numa_bound": bound, "bytes": nbytes,
               "reps": a.reps, "cpus": len(os.sched_getaffinity(0)), "results": res}
        s = json.dumps(out)
        print(s, flush=True)
        if a.out:
            with open(a.out, "a") as f:
                f.write(s + "\n")
    if a.mode == "ranks" and world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


if __name__ == "__main__":
    main()
