#!/usr/bin/env python
"""Recording modes of the fused integrators on config 3 (65,536 x 32): what trajectory output costs.

    python benchmarks/bench_recording.py [--out FILE]

For method in {RK4 paired persistent kernel, implicit midpoint}, save_every in {none, 50, 10, 1} and recording in
{full state [T,B,2n], node shapes [T,B,N], tip trace [T,B,1]}: ms per 50-step launch (CUDA events, best of 5),
element-steps/s and the bytes of frames written per second.  Full-state recording at save_every = 1 is the one mode
where the 96 B per element-step roofline binds for real (every step's state leaves the chip)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--nsteps", type=int, default=50)
    a = ap.parse_args()
    import torch

    from bench import build_ensemble, measured_peak
    from continuum_robot_b200 import output_selection
    from continuum_robot_b200.integrate import midpoint_steps, rk4_steps, with_selection

    dev = torch.device("cuda", 0)
    B, N = 65536, 32
    e, beam, x0 = build_ensemble(0, B, N, dev)
    X0 = torch.from_numpy(x0).to(dev)
    n2 = X0.shape[1]
    peak, _ = measured_peak()
    rows = []
    sels = {"full": None, "shape": output_selection(beam, "shape"), "tip": output_selection(beam, "tip")}
    for method in ("rk4", "midpoint"):
        h = e.h if method == "rk4" else 10 * e.h
        step = rk4_steps if method == "rk4" else midpoint_steps
        for se in (0, 50, 10, 1):
            for what, sel in sels.items():
                if se == 0 and what != "full":
                    continue
                X = X0.clone()
                T = a.nsteps // se if se else 0
                width = n2 if sel is None else len(sel)
                Y = torch.empty((T, B, width), dtype=torch.float64, device=dev) if T else None
                # the crb_system_t (and the lean-recording table) is built once, outside the timed calls
                system = beam.make_system(B)
                if sel is not None and T:
                    system = with_selection(beam, system, sel)
                kw = dict(Y_out=Y, save_every=se, system=system) if T else dict(system=system)
                best = 1e30
                for _ in range(6):
                    a0, b0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a0.record()
                    step(beam, X, 0.0, h, a.nsteps, **kw)
                    b0.record()
                    torch.cuda.synchronize()
                    best = min(best, a0.elapsed_time(b0))
                frame_bytes = T * B * width * 8
                rows.append({"method": method, "save_every": se or None, "recording": what if T else "none", "ms": best,
                             "element_steps_per_s": B * N * a.nsteps / (best * 1e-3),
                             "frame_bytes": frame_bytes, "frame_write_gbs": frame_bytes / (best * 1e-3) / 1e9,
                             "frame_write_frac_of_hbm_peak": frame_bytes / (best * 1e-3) / 1e9 / peak,
                             "finite": bool(torch.isfinite(X).all().item())})
                del Y
    out = {"what": "recording modes on config 3, %d fused steps per launch, one B200" % a.nsteps, "hbm_peak_gbs": peak, "rows": rows}
    s = json.dumps(out)
    print(s)
    if a.out:
        open(a.out, "w").write(s + "\n")


if __name__ == "__main__":
    main()
