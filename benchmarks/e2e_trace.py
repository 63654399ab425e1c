import os, sys, time
sys.path.insert(0, '/root/repo')
import torch
from bench import build_ensemble
from continuum_robot_b200.integrate import HostPipeline
dev = torch.device("cuda", 0)
e, beam, x0 = build_ensemble(0, 65536, 32, dev)
x_host = torch.from_numpy(x0).pin_memory()
pipe = HostPipeline(beam, 65536, chunk_members=int(sys.argv[1]) if len(sys.argv) > 1 else 0)
for _ in range(3):
    pipe.run(x_host, 0.0, e.h, int(sys.argv[2]) if len(sys.argv) > 2 else 20); pipe.synchronize()
