"""Planning time of the multi-start RRT guesses: forked host planners vs the batched device tree search."""
import sys, time
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO))
from nlotrajectories_b200.config import Config
from nlotrajectories_b200.initializer import rrt_multistart
from nlotrajectories_b200.train import scene_sdf
name = sys.argv[1] if len(sys.argv) > 1 else "benchmark_6"
sizes = [int(a) for a in sys.argv[2:]] or [64, 1024]
cfg = Config.load(next((REPO / "nlotrajectories_b200/benchmarks").glob(name + "*.yaml")))
host = {}
for P in sizes:
    if P <= 1024:
        t0 = time.time(); rrt_multistart(cfg, P); host[P] = time.time() - t0
import torch
from nlotrajectories_b200.rrt_device import rrt_multistart_device
rrt_multistart_device(cfg, 4)                                  # warm-up
exact = scene_sdf(cfg)
for P in sizes:
    torch.cuda.synchronize(); t0 = time.time()
    w = rrt_multistart_device(cfg, P)
    dt = time.time() - t0
    N = cfg.solver.N; nx = len(cfg.body.start_state)
    X = w[:, :nx * (N + 1)].reshape(P, N + 1, nx).astype(float)
    line = np.linspace(np.asarray(cfg.body.start_state, float), np.asarray(cfg.body.goal_state, float), N + 1)[:, :2]
    fallbacks = int(sum(np.allclose(X[i, :, :2], line, atol=1e-6) for i in range(P)))
    print(f"{name}: P={P}: device-batched trees + host post-processing {dt:.2f} s" + (f", forked host planners {host[P]:.2f} s" if P in host else "")
          + f"; min exact SDF along the splines {exact(X[..., 0], X[..., 1]).min():.4f}; straight-line fallbacks {fallbacks}", flush=True)
