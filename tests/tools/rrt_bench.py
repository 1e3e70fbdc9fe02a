"""Planning time of the multi-start RRT guesses: forked host planners vs the device tree search (one CUDA kernel, a warp per planner:
csrc/rrt_kernels.cu) with the path post-processing on a pool of host processes forked BEFORE the CUDA context exists.
Prints one JSON line per batch size.     python tests/tools/rrt_bench.py benchmark_6 64 1024 4096 16384"""
import json
import sys
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO))
from nlotrajectories_b200.config import Config                      # noqa: E402
from nlotrajectories_b200.initializer import rrt_multistart         # noqa: E402
from nlotrajectories_b200.rrt_device import make_post_pool          # noqa: E402
from nlotrajectories_b200.train import scene_sdf                    # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "benchmark_6"
sizes = [int(a) for a in sys.argv[2:]] or [64, 1024, 4096]
cfg = Config.load(next((REPO / "nlotrajectories_b200/benchmarks").glob(name + "*.yaml")))
host = {}
for P in sizes:
    if P <= 1024:
        t0 = time.time(); rrt_multistart(cfg, P); host[P] = time.time() - t0
pool = make_post_pool(cfg)                                           # fork now: no CUDA context yet
import torch                                                         # noqa: E402
from nlotrajectories_b200.rrt_device import _host_planner, cuda_rrt_paths, rrt_multistart_device   # noqa: E402

dev = torch.device("cuda", 0)
rrt_multistart_device(cfg, 4, device=dev, pool=pool)                 # warm-up (library load, context)
exact = scene_sdf(cfg)
hostp, bounds = _host_planner(cfg)
ini, b = cfg.solver.initializer, cfg.body
for P in sizes:
    t0 = time.time()
    paths = cuda_rrt_paths(cfg, [1234 + i for i in range(P)], b.start_state, b.goal_state, bounds, ini.step_size, ini.max_iter, hostp.inflation)
    t_kernel = time.time() - t0
    t0 = time.time()
    w = rrt_multistart_device(cfg, P, device=dev, pool=pool)
    dt = time.time() - t0
    N = cfg.solver.N; nx = len(cfg.body.start_state)
    X = w[:, :nx * (N + 1)].reshape(P, N + 1, nx).astype(float)
    line = np.linspace(np.asarray(cfg.body.start_state, float), np.asarray(cfg.body.goal_state, float), N + 1)[:, :2]
    fallbacks = int(sum(np.allclose(X[i, :, :2], line, atol=1e-6) for i in range(P)))
    print(json.dumps({"benchmark": name, "starts": P, "tree_search_kernel_s": t_kernel, "plans_total_s": dt, "host_planner_pool_s": host.get(P),
                      "planner_failures": int(sum(p is None for p in paths)), "straight_line_fallbacks": fallbacks,
                      "min_exact_sdf_along_splines": float(exact(X[..., 0], X[..., 1]).min())}), flush=True)
pool.close(); pool.join()
