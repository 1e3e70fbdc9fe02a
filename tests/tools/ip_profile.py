"""Where a batched interior-point iteration spends its GPU time (torch.profiler over a few iterations of B6 x P starts):
the evaluation kernels of this library vs the dense torch linear algebra of the caller."""
import sys
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
name = sys.argv[1] if len(sys.argv) > 1 else "benchmark_6"
P = int(sys.argv[2]) if len(sys.argv) > 2 else 64
weights = sys.argv[3] if len(sys.argv) > 3 else str(REPO / "tests/golden/sdf_benchmark_6_relu128.npz")
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 12
from nlotrajectories_b200.config import Config
from nlotrajectories_b200.initializer import rrt_multistart
cfg = Config.load(next((REPO / "nlotrajectories_b200/benchmarks").glob(name + "*.yaml")))
w0 = rrt_multistart(cfg, P, lift=True).astype(np.float64)
import torch
from torch.profiler import ProfilerActivity, profile
from oracle import sdf_oracle as so
from gpu_util import to_weights
from nlotrajectories_b200.problem import NlpProblem
from nlotrajectories_b200.sdf import LearnedSDF
from nlotrajectories_b200.solver import BatchedIPSolver, DeviceEvaluator
model = LearnedSDF(to_weights(so.from_npz(weights))) if cfg.solver.mode == "l4casadi" else None
prob = NlpProblem.from_config(cfg, model)
lb, ub = prob.bounds()
ev = DeviceEvaluator(prob)
BatchedIPSolver(ev, lb, ub, max_iter=3).solve(torch.from_numpy(w0).cuda())          # warm-up (cuSOLVER handles, allocator)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    BatchedIPSolver(ev, lb, ub, max_iter=iters).solve(torch.from_numpy(w0).cuda())
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=18, max_name_column_width=70))
