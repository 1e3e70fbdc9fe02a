"""What the ncu captures under profiles/ are taken on: a few device-resident steps of the bench workload (benchmark_6 x 65,536,
ReLU 2-128-128-1), one Hessian of the Lagrangian, and the learned-SDF kernel alone for the other shapes quoted in DESIGN.md
(H = 64 ReLU, shipped FourierMLP-128, 2^24 points).

    python tests/tools/profile_step.py [--problems 65536] [--points 16777216]
"""
import argparse
import sys
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))

from nlotrajectories_b200.config import Config          # noqa: E402
from nlotrajectories_b200.problem import NlpProblem     # noqa: E402
from nlotrajectories_b200.sdf import LearnedSDF, SdfWeights  # noqa: E402
from oracle import sdf_oracle as so                     # noqa: E402
from gpu_util import to_weights                         # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--problems", type=int, default=65536)
    ap.add_argument("--points", type=int, default=1 << 24)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    net = so.synthetic_mlp(128, 1, seed=0)
    model = LearnedSDF(to_weights(net))
    prob = NlpProblem.from_config(Config.load(next((REPO / "nlotrajectories_b200" / "benchmarks").glob("benchmark_6*.yaml"))), model)
    P = args.problems
    w = torch.from_numpy(prob.multistart_guess(P)).to(dev).T.contiguous()
    g, jac, f, grad = prob.alloc_outputs(P)
    for _ in range(args.steps):
        prob.eval_device(w, g, jac, f, grad)
    lam = torch.rand((prob.n_g, P), device=dev)
    prob.eval_hess_device(w, lam)
    torch.cuda.synchronize()
    x = torch.rand(args.points, device=dev) * 2 - 0.5
    y = torch.rand(args.points, device=dev) * 2 - 0.5
    out = tuple(torch.empty_like(x) for _ in range(3))
    shipped = SdfWeights.from_npz(REPO / "tests" / "golden" / "sdf_shipped_fourier128_weights.npz")
    for name, m in (("relu128", model), ("relu64", LearnedSDF(to_weights(so.synthetic_mlp(64, 1, seed=0)))), ("fourier128", LearnedSDF(shipped))):
        for _ in range(2):
            m.eval(x, y, out=out)
        torch.cuda.synchronize()
        print(name, m.precision, float(out[0][:4].sum()))
    print("profile_step done")


if __name__ == "__main__":
    main()
