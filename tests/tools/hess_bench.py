"""Throughput of the Hessian-of-the-Lagrangian evaluation and of the learned-SDF Hessian."""
import json
import sys
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
import torch
from oracle import sdf_oracle as so
from gpu_util import to_weights
from nlotrajectories_b200.config import Config
from nlotrajectories_b200.problem import NlpProblem
from nlotrajectories_b200.sdf import LearnedSDF


def timed(fn, reps):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


nets = {"relu128": so.synthetic_mlp(128, 1, seed=0), "shipped_fourier128": so.from_npz(REPO / "tests/golden/sdf_shipped_fourier128_weights.npz"),
        "tanh128": so.synthetic_mlp(128, 1, seed=13, act=so.ACT_TANH), "sigmoid128": so.synthetic_mlp(128, 1, seed=15, act=so.ACT_SIGMOID),
        "siren64": so.synthetic_siren(64, 1, omega0=30.0, seed=17), "siren128": so.synthetic_siren(128, 1, omega0=30.0, seed=40),
        "tanh64x2_fp32_general": so.synthetic_mlp(64, 2, seed=4, act=so.ACT_TANH)}
for netname, net in nets.items():
    model = LearnedSDF(to_weights(net))
    slow = netname.endswith("fp32_general")            # FP32 warp-per-point kernel (sdf_simt.cu): no tensor-tile Hessian for this shape
    n = 1 << (18 if slow else 22)
    x = torch.rand(n, device="cuda") * 2 - 0.5; y = torch.rand(n, device="cuda") * 2 - 0.5
    ms = timed(lambda: model.hess(x, y), 5)
    print(json.dumps({"row": "sdf_hess", "net": netname, "precision": model.precision, "points": n, "Gpts_s": n / ms / 1e6}), flush=True)
    for name, P in (("benchmark_6", 2048 if slow else 65536), ("benchmark_4", 512 if slow else 4096)):
        cfg = Config.load(next((REPO / "nlotrajectories_b200/benchmarks").glob(name + "*.yaml")))
        prob = NlpProblem.from_config(cfg, model)
        w = torch.from_numpy(prob.multistart_guess(P)).cuda().T.contiguous()
        lam = torch.randn((prob.n_g, P), device="cuda")
        out = torch.empty((prob.nnz_hess, P), device="cuda")
        ms = timed(lambda: prob.eval_hess_device(w, lam, None, out=out), 5)
        print(json.dumps({"row": "nlp_hess", "config": name, "net": netname, "problems": P, "nnz_hess": prob.nnz_hess, "ms": ms,
                          "hess_evals_per_s": P / ms * 1e3, "sdf_points_per_s": P * prob.n_sdf_points / ms * 1e3}), flush=True)
    model.close()
