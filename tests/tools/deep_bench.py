"""Throughput of the learned-SDF value + Jacobian kernel on networks with two and three H x H matrices (the reference's deeper
defaults: core/config.py:203-212, scripts/run_benchmark.py:64-83): tensor-tile path (deep kernel) vs the FP32 general path.
Prints one JSON line per network."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
from gpu_util import to_weights                              # noqa: E402
from nlotrajectories_b200.sdf import LearnedSDF              # noqa: E402
from oracle import sdf_oracle as so                          # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
nets = {
    "relu128x1": so.synthetic_mlp(128, 1, seed=0),
    "shipped_fourier128x1": so.from_npz(REPO / "tests" / "golden" / "sdf_shipped_fourier128_weights.npz"),
    "siren128x1": so.synthetic_siren(128, 1, omega0=30.0, seed=40),
    "relu128x2": so.synthetic_mlp(128, 2, seed=21),
    "relu128x3": so.synthetic_mlp(128, 3, seed=28),
    "relu64x2": so.synthetic_mlp(64, 2, seed=22),
    "fourier128x2_relu": so.synthetic_fourier(128, 2, scale=2.0, seed=23),
    "siren128x2": so.synthetic_siren(128, 2, omega0=30.0, seed=24),
    "tanh128x2": so.synthetic_mlp(128, 2, seed=25, act=so.ACT_TANH),
    "siren64x3": so.synthetic_siren(64, 3, omega0=30.0, seed=31),
}
x = torch.rand(n, device="cuda") * 2 - 0.5
y = torch.rand(n, device="cuda") * 2 - 0.5
out = tuple(torch.empty_like(x) for _ in range(3))
for name, net in nets.items():
    rec = {"net": name, "points": n, "flop_per_point": 4 * (3 * net.W0.shape[0] + len(net.hidden) * net.W0.shape[0] ** 2)}
    for prec in ("auto", "fp32"):
        m = LearnedSDF(to_weights(net), precision=prec)
        reps = 5 if prec == "auto" else 2
        for _ in range(2):
            m.eval(x, y, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            m.eval(x, y, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        rec[f"{m.precision}_gpts_per_s"] = n / ms / 1e6
        rec[f"{m.precision}_tflops"] = rec["flop_per_point"] * n / ms / 1e9
        m.close()
    print(json.dumps(rec), flush=True)
