"""Small invocations of every tensor-path kernel and the NLP kernels (a quick smoke of all kernel variants on a GPU box)."""
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
nets = {"relu64": so.synthetic_mlp(64, 1, seed=1), "relu128": so.synthetic_mlp(128, 1, seed=0), "relu256": so.synthetic_mlp(256, 1, seed=2),
        "tanh128": so.synthetic_mlp(128, 1, seed=13, act=so.ACT_TANH), "fourier128": so.from_npz(REPO / "tests/golden/sdf_shipped_fourier128_weights.npz")}
n = 128 * 148 * 2 + 37
x = torch.rand(n, device="cuda") * 2 - 0.5; y = torch.rand(n, device="cuda") * 2 - 0.5
for name, net in nets.items():
    m = LearnedSDF(to_weights(net))
    s, jx, jy = m.eval(x, y)
    if name == "fourier128":
        m.hess(x, y)
    torch.cuda.synchronize()
    print(name, m.precision, float(s.sum()), flush=True)
m = LearnedSDF(to_weights(nets["relu128"]))
prob = NlpProblem.from_config(Config.load(next((REPO / "nlotrajectories_b200/benchmarks").glob("benchmark_6*.yaml"))), m)
P = 333
w = torch.from_numpy(prob.multistart_guess(P)).cuda().T.contiguous()
g, jac, f, grad = prob.alloc_outputs(P)
prob.eval_device(w, g, jac, f, grad)
prob.eval_hess_device(w, torch.randn((prob.n_g, P), device="cuda"))
torch.cuda.synchronize()
print("nlp ok", float(f.sum()), flush=True)
