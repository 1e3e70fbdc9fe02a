"""Time the learned-SDF kernel alone on the shipped FourierMLP-128 (value+Jacobian) and report max error vs the fp64 oracle."""
import sys
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
import torch
from oracle import sdf_oracle as so
from gpu_util import to_weights
from nlotrajectories_b200.sdf import LearnedSDF
net = so.from_npz(REPO / "tests/golden/sdf_shipped_fourier128_weights.npz")
model = LearnedSDF(to_weights(net))
n = 1 << 22
x = torch.rand(n, device="cuda") * 2 - 0.5; y = torch.rand(n, device="cuda") * 2 - 0.5
out = (torch.empty_like(x), torch.empty_like(x), torch.empty_like(x))
for _ in range(3): model.eval(x, y, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): model.eval(x, y, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
m = 20000
Q = np.stack([x[:m].cpu().numpy(), y[:m].cpu().numpy()], 1).astype(np.float64)
s_ref, J_ref = so.value_jac(net.astype(np.float64), Q)
es = np.abs(out[0][:m].cpu().numpy() - s_ref).max()
ej = np.maximum(np.abs(out[1][:m].cpu().numpy() - J_ref[:, 0]), np.abs(out[2][:m].cpu().numpy() - J_ref[:, 1]))
print(f"shipped fourier128 {model.precision}: {n / ms / 1e6:.3f} G pts/s; max|ds|={es:.2e} |dJ| median {np.median(ej):.2e} p99 {np.quantile(ej, 0.99):.2e} "
      f"max {ej.max():.2e}; points with |dJ|>1e-5: {(ej > 1e-5).sum()} of {m} (ReLU kink ties)", flush=True)
