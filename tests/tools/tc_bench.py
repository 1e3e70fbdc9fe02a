"""Time the learned-SDF kernel alone (value+Jacobian, benchmark net) - used for quick perf iterations and ncu captures."""
import sys
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
import torch
from oracle import sdf_oracle as so
from gpu_util import to_weights
from nlotrajectories_b200.sdf import LearnedSDF
H = int(sys.argv[1]) if len(sys.argv) > 1 else 128
prec = sys.argv[2] if len(sys.argv) > 2 else "tc3xf16"
logn = int(sys.argv[3]) if len(sys.argv) > 3 else 24
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
act = {"relu": so.ACT_RELU, "tanh": so.ACT_TANH, "sigmoid": so.ACT_SIGMOID, "leaky": so.ACT_LEAKY_RELU}[sys.argv[5] if len(sys.argv) > 5 else "relu"]
net = so.synthetic_mlp(H, 1, seed=0, act=act)
model = LearnedSDF(to_weights(net), precision=prec)
n = 1 << logn
x = torch.rand(n, device="cuda") * 2 - 0.5; y = torch.rand(n, device="cuda") * 2 - 0.5
out = (torch.empty_like(x), torch.empty_like(x), torch.empty_like(x))
for _ in range(3): model.eval(x, y, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): model.eval(x, y, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"H={H} {prec}: {n / ms / 1e6:.3f} G pts/s ({ms:.3f} ms for 2^{logn}) -> {4*(3*H+H*H)*n/ms/1e9:.1f} TFLOP/s algorithmic", flush=True)
