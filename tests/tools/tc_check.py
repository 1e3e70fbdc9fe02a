"""Stand-alone check of the tcgen05 SDF kernel against the oracle (run under `timeout` on the GPU box)."""
import sys
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
import torch
from oracle import sdf_oracle as so
from gpu_util import kink_mask, to_weights
from nlotrajectories_b200.sdf import LearnedSDF

def run(net, n, prec, label):
    model = LearnedSDF(to_weights(net), precision=prec)
    rng = np.random.default_rng(5)
    P = rng.uniform(-0.5, 1.5, (n, 2)).astype(np.float32)
    sb = rng.uniform(0.5, 1.5, n).astype(np.float32)
    x = torch.from_numpy(P[:, 0].copy()).cuda(); y = torch.from_numpy(P[:, 1].copy()).cuda()
    s, jx, jy = model.eval(x, y)
    torch.cuda.synchronize()
    s_ref, J_ref = so.value_jac(net.astype(np.float64), P.astype(np.float64))
    tie = kink_mask(net, P)
    J = np.stack([jx.cpu().numpy(), jy.cpu().numpy()], 1)
    es = np.abs(s.cpu().numpy() - s_ref) / np.maximum(1, np.abs(s_ref))
    ej = (np.abs(J - J_ref) / np.maximum(1, np.abs(J_ref))).max(axis=1)
    print(f"{label:28s} prec={model.precision:9s} n={n:8d} max rel err s={es.max():.3e} J(no ties)={ej[~tie].max():.3e} ties={tie.sum()} "
          f"J>1e-5: {(ej[~tie] > 1e-5).sum()}", flush=True)
    model.close()

nets = {"relu128": so.synthetic_mlp(128, 1, seed=0), "relu64": so.synthetic_mlp(64, 1, seed=1),
        "shipped_fourier128": so.from_npz(REPO / "tests/golden/sdf_shipped_fourier128_weights.npz"),
        "tanh64": so.synthetic_mlp(64, 1, seed=4, act=so.ACT_TANH)}
for name, net in nets.items():
    for n in (128, 1000, 100003):
        for prec in ("tc3xf16", "fp32"):
            run(net, n, prec, name)
# timing
net = nets["relu128"]
for prec in ("tc3xf16", "fp32"):
    model = LearnedSDF(to_weights(net), precision=prec)
    n = 1 << 22
    x = torch.rand(n, device="cuda") * 2 - 0.5; y = torch.rand(n, device="cuda") * 2 - 0.5
    out = (torch.empty_like(x), torch.empty_like(x), torch.empty_like(x))
    for _ in range(3): model.eval(x, y, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): model.eval(x, y, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"timing {prec}: {n / ms / 1e6:.3f} G pts/s ({ms:.3f} ms for 2^22)", flush=True)
