"""The batched interior point on the fp64 oracle's functions (CPU, no GPU needed): solver development and diagnosis.
usage: ip_cpu.py <benchmark> <P> [max_iter] [init: multistart|rrt|rrt_lift] [weights.npz]"""
import sys, time
from pathlib import Path
import numpy as np, yaml
REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
import torch
from oracle import nlp_oracle as no, sdf_oracle as so
from solver_util import OracleEvaluator
from nlotrajectories_b200.config import Config
from nlotrajectories_b200.solver import BatchedIPSolver

name = sys.argv[1]; P = int(sys.argv[2]); max_iter = int(sys.argv[3]) if len(sys.argv) > 3 else 300
init = sys.argv[4] if len(sys.argv) > 4 else "multistart"
weights = sys.argv[5] if len(sys.argv) > 5 else None
ypath = next((REPO / "nlotrajectories_b200/benchmarks").glob(name + "*.yaml"))
cfg = Config.load(ypath)
spec = no.NlpSpec.from_yaml_dict(yaml.safe_load(open(ypath)))
net = so.from_npz(weights) if weights else None
lb, ub = no.bounds(spec) if hasattr(no, "bounds") else (None, None)
if init.startswith("rrt"):
    from nlotrajectories_b200.initializer import rrt_multistart
    w0 = rrt_multistart(cfg, P, lift=init == "rrt_lift").astype(np.float64)
else:
    w0 = no.multistart_guess(spec, P).astype(np.float64) if hasattr(no, "multistart_guess") else None
ev = OracleEvaluator(spec, net)
w0 = torch.from_numpy(w0)
import os
if os.environ.get("ELASTIC"):
    from nlotrajectories_b200.solver import ElasticEvaluator
    ev = ElasticEvaluator(ev, lb, ub, penalty=float(os.environ["ELASTIC"]))
    lb, ub = ev.lbg, ev.ubg
    w0 = ev.initial(w0)
t0 = time.time()
res = BatchedIPSolver(ev, lb, ub, verbose=int(sys.argv[6]) if len(sys.argv) > 6 else 1, max_iter=max_iter).solve(w0)
ok = res.converged.numpy(); stl = res.stalled.numpy(); its = res.iterations.numpy(); f = res.f.numpy()
if os.environ.get("ELASTIC"):
    print("elastic variables max per start", ev.split(res.w)[1].amax(1).numpy())
print(f"{name}: P={P} {time.time() - t0:.1f} s, evals {ev.evals}; converged {ok.mean() * 100:.1f}% stalled-feasible {stl.mean() * 100:.1f}% "
      f"iterations median {np.median(its[ok]) if ok.any() else -1:.0f} max {its[ok].max() if ok.any() else -1}; f over converged min/med/max "
      + (f"{f[ok].min():.6f} / {np.median(f[ok]):.6f} / {f[ok].max():.6f}" if ok.any() else "-") + f"; viol max {res.violation.numpy().max():.1e}")
