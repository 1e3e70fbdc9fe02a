"""Batched augmented-Lagrangian solver vs scipy SLSQP on the CPU oracle (same start, same functions)."""
import sys, time
from pathlib import Path
import numpy as np, yaml
REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
import torch
from scipy.optimize import minimize
from oracle import nlp_oracle as no, sdf_oracle as so
from gpu_util import to_weights
from nlotrajectories_b200.config import Config
from nlotrajectories_b200.problem import NlpProblem
from nlotrajectories_b200.sdf import LearnedSDF
from nlotrajectories_b200.solver import BatchedALSolver

name = sys.argv[1] if len(sys.argv) > 1 else "benchmark_1"
P = int(sys.argv[2]) if len(sys.argv) > 2 else 256
n_ref = int(sys.argv[3]) if len(sys.argv) > 3 else 2
ypath = next((REPO / "nlotrajectories_b200/benchmarks").glob(name + "*.yaml"))
cfg = Config.load(ypath)
spec = no.NlpSpec.from_yaml_dict(yaml.safe_load(open(ypath)))
net = so.from_npz(REPO / "tests/golden/sdf_shipped_fourier128_weights.npz")
model = LearnedSDF(to_weights(net)) if cfg.solver.mode == "l4casadi" else None
prob = NlpProblem.from_config(cfg, model)
w0 = prob.multistart_guess(P)
wd = torch.from_numpy(w0).cuda().T.contiguous()
solver = BatchedALSolver(prob, verbose=True)
t0 = time.time()
res = solver.solve(wd)
torch.cuda.synchronize()
dt = time.time() - t0
f = res.f.cpu().numpy(); v = res.violation.cpu().numpy(); st = res.stationarity.cpu().numpy()
ok = (v < 1e-4)
print(f"{name}: P={P} solved in {dt:.2f} s, {res.evaluations} batched evaluations; feasible(1e-4) {ok.mean()*100:.1f}% ; f min/med/max over feasible "
      f"{f[ok].min() if ok.any() else float('nan'):.5f} / {np.median(f[ok]) if ok.any() else float('nan'):.5f} / {f[ok].max() if ok.any() else float('nan'):.5f}; stationarity med {np.median(st):.2e}")
# CPU reference solve of the first n_ref starts with SLSQP on the fp64 oracle
n64 = net.astype(np.float64)
sdf = (lambda Q: so.value_jac(n64, Q)) if model else None
rows, cols, _ = no.jac_pattern(spec)
lb, ub = no.bounds(spec)
eq = lb == ub
def G(w): return no.eval_g_jac(spec, w[None], sdf)[0][0]
def J(w):
    jv = no.eval_g_jac(spec, w[None], sdf)[1][0]
    D = np.zeros((spec.n_g, spec.n_w)); D[rows, cols] = jv; return D
cons = [{"type": "eq", "fun": lambda w: G(w)[eq] - lb[eq], "jac": lambda w: J(w)[eq]}]
fin_l = (~eq) & np.isfinite(lb); fin_u = (~eq) & np.isfinite(ub)
cons.append({"type": "ineq", "fun": lambda w: G(w)[fin_l] - lb[fin_l], "jac": lambda w: J(w)[fin_l]})
if fin_u.any(): cons.append({"type": "ineq", "fun": lambda w: ub[fin_u] - G(w)[fin_u], "jac": lambda w: -J(w)[fin_u]})
ws = res.w.T.cpu().numpy()
for i in range(n_ref):
    t0 = time.time()
    r = minimize(lambda w: no.eval_f_grad(spec, w[None])[0][0], w0[i].astype(np.float64), jac=lambda w: no.eval_f_grad(spec, w[None])[1][0],
                 constraints=cons, method="SLSQP", options={"maxiter": 400, "ftol": 1e-10})
    gi = G(r.x); vi = np.maximum(0, np.maximum(lb - gi, gi - ub)).max()
    # also polish the GPU solution with SLSQP started at it (same basin by construction)
    r2 = minimize(lambda w: no.eval_f_grad(spec, w[None])[0][0], ws[i].astype(np.float64), jac=lambda w: no.eval_f_grad(spec, w[None])[1][0],
                  constraints=cons, method="SLSQP", options={"maxiter": 400, "ftol": 1e-12})
    print(f"start {i}: GPU AL f={f[i]:.6f} viol={v[i]:.1e} | SLSQP from same start f={r.fun:.6f} viol={vi:.1e} ({r.nit} its, {time.time()-t0:.1f}s, {r.message}) "
          f"| SLSQP polished from GPU solution f={r2.fun:.6f} (|df|={abs(r2.fun - f[i]):.2e}, |dw|max={np.abs(r2.x - ws[i]).max():.2e})")
