"""Batched interior-point solve on the GPU evaluation path vs CPU solves of the same problems on the fp64 oracle
(the same algorithm on the oracle's functions, and scipy SLSQP from the same start)."""
import sys, time
from pathlib import Path
import numpy as np, yaml
REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
import torch
from scipy.optimize import minimize
from oracle import nlp_oracle as no, sdf_oracle as so
from gpu_util import to_weights
from solver_util import OracleEvaluator
from nlotrajectories_b200.config import Config
from nlotrajectories_b200.problem import NlpProblem
from nlotrajectories_b200.sdf import LearnedSDF, SdfWeights
from nlotrajectories_b200.solver import BatchedIPSolver, DeviceEvaluator

name = sys.argv[1] if len(sys.argv) > 1 else "benchmark_1"
P = int(sys.argv[2]) if len(sys.argv) > 2 else 256
n_ref = int(sys.argv[3]) if len(sys.argv) > 3 else 2
weights = sys.argv[4] if len(sys.argv) > 4 and sys.argv[4] != "-" else None
init = sys.argv[5] if len(sys.argv) > 5 else "multistart"          # "multistart" (seeded lateral offsets) | "rrt" (seeded planner per start) | "rrt_lift"
ypath = next((REPO / "nlotrajectories_b200/benchmarks").glob(name + "*.yaml"))
cfg = Config.load(ypath)
spec = no.NlpSpec.from_yaml_dict(yaml.safe_load(open(ypath)))
net = so.from_npz(weights) if weights else None
w0 = None
if init.startswith("rrt"):                                          # host processes fork before the CUDA context exists
    from nlotrajectories_b200.initializer import rrt_multistart
    t0 = time.time()
    w0 = rrt_multistart(cfg, P, lift=init == "rrt_lift").astype(np.float64)
    print(f"{name}: {P} RRT initial guesses in {time.time() - t0:.1f} s (host)", flush=True)
model = LearnedSDF(to_weights(net)) if cfg.solver.mode == "l4casadi" else None
prob = NlpProblem.from_config(cfg, model)
lb, ub = prob.bounds()
if w0 is None:
    w0 = prob.multistart_guess(P).astype(np.float64)
ev = DeviceEvaluator(prob)
t0 = time.time()
elastic = float(sys.argv[6]) if len(sys.argv) > 6 else 0.0          # > 0: elastic mode with this penalty
if elastic > 0:
    from nlotrajectories_b200.solver import solve_elastic
    res = solve_elastic(ev, lb, ub, torch.from_numpy(w0).cuda(), penalty=elastic, verbose=True, max_iter=300)
else:
    res = BatchedIPSolver(ev, lb, ub, verbose=True, max_iter=300).solve(torch.from_numpy(w0).cuda())
torch.cuda.synchronize()
dt = time.time() - t0
f = res.f.cpu().numpy(); v = res.violation.cpu().numpy(); ok = res.converged.cpu().numpy(); its = res.iterations.cpu().numpy()
stl = res.stalled.cpu().numpy()
use = ok | (stl & (v <= 1e-4))
best = np.where(use, f, np.inf).argmin()
q = lambda m: f"{f[m].min():.6f} / {np.median(f[m]):.6f} / {f[m].max():.6f}" if m.any() else "-"
print(f"{name}: P={P} batched IP on GPU: {dt:.1f} s, {ev.evals} batched evaluations; converged {ok.mean() * 100:.1f}% (iterations median {np.median(its[ok]) if ok.any() else -1:.0f}), "
      f"stalled-feasible {stl.mean() * 100:.1f}%; f over converged min/med/max {q(ok)}; over stalled-feasible {q(stl)}; best start {best} f={f[best]:.6f} viol={v[best]:.1e}", flush=True)
if n_ref == 0:
    sys.exit(0)
# CPU: same algorithm on the fp64 oracle, and SLSQP, for the first n_ref starts
cpu = BatchedIPSolver(OracleEvaluator(spec, net), lb, ub, max_iter=300).solve(torch.from_numpy(w0[:n_ref]))
n64 = net.astype(np.float64) if net is not None else None
sdf = (lambda Q: so.value_jac(n64, Q)) if model else None
rows, cols, _ = no.jac_pattern(spec)
eq = lb == ub
def G(w): return no.eval_g_jac(spec, w[None], sdf)[0][0]
def J(w):
    jv = no.eval_g_jac(spec, w[None], sdf)[1][0]
    D = np.zeros((spec.n_g, spec.n_w)); D[rows, cols] = jv; return D
cons = [{"type": "eq", "fun": lambda w: G(w)[eq] - lb[eq], "jac": lambda w: J(w)[eq]}]
fin_l = (~eq) & np.isfinite(lb); fin_u = (~eq) & np.isfinite(ub)
cons.append({"type": "ineq", "fun": lambda w: G(w)[fin_l] - lb[fin_l], "jac": lambda w: J(w)[fin_l]})
if fin_u.any(): cons.append({"type": "ineq", "fun": lambda w: ub[fin_u] - G(w)[fin_u], "jac": lambda w: -J(w)[fin_u]})
wg = res.w.cpu().numpy()
for i in range(n_ref):
    t0 = time.time()
    r = minimize(lambda w: no.eval_f_grad(spec, w[None])[0][0], w0[i], jac=lambda w: no.eval_f_grad(spec, w[None])[1][0],
                 constraints=cons, method="SLSQP", options={"maxiter": 400, "ftol": 1e-10})
    gi = G(r.x); vi = np.maximum(0, np.maximum(lb - gi, gi - ub)).max()
    print(f"start {i}: GPU IP f={f[i]:.6f} viol={v[i]:.1e} conv={ok[i]} stalled={stl[i]} its={its[i]} | CPU IP (oracle fp64) f={cpu.f[i].item():.6f} conv={bool(cpu.converged[i])} "
          f"|dw|max={np.abs(cpu.w[i].numpy() - wg[i]).max():.2e} | SLSQP f={r.fun:.6f} viol={vi:.1e} ({r.nit} its, {time.time() - t0:.1f}s) |dw|max={np.abs(r.x - wg[i]).max():.2e}", flush=True)
