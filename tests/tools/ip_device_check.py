"""The device interior-point solver at BASELINE batch sizes (benchmark_4 x 4,096 on one GPU, benchmark_6 x 65,536): solves,
converged fraction, best objective, time, and where an iteration goes (CUDA events around each phase are not needed: the solver
is one C call; the per-kernel split comes from the ncu launch list of this command).

    python tests/tools/ip_device_check.py benchmark_4 4096 [--init multistart|rrt|rrt_lift] [--plans 512] [--max-iter 300] [--compare 64]

--init rrt: `--plans` seeded RRT plans on the host cores (core/trajectory_initialization.py:175-236 restated in initializer.py),
tiled over the batch with a small seeded jitter on the path (N(0, 0.005)) so that every start is distinct.
--compare n: also solve the first n starts with the dense torch solver (BatchedIPSolver) and report the objective differences.
Prints one JSON line."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("name"); ap.add_argument("P", type=int)
    ap.add_argument("--init", default="multistart"); ap.add_argument("--plans", type=int, default=512)
    ap.add_argument("--max-iter", type=int, default=300); ap.add_argument("--compare", type=int, default=0)
    ap.add_argument("--weights", default=None); ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--repeat", type=int, default=1)
    args = ap.parse_args()
    from nlotrajectories_b200.config import Config
    cfg = Config.load(next((REPO / "nlotrajectories_b200/benchmarks").glob(args.name + "*.yaml")))
    w0 = None
    t_init = 0.0
    if args.init.startswith("rrt"):                                          # host processes fork before the CUDA context exists
        from nlotrajectories_b200.initializer import rrt_multistart
        t0 = time.time()
        plans = rrt_multistart(cfg, min(args.plans, args.P), lift=args.init == "rrt_lift").astype(np.float64)
        t_init = time.time() - t0
    import torch
    from gpu_util import to_weights
    from nlotrajectories_b200.problem import NlpProblem
    from nlotrajectories_b200.sdf import LearnedSDF
    from nlotrajectories_b200.solver import BatchedIPSolver, DeviceEvaluator, DeviceIPSolver
    from oracle import sdf_oracle as so
    model = None
    if cfg.solver.mode == "l4casadi":
        wpath = args.weights or str(REPO / "tests" / "golden" / f"sdf_{args.name}_relu128.npz")
        model = LearnedSDF(to_weights(so.from_npz(wpath)))
    prob = NlpProblem.from_config(cfg, model)
    P = args.P
    if args.init.startswith("rrt"):
        reps = (P + len(plans) - 1) // len(plans)
        w0 = np.tile(plans, (reps, 1))[:P].copy()
        rng = np.random.default_rng(7)
        jitter = rng.normal(0.0, 0.005, (P, prob.N + 1, 2))
        jitter[:, 0] = 0.0; jitter[:, -1] = 0.0
        jitter[:len(plans)] = 0.0                                             # the plans themselves stay as planned
        X = w0[:, :prob.n_X].reshape(P, prob.N + 1, prob.nx)
        X[:, :, :2] += jitter
    else:
        w0 = prob.multistart_guess(P).astype(np.float64)
    solver = DeviceIPSolver(prob, max_problems=P, max_iter=args.max_iter, verbose=args.verbose)
    times = []
    for _ in range(args.repeat):
        torch.cuda.synchronize()
        t0 = time.time()
        res = solver.solve(w0)
        times.append(time.time() - t0)
    dt = min(times)
    f = res.f.numpy(); v = res.violation.numpy(); ok = res.converged.numpy(); st = res.stalled.numpy(); its = res.iterations.numpy()
    usable = ok | (st & (v <= 1e-4))
    best = int(np.where(usable, f, np.inf).argmin())
    line = {"benchmark": args.name, "starts": P, "init": args.init, "plans": (len(plans) if args.init.startswith("rrt") else None),
            "init_s": t_init, "solve_s": dt, "solve_s_all": times, "solves_per_s": P / dt, "max_iter": args.max_iter, "stats": solver.stats,
            "ms_per_iteration": 1e3 * dt / max(1, solver.stats["iterations"]),
            "converged_frac": float(ok.mean()), "stalled_feasible_frac": float((st & (v <= 1e-4)).mean()), "usable_frac": float(usable.mean()),
            "iterations_median_converged": float(np.median(its[ok])) if ok.any() else None,
            "objective_converged_min_med_max": [float(f[ok].min()), float(np.median(f[ok])), float(f[ok].max())] if ok.any() else None,
            "best": {"index": best, "objective": float(f[best]), "violation": float(v[best])} if usable.any() else None}
    if args.compare:
        n = min(args.compare, P)
        lb, ub = prob.bounds()
        t0 = time.time()
        ref = BatchedIPSolver(DeviceEvaluator(prob), lb, ub, max_iter=args.max_iter).solve(torch.from_numpy(w0[:n]).cuda())
        torch.cuda.synchronize()
        rt = time.time() - t0
        rok = ref.converged.cpu().numpy(); rf = ref.f.cpu().numpy()
        both = ok[:n] & rok
        d = np.abs(f[:n] - rf)
        line["dense_reference"] = {"starts": n, "solve_s": rt, "converged_frac": float(rok.mean()), "device_converged_frac_same_starts": float(ok[:n].mean()),
                                   "common_converged": int(both.sum()), "max_abs_objective_difference_common": float(d[both].max()) if both.any() else None,
                                   "median_abs_objective_difference_common": float(np.median(d[both])) if both.any() else None,
                                   "frac_common_within_1e-4": float((d[both] <= 1e-4).mean()) if both.any() else None}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
