"""``run-benchmark --config <yaml>`` for the batched GPU path (reference: scripts/run_benchmark.py:49-234).

Parses the reference's YAML unchanged.  Where the reference trains an SDF, builds one Opti instance and calls
IPOPT, this runner loads SDF weights, builds ``--batch`` multi-start problems, evaluates the NLP functions for the
whole batch on the GPU (sharded by problem index under torchrun) and reports the best start by merit
(objective + penalty on constraint violation).  New knobs are additive flags; the YAML schema is the reference's.
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

from .config import Config
from .distributed import init_process_group, merit, select_best, shard_range
from .problem import NlpProblem
from .sdf import LearnedSDF, SdfWeights


def synthetic_weights(cfg: Config, seed: int = 0) -> SdfWeights:
    """Seeded stand-in for the network the reference would train (scripts/run_benchmark.py:64-96): same
    architecture as the YAML's ``model`` section, W ~ N(0, 1/fan_in)."""
    from .sdf import ACT_COS_SCALE, ACT_SIN, activation_id
    m = cfg.model
    H, M = m.hidden_dim, m.n_hidden_mats()
    rng = np.random.default_rng(seed)
    W0 = rng.standard_normal((H, 2)) / np.sqrt(2.0)
    b0 = 0.1 * rng.standard_normal(H)
    hidden = [(rng.standard_normal((H, H)) / np.sqrt(H), 0.1 * rng.standard_normal(H)) for _ in range(M)]
    w_out = rng.standard_normal(H) / np.sqrt(H)
    if m.type == "mlp":
        a = activation_id(m.activation_function)
        return SdfWeights.pack("mlp", W0, b0, hidden, w_out, 0.05, a, a)
    if m.type == "fourier":
        return SdfWeights.pack("fourier", W0, b0, hidden, w_out, 0.05, ACT_COS_SCALE, activation_id(m.activation_function), 1.0)
    return SdfWeights.pack("siren", W0 / 30.0, b0, [(W / m.omega_0, b) for W, b in hidden], w_out, 0.0, ACT_SIN, ACT_SIN,
                           m.omega_0, m.omega_0)


def solve_benchmark(config_path: Path, batch: int = 64, weights: str | None = None, precision: str = "auto", verbose: bool = True,
                    lift: bool = False, elastic: float | None = None, device_rrt: bool = False, dense: bool = False, max_iter: int = 300):
    """``--solve``: every start of this rank's shard goes through the interior point - the device solver of the CUDA library
    (``DeviceIPSolver``: block-tridiagonal KKT kernels, batches of thousands of starts) or, with ``dense`` / ``elastic``, the dense
    torch solver of ``solver.py`` (a few hundred starts).  The best usable objective across ranks is selected with the same
    all-gather + broadcast as the evaluation path."""
    from .distributed import env_rank_world
    rank, local_rank, world = env_rank_world()
    cfg = Config.load(config_path)
    lo, hi = shard_range(batch, rank, world)
    w0_host = None
    if cfg.solver.initializer.mode == "rrt" and not device_rrt:
        # the YAML's initializer (scripts/run_benchmark.py:114-127), one seed per start, planned on forked host processes BEFORE this
        # process creates its CUDA context or NCCL threads (forking a process that holds either is unsupported)
        from .initializer import rrt_multistart
        w0_host = rrt_multistart(cfg, hi - lo, first=lo, lift=lift).astype(np.float64)
    post_pool = None
    if cfg.solver.initializer.mode == "rrt" and device_rrt:      # the post-processing workers fork now, while this process holds no CUDA context
        from .rrt_device import make_post_pool
        post_pool = make_post_pool(cfg, lift)
    import torch
    from .solver import BatchedIPSolver, DeviceEvaluator, DeviceIPSolver, solve_elastic
    init_process_group("nccl")
    torch.cuda.set_device(local_rank)
    if post_pool is not None:                                    # thousands of starts: the tree search batched on the GPU (rrt_device.py)
        from .rrt_device import rrt_multistart_device
        w0_host = rrt_multistart_device(cfg, hi - lo, first=lo, lift=lift, device=torch.device("cuda", local_rank), pool=post_pool).astype(np.float64)
        post_pool.close(); post_pool.join()
    sdf = None
    if cfg.solver.mode == "l4casadi":
        sdf = LearnedSDF(SdfWeights.load(weights) if weights else synthetic_weights(cfg), device=local_rank, precision=precision)
    prob = NlpProblem.from_config(cfg, sdf, device=local_rank)
    dev = torch.device("cuda", local_rank)
    if w0_host is None:
        w0_host = prob.multistart_guess(hi - lo, first=lo).astype(np.float64)
    lb, ub = prob.bounds()
    t0 = time.time()
    if elastic:
        res = solve_elastic(DeviceEvaluator(prob), lb, ub, torch.from_numpy(w0_host).to(dev), penalty=float(elastic), verbose=verbose and rank == 0,
                            max_iter=max_iter)
        name = "dense batched interior point, elastic mode"
    elif dense:
        res = BatchedIPSolver(DeviceEvaluator(prob), lb, ub, verbose=verbose and rank == 0, max_iter=max_iter).solve(torch.from_numpy(w0_host).to(dev))
        name = "dense batched interior point (torch.linalg)"
    else:
        res = DeviceIPSolver(prob, max_problems=hi - lo, max_iter=max_iter, verbose=verbose and rank == 0).solve(w0_host)
        name = "device interior point (block-tridiagonal KKT kernels)"
    torch.cuda.synchronize()
    dt = time.time() - t0
    w_res, f_res, v_res = res.w.to(dev), res.f.to(dev), res.violation.to(dev)
    usable = res.converged.to(dev) | (res.stalled.to(dev) & (v_res <= 1e-4))
    score = torch.where(usable, f_res, f_res + 1e3 * (1.0 + v_res)).float()
    best_val, best_idx, w_best = select_best(score, w_res.float().T.contiguous(), lo, prob.n_w)
    out = {"config": str(config_path), "batch": batch, "world": world, "solver": name + ", tol 1e-4, exact Hessian",
           "solve_s_rank0": dt, "solves_per_s_rank0": (hi - lo) / dt, "converged_fraction_rank0": float(res.converged.float().mean().item()),
           "stalled_feasible_fraction_rank0": float(res.stalled.float().mean().item()), "best_objective": best_val, "best_start": best_idx}
    if rank == 0 and verbose:
        print(json.dumps(out))
    return out, w_best.cpu().numpy()


def run_benchmark(config_path: Path, batch: int = 4096, weights: str | None = None, precision: str = "auto",
                  repeats: int = 3, verbose: bool = True, sdf_metrics: bool = False):
    import torch
    rank, local_rank, world = init_process_group("nccl")
    torch.cuda.set_device(local_rank)
    cfg = Config.load(config_path)
    sdf = None
    if cfg.solver.mode == "l4casadi":
        if verbose and rank == 0:
            print(f"Using model type: {cfg.model.type}")                   # scripts/run_benchmark.py:56
        w = SdfWeights.load(weights) if weights else synthetic_weights(cfg)
        if weights is None and rank == 0:
            print("[run-benchmark] no --weights given: using seeded synthetic SDF weights of the YAML's architecture", file=sys.stderr)
        sdf = LearnedSDF(w, device=local_rank, precision=precision)
    prob = NlpProblem.from_config(cfg, sdf, device=local_rank)
    lo, hi = shard_range(batch, rank, world)
    P = hi - lo
    dev = torch.device("cuda", local_rank)
    w_soa = torch.from_numpy(prob.multistart_guess(P, first=lo)).to(dev).T.contiguous()
    g, jac, f, grad = prob.alloc_outputs(P, dev)
    prob.eval_device(w_soa, g, jac, f, grad)
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(repeats):
        prob.eval_device(w_soa, g, jac, f, grad)
    torch.cuda.synchronize()
    dt = (time.time() - t0) / repeats
    lb, ub = prob.bounds()
    big = 3.0e38
    lbd = torch.from_numpy(np.clip(lb, -big, big).astype(np.float32)).to(dev)
    ubd = torch.from_numpy(np.clip(ub, -big, big).astype(np.float32)).to(dev)
    viol = prob.violation(g, lbd, ubd)
    best_val, best_idx, w_best = select_best(merit(f, viol), w_soa, lo, prob.n_w)
    out = {"config": str(config_path), "batch": batch, "world": world, "n_w": prob.n_w, "n_g": prob.n_g, "nnz_jac": prob.nnz,
           "sdf_points_per_eval": prob.n_sdf_points if sdf is not None else 0, "precision": sdf.precision if sdf else None,
           "eval_ms_rank0": dt * 1e3, "problem_evals_per_s_rank0": P / dt, "best_merit": best_val, "best_start": best_idx}
    if sdf_metrics and sdf is not None and rank == 0:              # scripts/run_benchmark.py:34-46, 203-205 (1000^2 grid)
        from .metrics import compute_metrics
        from .train import scene_sdf
        names = ("mse", "iou", "hausdorff", "chamfer", "surface_loss")
        out["sdf_metrics"] = dict(zip(names, compute_metrics(sdf, scene_sdf(cfg))))
    if rank == 0 and verbose:
        print(json.dumps(out))
    return out, w_best.cpu().numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=str, required=True, help="Path to benchmark YAML config")
    ap.add_argument("--batch", type=int, default=4096, help="number of multi-start problems (whole job)")
    ap.add_argument("--weights", type=str, default=None, help="SDF weights (.nlow, .npz, TorchScript .pt or state_dict .pt)")
    ap.add_argument("--precision", default="auto", choices=["auto", "fp32", "tc3xf16"])
    ap.add_argument("--solve", action="store_true", help="solve every start (batched interior point) instead of evaluating the initial guesses")
    ap.add_argument("--lift", action="store_true", help="--solve: fill heading / speed / steering of the RRT guesses from the planned path")
    ap.add_argument("--device-rrt", action="store_true", help="--solve: plan the RRT paths of all starts with the CUDA planner (one warp per start) instead of on host processes")
    ap.add_argument("--elastic", type=float, default=None, metavar="PENALTY", help="--solve: elastic mode (exact l1 penalty) on the inequality rows")
    ap.add_argument("--dense", action="store_true", help="--solve: the dense torch solver instead of the device solver (at most 512 starts)")
    ap.add_argument("--max-iter", type=int, default=300, help="--solve: iteration limit")
    ap.add_argument("--sdf-metrics", action="store_true", help="also report the learned SDF's quality metrics on the 1000^2 grid")
    a = ap.parse_args()
    if a.solve:
        batch = a.batch
        if (a.dense or a.elastic) and batch > 512:
            print(f"[run-benchmark] --dense / --elastic factorise dense {batch} x n_w^2 fp64 matrices: batch clamped from {batch} to 512", file=sys.stderr)
            batch = 512
        solve_benchmark(Path(a.config), batch, a.weights, a.precision, lift=a.lift, elastic=a.elastic, device_rrt=a.device_rrt, dense=a.dense,
                        max_iter=a.max_iter)
    else:
        run_benchmark(Path(a.config), a.batch, a.weights, a.precision, sdf_metrics=a.sdf_metrics)
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
