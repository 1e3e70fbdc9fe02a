"""Benchmark of the NLP-evaluation hot path (contract in the task statement).

    python bench.py --gpus N --steps K --warmup W                 # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W # the CPU reference arm

Workload (BASELINE.json configs[3], the config the metric is quoted on): benchmark_6_ackermann_wave
with 65,536 multi-start problems IN TOTAL, sharded by problem index over the N GPUs (strong scaling:
rank r owns [r*65536/N, (r+1)*65536/N)); with N > 1 the same run also measures the
weak-scaling variant (65,536 starts per GPU) and reports it in a "weak" block beside the headline
(`--scaling weak` makes that the headline instead).  One step = one pass of the hot path over the batch:
g(w), all structural non-zeros of dg/dw, f(w), grad f(w) -> 324 learned-SDF value+Jacobian points per
problem.  The SDF network is the YAML's model (mlp, ReLU, 2->128->128->1) with seeded synthetic weights.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

if "reference" in sys.argv[1:]:
    # the CPU arm uses every host core at every N: torch.distributed.run exports OMP_NUM_THREADS=1 for N > 1, and the BLAS / OpenMP
    # runtimes read these variables when numpy / torch are first imported
    for _v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

METRIC = "learned-SDF value+Jacobian points/s"
UNIT = "points/s"
YAML = "benchmark_6_ackermann_wave.yaml"


def load_peaks():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for ln in self.lines:
            f = [c.strip() for c in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def synthetic_net():
    """The YAML's model: l4c.naive.MultiLayerPerceptron(2, 128, 1, 2, 'ReLU') -> 2->128->128->1, seeded weights
    (W ~ N(0, 1/fan_in))."""
    rng = np.random.default_rng(0)
    H = 128
    W0 = rng.standard_normal((H, 2)) / np.sqrt(2.0)
    b0 = 0.1 * rng.standard_normal(H)
    W1 = rng.standard_normal((H, H)) / np.sqrt(H)
    b1 = 0.1 * rng.standard_normal(H)
    w_out = rng.standard_normal(H) / np.sqrt(H)
    return dict(W0=W0.astype(np.float32), b0=b0.astype(np.float32), W1=W1.astype(np.float32), b1=b1.astype(np.float32),
                w_out=w_out.astype(np.float32), b_out=np.float32(0.05))


FLOP_PER_POINT = 4 * (3 * 128 + 128 * 128)      # value + Jacobian, 2->128->128->1 = 67,072
TOTAL_PROBLEMS = 65536                          # BASELINE.json configs[3]


def ncu_dram_bytes_per_point():
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel per SDF point, read from the committed summary of
    the `ncu --set full` capture (profiles/sdf_tc_kernel_dram.json, written by tools/ncu_summary.py with the capture's point
    count).  None when the file is absent."""
    p = REPO / "profiles" / "sdf_tc_kernel_dram.json"
    if not p.exists():
        return None, None
    d = json.loads(p.read_text())
    return (d["dram_bytes_read"] + d["dram_bytes_write"]) / d["points"], d
BYTES_PER_EVAL = (727 + 1057 + 3225) * 4        # read w, write g and nnz(J) = 20,036 B / problem-eval
DYN_BYTES_PER_PROBLEM = (567 + 160 + 560 + 2080) * 4   # K2 on benchmark_6 = 13,468 B / problem
POINTS_PER_PROBLEM = 324


# ---------------------------------------------------------------------------------------------------------
# CPU baseline: strong batched CPU port (baseline B + C) on a bounded sample
# ---------------------------------------------------------------------------------------------------------
def cpu_step_factory(n_problems: int):
    """One step of the same workload on the host cores: numpy assembly (oracle.nlp_oracle, fp64) with the SDF
    evaluated by batched torch-CPU fp32 forward + hand-written reverse pass over all host threads."""
    import torch
    import yaml
    from oracle import nlp_oracle as no
    torch.set_num_threads(os.cpu_count() or 1)           # every host core at every N (torchrun's default is 1 thread per rank)
    net = synthetic_net()
    t = {k: torch.from_numpy(np.asarray(v)) for k, v in net.items()}
    W0t, W1t = t["W0"].T.contiguous(), t["W1"].T.contiguous()

    def sdf(P):
        with torch.no_grad():
            p = torch.from_numpy(np.ascontiguousarray(P, np.float32))
            a0 = torch.addmm(t["b0"], p, W0t)
            h0 = torch.relu(a0)
            a1 = torch.addmm(t["b1"], h0, W1t)
            s = torch.relu(a1) @ t["w_out"] + t["b_out"]
            g1 = (a1 > 0).float() * t["w_out"]
            g0 = (g1 @ t["W1"]) * (a0 > 0).float()
            J = g0 @ t["W0"]
        return s.numpy(), J.numpy()

    spec = no.NlpSpec.from_yaml_dict(yaml.safe_load(open(REPO / "nlotrajectories_b200" / "benchmarks" / YAML)))
    w = no.multistart_guess(spec, n_problems)

    def step():
        no.eval_g_jac(spec, w, sdf)
        no.eval_f_grad(spec, w)
    return step, torch.get_num_threads()


def time_cpu(n_problems: int, steps: int, warmup: int):
    step, threads = cpu_step_factory(n_problems)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return n_problems * POINTS_PER_PROBLEM / dt, dt, threads


def reference_call_pattern_rate(max_calls: int = 300):
    """Baseline A: the reference's real call pattern - the shipped TorchScript
    modules nn_sdf.pt + jac_nn_sdf.pt called one 1x2 point at a time on one thread."""
    ref = REPO / "oracle" / "_ref"
    if not (ref / "nn_sdf.pt").exists():
        return None
    import torch
    torch.set_num_threads(1)
    f = torch.jit.load(str(ref / "nn_sdf.pt")); j = torch.jit.load(str(ref / "jac_nn_sdf.pt"))
    p = torch.rand(1, 2)
    for _ in range(30):
        f(p); j(p)
    t0 = time.perf_counter()
    for _ in range(max_calls):
        f(p); j(p)
    return max_calls / (time.perf_counter() - t0)


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    n_problems = args.cpu_problems
    rate, dt, threads = time_cpu(n_problems, args.steps, args.warmup)
    cores = os.cpu_count()
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{YAML} NLP eval (g, nnz(dg/dw), f, grad f); CPU sample of {n_problems} multi-start problems per step "
                               f"({n_problems * POINTS_PER_PROBLEM} SDF points), mlp ReLU 2-128-128-1 synthetic weights"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "host_cpus": cores, "kind": "port",
                         "sample": f"{n_problems} problems/step x {args.steps} steps; numpy fp64 assembly (oracle.nlp_oracle) + torch-CPU fp32 "
                                   "batched SDF forward + analytic reverse on all host threads"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# CUDA arm
# ---------------------------------------------------------------------------------------------------------
def plan_solver_starts(args):
    """The solve block plans one seeded RRT guess per start ON THE DEVICE (nlo_rrt_paths: tree search, corner midpoints and shortcut in
    one kernel, a warp per planner); only the cubic spline and the lifting of each path run on host processes, which are forked here,
    BEFORE this process creates its CUDA context or NCCL threads (a process that holds either must not fork)."""
    if args.solve_problems <= 0:
        return None
    from nlotrajectories_b200.config import Config
    from nlotrajectories_b200.rrt_device import make_post_pool
    cfg = Config.load(REPO / "nlotrajectories_b200" / "benchmarks" / YAML)
    world = int(os.environ.get("WORLD_SIZE", 1))
    workers = max(1, min(32, (os.cpu_count() or 1) // world))
    return {"cfg": cfg, "pool": make_post_pool(cfg, lift=True, workers=workers), "workers": workers}


def run_solver(args, planned, rank, local_rank, world, dev):
    """The caller on both sides of the hot path at BASELINE size: `--solve-problems` benchmark_6 starts (sharded by problem index over
    the GPUs) solved by the device interior point (nlo_ip_*: block-tridiagonal KKT kernels), with the SDF network trained on the
    YAML's scene (tests/golden/sdf_benchmark_6_relu128.npz: train.py, 10^6 samples); best-of-batch selection over the SOLVED starts
    is the one collective of the path."""
    import torch
    import torch.distributed as dist
    from nlotrajectories_b200.config import Config
    from nlotrajectories_b200.distributed import select_best, shard_range
    from nlotrajectories_b200.problem import NlpProblem
    from nlotrajectories_b200.sdf import LearnedSDF, SdfWeights
    from nlotrajectories_b200.solver import DeviceIPSolver
    wpath = REPO / "tests" / "golden" / "sdf_benchmark_6_relu128.npz"
    model = LearnedSDF(SdfWeights.from_npz(wpath), device=local_rank, precision=args.precision)
    prob = NlpProblem.from_config(Config.load(REPO / "nlotrajectories_b200" / "benchmarks" / YAML), model, device=local_rank)
    total = args.solve_problems
    lo, hi = shard_range(total, rank, world)
    P = hi - lo
    from nlotrajectories_b200.rrt_device import rrt_multistart_device
    t0 = time.perf_counter()
    w0 = rrt_multistart_device(planned["cfg"], P, first=lo, lift=True, device=dev, pool=planned["pool"], workers=planned["workers"]).astype(np.float64)
    plan_s = time.perf_counter() - t0
    planned["pool"].close(); planned["pool"].join()
    solver = DeviceIPSolver(prob, max_problems=max(P, 1), max_iter=args.solve_max_iter)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = solver.solve(w0)
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    ok, st, v, f = res.converged.numpy(), res.stalled.numpy(), res.violation.numpy(), res.f.numpy()
    usable = ok | (st & (v <= 1e-4))
    counts = torch.tensor([ok.sum(), (st & (v <= 1e-4)).sum(), P], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    score = torch.from_numpy(np.where(usable, f, f + 1e3 * (1.0 + v))).float().to(dev)
    best_val, best_idx, _ = select_best(score, res.w.float().T.contiguous().to(dev), lo, prob.n_w)
    n_ok, n_st, n_all = (float(c) for c in counts.tolist())
    return {"workload": f"{YAML} x {total} starts sharded over {world} GPU(s): one seeded RRT plan per start (device tree search + shortcut, nlo_rrt_paths; "
                        "spline and lifting on host processes)",
            "sdf_model": "tests/golden/sdf_benchmark_6_relu128.npz (ReLU 2-128-128-1 trained on the YAML's scene)",
            "solver": "device interior point, tol 1e-4, exact Hessian, block-tridiagonal KKT kernels (nlo_ip_solve)", "max_iter": args.solve_max_iter,
            "solve_s": float(t.item()), "solves_per_s": n_all / float(t.item()), "converged_frac": n_ok / max(n_all, 1.0),
            "stalled_feasible_frac": n_st / max(n_all, 1.0), "stats_rank0": solver.stats, "rrt_plan_s_rank0": plan_s,
            "best_of_batch": {"objective": best_val, "global_index": best_idx, "selection": "all-gather of (objective, index) + broadcast of the winner over the SOLVED starts"}}


def run_ours(args):
    planned = plan_solver_starts(args)
    import torch
    import torch.distributed as dist
    from nlotrajectories_b200 import lib
    from nlotrajectories_b200.config import Config
    from nlotrajectories_b200.distributed import bind_to_gpu_numa, init_process_group, merit, select_best, shard_range
    from nlotrajectories_b200.problem import NlpProblem
    from nlotrajectories_b200.sdf import LearnedSDF, SdfWeights

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    rank, local_rank, world = init_process_group("nccl")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_cpus = bind_to_gpu_numa(local_rank) if world > 1 else 0      # before any pinned allocation
    L = lib.load()
    peaks, peak_src = load_peaks()
    net = synthetic_net()
    weights = SdfWeights.pack("mlp", net["W0"], net["b0"], [(net["W1"], net["b1"])], net["w_out"], float(net["b_out"]), 0, 0)
    model = LearnedSDF(weights, device=local_rank, precision=args.precision)
    prob = NlpProblem.from_config(Config.load(REPO / "nlotrajectories_b200" / "benchmarks" / YAML), model, device=local_rank)
    steps, warmup = args.steps, max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def measure(P, first, n_steps, with_e2e):
        """Device-resident and host-buffer timing of this rank's shard of P problems (global indices first..first+P-1)."""
        w_host = prob.multistart_guess(P, first=first)                   # (P, n_w) problem-major, seeded by global index
        w_pin = torch.empty((P, prob.n_w), dtype=torch.float32).pin_memory()
        w_pin.copy_(torch.from_numpy(w_host))
        w = w_pin.to(dev).T.contiguous()                                  # SoA (n_w, P), resident in HBM
        g, jac, f, grad = prob.alloc_outputs(P, dev)
        prob.reserve(P)
        for _ in range(warmup):
            prob.eval_device(w, g, jac, f, grad)
        barrier()
        l0 = L.nlo_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(n_steps):
            prob.eval_device(w, g, jac, f, grad)
        e1.record()
        barrier()
        out = {"P": P, "w": w, "g": g, "f": f, "launches": int(L.nlo_launch_count() - l0), "w_pin": w_pin,
               "ms_per_step": max_over_ranks(e0.elapsed_time(e1)) / n_steps,
               "resident_bytes": (w.numel() + g.numel() + jac.numel() + grad.numel()) * 4}
        if with_e2e:
            out.update(measure_e2e(P, w_pin))
        return out

    def measure_e2e(P, w_pin):
        """Host buffers in, host buffers out, through the C-ABI host entry points; synchronous calls, wall clock, max over ranks."""
        res = {}
        w_np = w_pin.numpy()
        e2e_steps = max(2, min(steps, 5))
        pin = lambda shape: torch.empty(shape, dtype=torch.float32).pin_memory().numpy()
        lay = prob.compact_layout()
        full = {"g": pin((P, prob.n_g)), "jac": pin((P, prob.nnz)), "f": pin((P,)), "grad_f": pin((P, prob.n_w))}
        comp = {"g": pin((P, len(lay["g_var_rows"]))), "jac": pin((P, len(lay["jac_var_nz"]))), "f": pin((P,)),
                "grad_f": pin((P, len(lay["grad_var_idx"])))}
        for name, call, bufs in (("full", lambda: prob.eval_host(w_np, out=full), full),
                                 ("compact", lambda: prob.eval_host_compact(w_np, out=comp), comp)):
            call()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                call()                                                    # synchronous: results are in host memory on return
            barrier()
            res[name] = {"s": max_over_ranks((time.perf_counter() - t0) / e2e_steps), "h2d": int(w_np.nbytes),
                         "d2h": int(sum(v.nbytes for v in bufs.values()))}
        return {"e2e": res}

    sampler = ClockSampler(local_rank)          # samples SM clocks / throttle reasons across every timed region below
    sampler.start()
    # ---- headline: the 65,536-start batch sharded over the ranks (strong), or 65,536 per rank (weak) ---------------------
    total = args.problems
    lo, hi = shard_range(total, rank, world)
    weak = None
    if args.scaling == "strong":
        head = measure(hi - lo, lo, steps, with_e2e=True)
        if world > 1:                            # the weak-scaling variant beside it (fewer steps: it is a side note)
            weak = measure(total, rank * total, max(5, steps // 4), with_e2e=False)
    else:
        lo = rank * total
        head = measure(total, lo, steps, with_e2e=True)
    P = head["P"]
    w, g, f = head["w"], head["g"], head["f"]
    n_pts = P * prob.n_sdf_points
    pts_total = int(sum_over_ranks(n_pts))
    ms_per_step = head["ms_per_step"]
    value = pts_total / (ms_per_step * 1e-3)

    # ---- the dominant kernel alone (the fused learned-SDF value+Jacobian kernel on this rank's footprint points) ---------
    xs = torch.rand(n_pts, device=dev) * 2 - 0.5
    ys = torch.rand(n_pts, device=dev) * 2 - 0.5
    so_, jx_, jy_ = torch.empty_like(xs), torch.empty_like(xs), torch.empty_like(xs)
    for _ in range(3):
        model.eval(xs, ys, out=(so_, jx_, jy_))
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(steps):
        model.eval(xs, ys, out=(so_, jx_, jy_))
    k1.record()
    torch.cuda.synchronize()
    k_ms = k0.elapsed_time(k1) / steps
    # K2 alone (Euler defects + banded Jacobian values): the HBM-bound kernel of the path
    jac_d = torch.empty((prob.nnz, P), dtype=torch.float32, device=dev)
    for _ in range(3):
        prob.eval_dynamics_device(w, g, jac_d)
    torch.cuda.synchronize()
    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d0.record()
    for _ in range(steps):
        prob.eval_dynamics_device(w, g, jac_d)
    d1.record()
    torch.cuda.synchronize()
    dyn_ms = d0.elapsed_time(d1) / steps
    del jac_d
    prob.eval_device(w, g, None, f, None)       # g again in full (the K2 timing above rewrote only its defect rows)
    clocks = sampler.stop()

    # ---- best-of-batch selection (the only collective on the path; outside the timed region) ------------------
    lb, ub = prob.bounds()
    big = 3.0e38
    lbd = torch.from_numpy(np.clip(lb, -big, big).astype(np.float32)).to(dev)
    ubd = torch.from_numpy(np.clip(ub, -big, big).astype(np.float32)).to(dev)
    viol = prob.violation(g, lbd, ubd)
    best_val, best_idx, _ = select_best(merit(f, viol), w, lo, prob.n_w)
    solved = None
    if planned is not None:
        for blk in (head, weak):                          # free the timed batch before the solver allocates its working set
            if blk is not None:
                for k in ("w", "g", "f", "w_pin"):
                    blk.pop(k, None)
        del w, g, f, viol, xs, ys, so_, jx_, jy_
        torch.cuda.empty_cache()
        solved = run_solver(args, planned, rank, local_rank, world, dev)

    if rank != 0:
        return
    prec = model.precision
    if prec == "tc3xf16":
        peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
        bound = "tensor"
        peak_note = (f"{peak_src} cuBLAS bf16 sustained; per point the kernel executes 5 fp16 MMA passes for 2 algorithmic contractions "
                     "(forward: hi.hi, hi.lo, lo.hi; reverse of the ReLU layer: exact 0/1 mask x hi/lo of diag(w2)W1) plus one K=16 MMA that evaluates "
                     "layer 0 (sdf_tc_rr_kernel), so its algorithmic ceiling is peak x 2/5.125")
    else:
        sm_mhz = clocks.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
        peak = model_sm_count(L, local_rank) * 128 * 2 * sm_mhz * 1e6 / 1e12
        bound = "fp32_simt"
        peak_note = f"FP32 FMA pipe = SMs x 128 lanes x 2 x {sm_mhz:.0f} MHz (median SM clock sampled during the run)"
    achieved = FLOP_PER_POINT * n_pts / (k_ms * 1e-3) / 1e12
    dram_pp, dram_src = ncu_dram_bytes_per_point() if prec == "tc3xf16" else (None, None)
    e2e, e2e_full = head["e2e"]["compact"], head["e2e"]["full"]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32" if prec == "fp32" else "f32 (tcgen05 split-fp16 hi/lo tiles, fp32 accumulate)", "data": "synthetic",
        "config": {"workload": (f"{YAML} x {total} multi-starts " + ("sharded by problem index over the GPUs" if args.scaling == "strong" else "per GPU")
                                + ": NLP eval g + nnz(dg/dw) + f + grad f, SoA fp32 resident in HBM"),
                   "problems_total": pts_total // prob.n_sdf_points, "problems_per_gpu": P,
                   "sdf_points_per_step_per_gpu": n_pts, "n_w": prob.n_w, "n_g": prob.n_g, "nnz_jac": prob.nnz,
                   "sdf_model": "mlp ReLU 2-128-128-1 (synthetic: numpy default_rng(0), W ~ N(0, 1/fan_in), b ~ 0.1 N(0, 1))", "sdf_precision": prec,
                   "l2": f"inputs + outputs resident per GPU {head['resident_bytes'] / 1e6:.0f} MB per step (every byte read or written once per step) vs the 126 MB L2",
                   "parallelism": f"problem-sharded x{world}, no data-path collective"},
        "nlp_evals_per_s": pts_total / prob.n_sdf_points / (ms_per_step * 1e-3),
        "clocks": clocks,
        "e2e": {"value": pts_total / e2e["s"], "unit": UNIT, "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                "ms_per_step": e2e["s"] * 1e3,
                "api": "nlo_nlp_eval_host_compact (pinned host buffers, problem-major; returns every entry of g, nnz(dg/dw), grad f that varies with w - "
                       "the rest are copies of w and constants published once by nlo_nlp_compact_layout)",
                "full_form": {"api": "nlo_nlp_eval_host (every entry, constants included)", "value": pts_total / e2e_full["s"], "ms_per_step": e2e_full["s"] * 1e3,
                              "h2d_bytes_per_step": e2e_full["h2d"], "d2h_bytes_per_step": e2e_full["d2h"]},
                "cpu_affinity": (f"rank pinned to the {numa_cpus} CPUs NVML reports local to its GPU" if numa_cpus else "unchanged")},
        "gpu_launches": head["launches"],
        "roofline": {"kernel": "sdf_tc_rr_kernel (tcgen05; ReLU / ReLU form of sdf_tc_kernel)" if prec == "tc3xf16" else "sdf_simt_kernel", "bound": bound, "achieved": achieved,
                     "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "traffic": (dram_pp * n_pts if dram_pp is not None else None), "traffic_unit": "bytes per launch (ncu dram read+write)",
                     "traffic_source": (f"profiles/sdf_tc_kernel_dram.json: {dram_src['dram_bytes_read']:.0f} B read + {dram_src['dram_bytes_write']:.0f} B written "
                                        f"for {dram_src['points']} points ({dram_src.get('source', '')}) = {dram_pp:.2f} B/point, scaled to this launch"
                                        if dram_src else None),
                     "algorithmic_bytes_per_launch": 20 * n_pts,
                     "flop_per_point": FLOP_PER_POINT, "points_per_launch": n_pts, "kernel_ms": k_ms, "peak_source": peak_note,
                     "kernel_note": "timed alone in its plain form (points read from memory, s / jx / jy written); inside the step the same tile "
                                    "code runs in its rows form (K3 fused: footprint points formed from the poses, results written straight "
                                    "into g and dg/dw), two launches per evaluation",
                     "kernel_share_of_step": k_ms / ms_per_step,
                     "dynamics_kernel_hbm": {"kernel": "nlp_phase0_kernel (defect rows only)", "algorithmic_bytes_per_launch": DYN_BYTES_PER_PROBLEM * P,
                                             "kernel_ms": dyn_ms, "achieved_gbs": DYN_BYTES_PER_PROBLEM * P / (dyn_ms * 1e-3) / 1e9,
                                             "peak_gbs": peaks["hbm_gbs"], "frac": DYN_BYTES_PER_PROBLEM * P / (dyn_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                             "note": "SURVEY.md 8(d): read (N+1)nx + N nu, write N nx residuals + 26 N Jacobian values per problem"},
                     "step_hbm": {"algorithmic_bytes_per_step": BYTES_PER_EVAL * P, "achieved_gbs": BYTES_PER_EVAL * P / (ms_per_step * 1e-3) / 1e9,
                                  "peak_gbs": peaks["hbm_gbs"]}},
        "best_of_batch": {"merit": best_val, "global_index": best_idx, "note": "selection over the EVALUATED initial guesses of the timed batch "
                          "(synthetic SDF weights); the selection over SOLVED starts is in the `solve` block"},
    }
    if solved is not None:
        line["solve"] = solved
    if weak is not None:
        line["weak"] = {"problems_per_gpu": weak["P"], "ms_per_step": weak["ms_per_step"], "unit": UNIT,
                        "value": world * weak["P"] * prob.n_sdf_points / (weak["ms_per_step"] * 1e-3)}
    if world == 1 and not args.no_cpu_baseline:
        cpu_steps = 7
        rate, dt, threads = time_cpu(args.cpu_problems, cpu_steps, 1)
        cb = {"value": rate, "unit": UNIT, "cores": threads, "host_cpus": os.cpu_count(), "kind": "port",
              "sample": f"{args.cpu_problems} of the {P} problems per step x {cpu_steps} steps ({dt * cpu_steps:.1f} s of CPU work); numpy fp64 assembly "
                        "(oracle.nlp_oracle) + torch-CPU fp32 batched SDF forward + analytic reverse on all host threads"}
        a = reference_call_pattern_rate()
        if a is not None:
            cb["reference_call_pattern_points_per_s_1thread"] = a
        line["cpu_baseline"] = cb
    print(json.dumps(line), flush=True)


def model_sm_count(L, device):
    return int(L.nlo_device_sm_count(device))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--problems", type=int, default=TOTAL_PROBLEMS, help="multi-start problems: in total (strong) / per GPU (weak)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: --problems in total, sharded over the GPUs (BASELINE configs[3]); weak: --problems per GPU")
    ap.add_argument("--precision", default="auto", choices=["auto", "fp32", "tc3xf16"])
    ap.add_argument("--cpu-problems", type=int, default=8192, help="problems per CPU-baseline step (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--solve-problems", type=int, default=16384, help="starts solved by the device interior point after the timed evaluation (0: skip)")
    ap.add_argument("--solve-max-iter", type=int, default=300)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()
