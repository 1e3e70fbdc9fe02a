"""Evaluation time of the analytic-obstacle configurations (solver.mode casadi: benchmarks 1, 2, 5) at 65,536 starts: one launch
(defects, copy rows, footprint + union SDF + rows, gradient, objective), against the algorithmic traffic w + g + nnz(dg/dw)."""
import json
import sys
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO))
from nlotrajectories_b200 import lib                       # noqa: E402
from nlotrajectories_b200.config import Config             # noqa: E402
from nlotrajectories_b200.problem import NlpProblem        # noqa: E402

peaks = json.loads((REPO / "MEASURED_PEAKS.json").read_text()) if (REPO / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6650.0}
P = 65536
L = lib.load()
for name in ("benchmark_1", "benchmark_2", "benchmark_5"):
    prob = NlpProblem.from_config(Config.load(next((REPO / "nlotrajectories_b200/benchmarks").glob(name + "*.yaml"))), None)
    w = torch.from_numpy(prob.multistart_guess(P)).cuda().T.contiguous()
    g, jac, f, grad = prob.alloc_outputs(P)
    for _ in range(3):
        prob.eval_device(w, g, jac, f, grad)
    torch.cuda.synchronize()
    l0 = L.nlo_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        prob.eval_device(w, g, jac, f, grad)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    algo = (prob.n_w + prob.n_g + prob.nnz) * 4 * P
    full = algo + (prob.n_w + 1) * 4 * P                   # + grad f and f
    print(json.dumps({"config": name, "problems": P, "ms_per_eval": ms, "launches_per_eval": (L.nlo_launch_count() - l0) / 20,
                      "algorithmic_gbs_w_g_jac": algo / ms / 1e6, "frac_of_hbm_peak": algo / ms / 1e6 / peaks["hbm_gbs"],
                      "gbs_incl_grad_f": full / ms / 1e6, "evals_per_s": P / ms * 1e3}), flush=True)
