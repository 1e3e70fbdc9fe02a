"""SURVEY.md 8(d) measurement sweep (run on the GPU box): synthetic learned-SDF sweep over widths and point counts,
benchmark-shaped NLP evaluation for B1/B3/B4/B6, and the single-point CasADi-ABI latency.  Prints one JSON line per row."""
import ctypes as C
import json
import sys
import time
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
import torch
from oracle import sdf_oracle as so
from gpu_util import to_weights
from nlotrajectories_b200 import lib
from nlotrajectories_b200.config import Config
from nlotrajectories_b200.problem import NlpProblem
from nlotrajectories_b200.sdf import LearnedSDF

PEAKS = json.loads((REPO / "MEASURED_PEAKS.json").read_text()) if (REPO / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}


def timed(fn, reps):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def sdf_sweep(max_log=28):
    nets = {"relu64": so.synthetic_mlp(64, 1, seed=1), "relu128": so.synthetic_mlp(128, 1, seed=0), "relu256": so.synthetic_mlp(256, 1, seed=2),
            "shipped_fourier128": so.from_npz(REPO / "tests/golden/sdf_shipped_fourier128_weights.npz")}
    zoo = {"tanh128": so.synthetic_mlp(128, 1, seed=13, act=so.ACT_TANH), "sigmoid128": so.synthetic_mlp(128, 1, seed=15, act=so.ACT_SIGMOID),
           "leaky128": so.synthetic_mlp(128, 1, seed=16, act=so.ACT_LEAKY_RELU), "siren64x1": so.synthetic_siren(64, 1, omega0=30.0, seed=17),
           "fourier128_tanh": so.synthetic_fourier(128, 1, scale=2.0, seed=21, act=so.ACT_TANH)}
    runs = ([(name, net, "auto") for name, net in nets.items()] + [("relu128", nets["relu128"], "fp32"), ("relu256", nets["relu256"], "fp32")]
            + [(name, net, "zoo") for name, net in zoo.items()])
    for name, net, prec in runs:
        model = LearnedSDF(to_weights(net), precision="auto" if prec == "zoo" else prec)
        for logn in (20, 22, 24, 26, 28):
            if logn > max_log or (model.precision == "fp32" and logn > 22) or (prec == "zoo" and logn != 24): continue
            n = 1 << logn
            x = torch.rand(n, device="cuda") * 2 - 0.5; y = torch.rand(n, device="cuda") * 2 - 0.5; sb = torch.rand(n, device="cuda") + 0.5
            out = (torch.empty_like(x), torch.empty_like(x), torch.empty_like(x))
            reps = max(2, min(20, (1 << 26) // n))
            ms_vj = timed(lambda: model.eval(x, y, out=out), reps)
            ms_adj = timed(lambda: model.eval(x, y, sb, out=out), reps)
            ms_v = timed(lambda: model.eval(x, y, want_jac=False, out=(out[0], None, None)), reps)
            fl = net.flops_value_jac()
            print(json.dumps({"row": "sdf_sweep", "net": name, "precision": model.precision, "points": n,
                              "value_jac_Gpts_s": n / ms_vj / 1e6, "adjoint_Gpts_s": n / ms_adj / 1e6, "value_only_Gpts_s": n / ms_v / 1e6,
                              "algorithmic_TFLOPs": fl * n / ms_vj / 1e9, "frac_of_bf16_peak": fl * n / ms_vj / 1e9 / PEAKS["bf16_tflops_sustained"],
                              "hbm_GBs_at_20B_per_point": 20 * n / ms_vj / 1e6}), flush=True)
            del x, y, sb, out
        model.close()


def nlp_rows():
    net = so.synthetic_mlp(128, 1, seed=0)
    for name, P in (("benchmark_1", 1), ("benchmark_3", 1), ("benchmark_4", 4096), ("benchmark_6", 65536), ("benchmark_1", 65536)):
        cfg = Config.load(next((REPO / "nlotrajectories_b200/benchmarks").glob(name + "*.yaml")))
        model = LearnedSDF(to_weights(net)) if cfg.solver.mode == "l4casadi" else None
        prob = NlpProblem.from_config(cfg, model)
        w = torch.from_numpy(prob.multistart_guess(P)).cuda().T.contiguous()
        g, jac, f, grad = prob.alloc_outputs(P)
        ms = timed(lambda: prob.eval_device(w, g, jac, f, grad), 20 if P > 1 else 200)
        byt = (prob.n_w + prob.n_g + prob.nnz) * 4
        row = {"row": "nlp_eval", "config": name, "problems": P, "ms_per_eval": ms, "problem_evals_per_s": P / ms * 1e3,
               "algorithmic_GBs": byt * P / ms / 1e6, "frac_of_hbm_peak": byt * P / ms / 1e6 / PEAKS["hbm_gbs"],
               "sdf_points_per_s": (prob.n_sdf_points * P / ms * 1e3) if model else 0}
        if P == 1:
            wh = prob.multistart_guess(1)
            t0 = time.perf_counter()
            for _ in range(200): prob.eval_host(wh)
            row["host_call_latency_us"] = (time.perf_counter() - t0) / 200 * 1e6
        print(json.dumps(row), flush=True)


def casadi_latency():
    net = so.from_npz(REPO / "tests/golden/sdf_shipped_fourier128_weights.npz")
    model = LearnedSDF(to_weights(net)); model.bind_casadi(1)
    L = lib.load(); DP = C.POINTER(C.c_double)
    p = np.array([0.3, 0.7]); out = np.zeros(4)
    dp = lambda a: a.ctypes.data_as(DP)
    res = {}
    for fn, args in (("nn_sdf", (DP * 1)(dp(p))), ("jac_nn_sdf", (DP * 2)(dp(p), None)), ("adj1_nn_sdf", (DP * 3)(dp(p), None, dp(np.ones(1)))),
                     ("jac_adj1_nn_sdf", (DP * 4)(dp(p), None, dp(np.ones(1)), None))):
        r = (DP * 3)(dp(out), None, None)
        f = getattr(L, fn)
        for _ in range(50): f(args, r, None, None, 0)
        t0 = time.perf_counter()
        for _ in range(500): f(args, r, None, None, 0)
        res[fn + "_us_per_call"] = (time.perf_counter() - t0) / 500 * 1e6
    print(json.dumps({"row": "casadi_abi_single_point_latency", **res}), flush=True)


if __name__ == "__main__":
    sdf_sweep(int(sys.argv[1]) if len(sys.argv) > 1 else 28)
    nlp_rows()
    casadi_latency()
