"""Generate the golden fixtures under tests/golden/ from the REFERENCE ITSELF (run in the build container).

TEST INFRASTRUCTURE (see oracle/__init__.py).  ``/root/reference`` does not exist on the GPU box,
so this script is run once here and its small outputs are committed:

1. ``sdf_shipped_fourier128.npz`` - weights extracted from ``_l4c_generated/nn_sdf.pt`` and the
   outputs of the four TorchScript modules ``{nn_sdf,jac_nn_sdf,adj1_nn_sdf,jac_adj1_nn_sdf}.pt``
   (the functions ``_l4c_generated/nn_sdf.cpp:57-104`` calls) at seeded points.
2. ``nlp_<benchmark>.npz`` - g(w), dg/dw (structural non-zeros, CCS order), f(w), grad f(w), lbg, ubg
   obtained by executing the reference's UNMODIFIED ``core/runner.py`` (+dynamics, geometry, utils,
   sdf) under ``oracle/casadi_stub`` (sympy) and lambdifying what ``Opti`` recorded.  The learned
   SDF node is bound to the fp64 numpy SDF oracle of fixture 1.
3. ``oracle/_ref/*.pt`` - verbatim copies of the four TorchScript artefacts (git-ignored; used only
   by ``bench.py``'s CPU-baseline legs as ``kind: "reference"``).

Usage:  python oracle/make_golden.py [--reference /root/reference]
"""
from __future__ import annotations

import argparse
import shutil
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))

from oracle import nlp_oracle as no  # noqa: E402
from oracle import sdf_oracle as so  # noqa: E402

GOLD = REPO / "tests" / "golden"
SEED_POINTS = 1


def shipped_net(ref: Path) -> so.SdfNet:
    import torch
    m = torch.jit.load(str(ref / "_l4c_generated" / "nn_sdf.pt"))
    c = {k: v.detach().numpy() for k, v in m.code_with_constants[1].const_mapping.items()}
    # graph: h0 = cos(p @ c0 + c1) * 10 ; h1 = relu(h0 @ c3 + c2) ; s = h1 @ c5 + c4
    return so.SdfNet("fourier", c["c0"].T.copy(), c["c1"].copy(), [(c["c3"].T.copy(), c["c2"].copy())],
                     c["c5"][:, 0].copy(), float(c["c4"][0]), so.ACT_COS_SCALE, so.ACT_RELU, 10.0, 1.0)


def make_sdf_golden(ref: Path) -> so.SdfNet:
    import torch
    torch.set_num_threads(1)
    net = shipped_net(ref)
    gen = ref / "_l4c_generated"
    f_val = torch.jit.load(str(gen / "nn_sdf.pt"))
    f_jac = torch.jit.load(str(gen / "jac_nn_sdf.pt"))
    f_adj = torch.jit.load(str(gen / "adj1_nn_sdf.pt"))
    f_hes = torch.jit.load(str(gen / "jac_adj1_nn_sdf.pt"))
    rng = np.random.default_rng(SEED_POINTS)
    survey_pts = np.array([[0, 0], [0.5, 0.5], [0.3, 0.7], [1.0, 1.0], [0.8, 0.2], [-0.25, 1.25]], np.float32)
    P = np.concatenate([survey_pts, rng.uniform(-0.5, 1.5, (506, 2)).astype(np.float32)])
    sbar = np.concatenate([np.full(6, 2.0, np.float32), rng.uniform(0.5, 1.5, 506).astype(np.float32)])
    n = P.shape[0]
    val = f_val(torch.from_numpy(P)).numpy()[:, 0]
    adj = f_adj(torch.from_numpy(P), torch.from_numpy(sbar[:, None])).numpy()
    jac = np.zeros((n, 2), np.float32)
    hes = np.zeros((n, 2, 2), np.float32)
    for i in range(n):  # jac / jac_adj1 graphs are shape-specialised to 1x2
        p = torch.from_numpy(P[i:i + 1])
        jac[i] = f_jac(p).numpy().reshape(2)
        hes[i] = f_hes(p, torch.from_numpy(sbar[i:i + 1, None])).numpy().reshape(2, 2)
    so.to_npz(net, GOLD / "sdf_shipped_fourier128_weights.npz")
    np.savez(GOLD / "sdf_shipped_fourier128.npz", P=P, sbar=sbar, value=val, jac=jac, adj1=adj, jac_adj1=hes)
    # copies of the artefacts for the bench's "reference" CPU leg (git-ignored)
    out = REPO / "oracle" / "_ref"
    out.mkdir(exist_ok=True)
    for name in ("nn_sdf", "jac_nn_sdf", "adj1_nn_sdf", "jac_adj1_nn_sdf"):
        shutil.copyfile(gen / f"{name}.pt", out / f"{name}.pt")
    print(f"sdf golden: {n} points; value range [{val.min():.3f}, {val.max():.3f}]")
    return net


def record_reference_nlp(ref: Path, yaml_path: Path):
    """Execute the reference's RunBenchmark.run() under the stub; return the recording Opti."""
    stub = str(REPO / "oracle" / "casadi_stub")
    src = str(ref / "src")
    for p in (src, stub):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    import casadi as ca  # the stub
    import l4casadi as l4c  # the stub
    import yaml
    from nlotrajectories.core.config import Config
    from nlotrajectories.core.runner import RunBenchmark
    from nlotrajectories.core.sdf.l4casadi import NNObstacle
    from nlotrajectories.core.trajectory_initialization import LinearInitializer

    assert "casadi_stub" in ca.__file__
    raw = yaml.safe_load(open(yaml_path))
    config = Config(**raw)
    obstacles = config.get_obstacles() if config.solver.mode == "casadi" else None
    if config.solver.mode == "l4casadi":
        # scripts/run_benchmark.py:97-101 without the training step: the model is the external node
        obstacles = NNObstacle(None, l4c.L4CasADi(None, device="cpu"))
    geometry = config.body.create_geometry()
    runner = RunBenchmark(
        dynamics=config.body.create_dynamics(), geometry=geometry,
        x0=ca.MX(config.body.start_state), x_goal=ca.MX(config.body.goal_state),
        N=config.solver.N, dt=config.solver.dt, sdf_func=obstacles.approximated_sdf,
        control_bounds=tuple(config.body.control_bounds), use_slack=config.solver.use_slack,
        slack_penalty=config.solver.slack_penalty, use_smooth=config.solver.use_smooth,
        smooth_weight=config.solver.smooth_weight,
        initializer=LinearInitializer(N=config.solver.N, x0=np.array(config.body.start_state),
                                      x_goal=np.array(config.body.goal_state)),
        enforce_heading=config.solver.enforce_heading)
    X_opt, U_opt, opti, X_init, status = runner.run()
    assert status == "success"
    return raw, opti


def lambdify_reference(opti, net64: so.SdfNet):
    import sympy as sp
    wsym = opti.w_symbols()
    index = {s: i for i, s in enumerate(wsym)}

    def _v(fn):
        def call(x, y):
            P = np.stack([np.atleast_1d(np.asarray(x, float)), np.atleast_1d(np.asarray(y, float))], axis=-1)
            out = fn(P)
            return out[0] if np.ndim(x) == 0 else out
        return call
    mods = [{"sdf": _v(lambda P: so.forward(net64, P)),
             "sdf_dx": _v(lambda P: so.value_jac(net64, P)[1][:, 0]),
             "sdf_dy": _v(lambda P: so.value_jac(net64, P)[1][:, 1])}, "numpy"]
    g_exprs = list(opti.g_rows)
    rows, cols, entries = [], [], []
    for r, e in enumerate(g_exprs):
        for s in sorted(e.free_symbols, key=lambda s: index[s]):
            d = sp.diff(e, s)
            if d != 0:
                rows.append(r); cols.append(index[s]); entries.append(d)
    rows, cols = np.array(rows), np.array(cols)
    perm = np.lexsort((rows, cols))
    rows, cols = rows[perm], cols[perm]
    entries = [entries[i] for i in perm]
    grad_exprs = [sp.diff(opti.f, s) for s in wsym]
    f_g = sp.lambdify(wsym, g_exprs, modules=mods, cse=True)
    f_j = sp.lambdify(wsym, entries, modules=mods, cse=True)
    f_f = sp.lambdify(wsym, [opti.f] + grad_exprs, modules=mods, cse=True)
    return rows, cols, f_g, f_j, f_f


def hessian_reference(opti, net64: so.SdfNet, lam: np.ndarray, sigma: float):
    """Upper triangle of hess(sigma * f + lam^T g) from the reference's recorded expressions, by symbolic
    differentiation (the learned SDF's second derivatives are the opaque sdf_dxx/dxy/dyy of the stub, bound to the
    numpy restatement of jac_adj1_nn_sdf)."""
    import sympy as sp
    wsym = opti.w_symbols()
    index = {s: i for i, s in enumerate(wsym)}

    def _v(fn):
        def call(x, y):
            P = np.stack([np.atleast_1d(np.asarray(x, float)), np.atleast_1d(np.asarray(y, float))], axis=-1)
            out = fn(P)
            return out[0] if np.ndim(x) == 0 else out
        return call
    hs = lambda P: so.jac_adj1(net64, P, np.ones(len(P)))
    mods = [{"sdf": _v(lambda P: so.forward(net64, P)),
             "sdf_dx": _v(lambda P: so.value_jac(net64, P)[1][:, 0]), "sdf_dy": _v(lambda P: so.value_jac(net64, P)[1][:, 1]),
             "sdf_dxx": _v(lambda P: hs(P)[:, 0, 0]), "sdf_dxy": _v(lambda P: hs(P)[:, 0, 1]), "sdf_dyy": _v(lambda P: hs(P)[:, 1, 1])}, "numpy"]
    acc = {}
    for coef, e in [(sigma, opti.f)] + [(float(l), e) for l, e in zip(lam, opti.g_rows)]:
        if coef == 0.0:
            continue
        syms = sorted(e.free_symbols, key=lambda s: index[s])
        for ia, a in enumerate(syms):
            da = sp.diff(e, a)
            if da == 0:
                continue
            for b in syms[ia:]:
                if b not in da.free_symbols:
                    continue
                d2 = sp.diff(da, b)
                if d2 != 0:
                    key = (index[a], index[b])
                    acc[key] = acc.get(key, 0) + coef * d2
    keys = sorted(acc, key=lambda rc: (rc[1], rc[0]))
    fn = sp.lambdify(wsym, [acc[k] for k in keys], modules=mods, cse=True)
    return np.array([k[0] for k in keys]), np.array([k[1] for k in keys]), fn


def make_hess_golden(ref: Path, net: so.SdfNet, names, P: int = 2):
    net64 = net.astype(np.float64)
    bench_dir = ref / "src" / "nlotrajectories" / "benchmarks"
    for name in names:
        yaml_path = next(bench_dir.glob(f"{name}*.yaml"))
        raw, opti = record_reference_nlp(ref, yaml_path)
        spec = no.NlpSpec.from_yaml_dict(raw)
        rng = np.random.default_rng(11)
        lam = rng.normal(0, 1, spec.n_g).astype(np.float32).astype(np.float64)
        sigma = 0.75
        rows, cols, fn = hessian_reference(opti, net64, lam, sigma)
        w = golden_w(spec, P, seed=7).astype(np.float32).astype(np.float64)
        vals = np.array([np.asarray(fn(*wi), float) for wi in w])
        np.savez(GOLD / f"nlp_hess_{name}.npz", yaml=np.array(yaml_path.name), w=w, lam=lam, sigma=np.float64(sigma),
                 hess_rows=rows, hess_cols=cols, hess_vals=vals)
        print(f"{name}: hessian of the Lagrangian, {len(rows)} structurally non-zero upper-triangle entries")


def golden_w(spec: no.NlpSpec, P: int, seed: int) -> np.ndarray:
    """Multi-start guesses plus noise on every block so that no Jacobian entry is trivially zero."""
    rng = np.random.default_rng(seed)
    w = no.multistart_guess(spec, P)
    X = w[:, :spec.n_X].reshape(P, spec.N + 1, spec.nx)
    X[:, :, 2:] += rng.normal(0, 0.2, X[:, :, 2:].shape)
    w[:, spec.n_X:spec.n_X + spec.n_U] = rng.normal(0, 0.5, (P, spec.n_U))
    if spec.use_slack:
        w[:, spec.n_X + spec.n_U:] = np.abs(rng.normal(0, 0.05, (P, spec.N + 1)))
    return w


def make_nlp_golden(ref: Path, net: so.SdfNet, names, P: int = 3):
    net64 = net.astype(np.float64)
    bench_dir = ref / "src" / "nlotrajectories" / "benchmarks"
    for name in names:
        yaml_path = next(bench_dir.glob(f"{name}*.yaml"))
        raw, opti = record_reference_nlp(ref, yaml_path)
        spec = no.NlpSpec.from_yaml_dict(raw)
        rows, cols, f_g, f_j, f_f = lambdify_reference(opti, net64)
        w = golden_w(spec, P, seed=7).astype(np.float32).astype(np.float64)   # fp32-representable: identical inputs on both sides
        assert len(opti.w_symbols()) == spec.n_w, (len(opti.w_symbols()), spec.n_w)
        g = np.array([np.asarray(f_g(*wi), float) for wi in w])
        jv = np.array([np.asarray(f_j(*wi), float) for wi in w])
        fg = np.array([np.asarray(f_f(*wi), float) for wi in w])
        np.savez(GOLD / f"nlp_{name}.npz", yaml=np.array(yaml_path.name), w=w, g=g, jac_rows=rows, jac_cols=cols,
                 jac_vals=jv, f=fg[:, 0], grad_f=fg[:, 1:], lbg=np.array(opti.lbg), ubg=np.array(opti.ubg),
                 w_init_linear=opti.w_initial())
        print(f"{name}: n_w={spec.n_w} n_g={len(opti.g_rows)} nnz={len(rows)}")


def make_metrics_golden(ref: Path):
    """The reference's own core/metrics.py (numpy + torch only, importable here) on small seeded fields: a circle's exact SDF
    as target, smooth perturbations of it as predictions, plus the empty-band / empty-occupancy corner cases."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_metrics", ref / "src/nlotrajectories/core/metrics.py")
    rm = importlib.util.module_from_spec(spec); spec.loader.exec_module(rm)
    n = 96
    x = np.linspace(-1.0, 2.0, n); X, Y = np.meshgrid(x, x)
    target = np.hypot(X - 0.5, Y - 0.4) - 0.45
    out = {"X": X, "Y": Y, "target": target}
    rng = np.random.default_rng(7)
    cases = {"shifted": np.hypot(X - 0.53, Y - 0.38) - 0.43,
             "wavy": target + 0.03 * np.sin(7 * X) * np.cos(5 * Y),
             "noisy": target + 0.02 * rng.standard_normal(target.shape),
             "far": target + 5.0}                                        # no surface band, no occupancy in the prediction
    for name, pred in cases.items():
        eps = 0.04
        vals = [rm.mse(target, pred), rm.iou(target, pred, threshold=0.0), rm.hausdorff(pred, target, X, Y, eps=eps),
                rm.chamfer(pred, target, X, Y, eps=eps), rm.surface_loss(target, pred, eps=eps)]
        out["pred_" + name] = pred
        out["vals_" + name] = np.array([np.nan if v is None else float(v) for v in vals])
    out["eps"] = np.array(0.04)
    out["iou_both_empty"] = np.array(float(rm.iou(target + 5.0, target + 6.0)))
    np.savez_compressed(GOLD / "metrics_reference.npz", **out)
    print("wrote", GOLD / "metrics_reference.npz", {k: v for k, v in out.items() if k.startswith("vals_")})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--only", default="")
    ap.add_argument("--hessian", action="store_true", help="(re)generate only the Hessian-of-the-Lagrangian fixtures")
    ap.add_argument("--metrics", action="store_true", help="(re)generate only the SDF-quality metric fixtures")
    a = ap.parse_args()
    ref = Path(a.reference)
    GOLD.mkdir(parents=True, exist_ok=True)
    if a.metrics:
        make_metrics_golden(ref)
        return
    net = None if a.hessian else make_sdf_golden(ref)
    names = ["benchmark_1", "benchmark_2", "benchmark_3", "benchmark_4", "benchmark_6"]
    if a.only:
        names = a.only.split(",")
    if a.hessian:
        from oracle import sdf_oracle as _so
        make_hess_golden(ref, _so.from_npz(GOLD / "sdf_shipped_fourier128_weights.npz"), names)
        return
    make_nlp_golden(ref, net, names)
    make_hess_golden(ref, net, names)


if __name__ == "__main__":
    main()
