"""The device interior-point solver (csrc/ip_solver.cu, C ABI nlo_ip_*).

CPU (-m "not gpu"): the solver's per-problem bodies, table builders and iteration loop - the very code the CUDA kernels wrap -
compiled for the host (tests/tools/ip_host_emul.cpp) and driven by the numpy oracle: the block-tridiagonal Newton step against a
dense solve of the same matrix, and whole solves against the dense torch solver.
GPU (-m gpu): the CUDA path through the C ABI against dense fp64 linear algebra on the same inputs and against the dense solver."""
import numpy as np
import pytest
import yaml

import ip_emul
from conftest import GOLDEN, bench_yaml
from oracle import nlp_oracle as no
from oracle import sdf_oracle as so
from solver_util import OracleEvaluator

BENCHES = ["benchmark_1", "benchmark_2", "benchmark_3", "benchmark_4", "benchmark_5", "benchmark_6"]


def _spec(name):
    return no.NlpSpec.from_yaml_dict(yaml.safe_load(open(bench_yaml(name))))


def _dense_step(op, jac, hess, omega, rhs, delta):
    """(H + J^T diag(omega) J + delta I) dw = rhs by dense fp64 Cholesky, problem by problem."""
    P = jac.shape[0]
    cols = np.repeat(np.arange(op.n_w), np.diff(op.jcolind))
    hcols = np.repeat(np.arange(op.n_w), np.diff(op.hcolind))
    dw = np.empty((P, op.n_w))
    for p in range(P):
        J = np.zeros((op.n_g, op.n_w)); J[op.jrow, cols] = jac[p].astype(np.float64)
        H = np.zeros((op.n_w, op.n_w)); H[op.hrow, hcols] = hess[p].astype(np.float64); H[hcols, op.hrow] = hess[p].astype(np.float64)
        K = H + J.T @ (omega[p][:, None] * J) + delta[p] * np.eye(op.n_w)
        c = np.linalg.cholesky(K)
        dw[p] = np.linalg.solve(c.T, np.linalg.solve(c, rhs[p]))
    return dw


def _random_kkt_inputs(op, P, seed, rho=1e8):
    """Inputs with the magnitudes the solver sees: evaluated Jacobian values, a positive definite Hessian stand-in on the
    structural pattern (diagonal dominant), row weights rho on the equalities and barrier weights on the inequalities."""
    rng = np.random.default_rng(seed)
    spec = op.spec
    w = no.multistart_guess(spec, P) + rng.normal(0, 0.05, (P, spec.n_w))
    _, jac = no.eval_g_jac(spec, w, op.sdf)
    hcols = np.repeat(np.arange(op.n_w), np.diff(op.hcolind))
    hess = rng.normal(0, 0.2, (P, op.nnzh))
    diag = op.hrow == hcols
    hess[:, diag] = 1.0 + np.abs(hess[:, diag])
    eq = op.lb == op.ub
    omega = np.where(eq[None, :], rho, 10.0 ** rng.uniform(-3, 4, (P, op.n_g)))
    rhs = rng.normal(0, 1.0, (P, op.n_w)) * 10.0 ** rng.uniform(-2, 4, (P, 1))
    return jac.astype(np.float32), hess.astype(np.float32), omega, rhs


@pytest.mark.parametrize("name", BENCHES)
def test_block_tridiagonal_step_equals_dense_step(name, shipped_net):
    """Every shipped YAML's stage structure: the banded Newton step equals a dense Cholesky solve of the same condensed matrix."""
    spec = _spec(name)
    op = ip_emul.OracleProblem(spec, shipped_net if spec.sdf_mode == "l4casadi" else None)
    P = 3
    jac, hess, omega, rhs = _random_kkt_inputs(op, P, seed=1)
    delta = np.full(P, 50.0)                                   # enough that the stand-in Hessian is positive definite: no retry
    dw, d_out = ip_emul.kkt_step(op, jac, hess, omega, rhs, delta)
    assert np.array_equal(d_out, delta)
    ref = _dense_step(op, jac, hess, omega, rhs, delta)
    scale = np.abs(ref).max(axis=1, keepdims=True)
    assert np.abs(dw - ref).max() <= 1e-8 * scale.max(), np.abs(dw - ref).max() / scale.max()


def test_inertia_correction_grows_delta_until_positive_definite():
    """An indefinite Hessian block makes the first factorisations fail: delta grows x8 from 1e-4 until the Cholesky succeeds, then
    the problem is solved once more with twice that value (the reference solver's rule), and the step solves that system."""
    spec = _spec("benchmark_1")
    op = ip_emul.OracleProblem(spec)
    P = 2
    jac, hess, omega, rhs = _random_kkt_inputs(op, P, seed=2)
    hcols = np.repeat(np.arange(op.n_w), np.diff(op.hcolind))
    hess[1, op.hrow == hcols] = -3.0                           # problem 1: negative curvature everywhere the constraints leave free
    rhs *= 1e-3 / np.abs(rhs).max()                            # small steps: only the inertia test decides (a step beyond 1e3 is refused too)
    dw, d_out = ip_emul.kkt_step(op, jac, hess, omega, rhs, np.array([50.0, 0.0]))
    assert d_out[0] == 50.0                                    # problem 0 needs nothing beyond its previous value
    k = np.log(d_out[1] / 2.0 / 1e-4) / np.log(8.0)
    assert abs(k - round(k)) < 1e-9 and d_out[1] / 2.0 > 3.0   # 1e-4 * 8^k, beyond the negative curvature

    def positive_definite(delta):
        try:
            _dense_step(op, jac[1:], hess[1:], omega[1:], rhs[1:], np.array([delta]))
            return True
        except np.linalg.LinAlgError:
            return False
    assert positive_definite(d_out[1] / 2.0) and not positive_definite(d_out[1] / 16.0)
    ref = _dense_step(op, jac, hess, omega, rhs, d_out)
    assert np.abs(dw - ref).max() <= 1e-7 * np.abs(ref).max()


@pytest.mark.parametrize("name,P,max_iter", [("benchmark_1", 3, 200), ("benchmark_5", 2, 200)])
def test_emulated_device_solver_follows_the_dense_solver(name, P, max_iter):
    """Whole solves: same iteration counts and the same solutions as the dense torch solver on the same (oracle) functions; the
    only difference is that the device loop evaluates at the fp32 rounding of w, like the CUDA evaluation kernels."""
    import torch
    from nlotrajectories_b200.solver import BatchedIPSolver
    spec = _spec(name)
    lb, ub = no.bounds(spec)
    w0 = no.multistart_guess(spec, P)
    r = ip_emul.solve(ip_emul.OracleProblem(spec), w0, max_iter=max_iter)
    ref = BatchedIPSolver(OracleEvaluator(spec), lb, ub, max_iter=max_iter).solve(torch.from_numpy(w0))
    assert np.array_equal(r["status"] == 1, ref.converged.numpy()) and r["status"].all()
    assert np.array_equal(r["iterations"], ref.iterations.numpy())
    assert np.abs(r["f"] - ref.f.numpy()).max() < 1e-5
    assert np.abs(r["w"] - ref.w.numpy()).max() < 5e-3
    assert r["violation"].max() < 1e-4
    assert r["stats"]["hessians"] <= r["stats"]["iterations"] and r["stats"]["compactions"] >= 1


def test_emulated_device_solver_with_learned_sdf_benchmark_3():
    """Rectangular unicycle around the learned (ReLU) SDF: the structure with footprint rows and slack."""
    import torch
    from nlotrajectories_b200.solver import BatchedIPSolver
    spec = _spec("benchmark_3")
    net = so.from_npz(GOLDEN / "sdf_benchmark_3_relu128.npz")
    lb, ub = no.bounds(spec)
    w0 = no.multistart_guess(spec, 1)
    r = ip_emul.solve(ip_emul.OracleProblem(spec, net), w0, max_iter=60)
    ref = BatchedIPSolver(OracleEvaluator(spec, net), lb, ub, max_iter=60).solve(torch.from_numpy(w0))
    assert r["status"][0] == 1 and bool(ref.converged[0])
    assert abs(r["f"][0] - ref.f[0].item()) < 1e-4


# ---------------------------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------------------------
def _gpu_problem(name, net):
    from gpu_util import to_weights
    from nlotrajectories_b200.config import Config
    from nlotrajectories_b200.problem import NlpProblem
    from nlotrajectories_b200.sdf import LearnedSDF
    cfg = Config.load(bench_yaml(name))
    model = LearnedSDF(to_weights(net)) if cfg.solver.mode == "l4casadi" else None
    return NlpProblem.from_config(cfg, model)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["benchmark_1", "benchmark_4", "benchmark_5", "benchmark_6"])
def test_device_newton_step_equals_dense_step_on_64_starts(name, library, capsys):
    """nlo_ip_kkt_step (assembly + block-tridiagonal Cholesky kernels) against torch's dense fp64 Cholesky of the same condensed
    matrix, built from the CUDA evaluation and Hessian of 64 starts with the weights of an interior-point iteration."""
    import torch
    from nlotrajectories_b200.solver import DeviceEvaluator, DeviceIPSolver
    net = so.from_npz(GOLDEN / f"sdf_{name}_relu128.npz") if name in ("benchmark_4", "benchmark_6") else None
    prob = _gpu_problem(name, net)
    P = 64
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(0)
    w = prob.multistart_guess(P).astype(np.float64) + rng.normal(0, 0.02, (P, prob.n_w))
    wd = torch.from_numpy(w).to(dev)
    ev = DeviceEvaluator(prob)
    f, grad, g, J = ev.eval(wd)
    lb, ub = prob.bounds()
    eq = torch.from_numpy(lb == ub).to(dev)
    lam = torch.from_numpy(rng.normal(0, 1.0, (P, prob.n_g))).to(dev)
    H = ev.hess(wd, torch.ones(P, dtype=torch.float64, device=dev), lam)
    omega = torch.where(eq[None, :], torch.full((P, prob.n_g), 1e8, dtype=torch.float64, device=dev),
                        torch.from_numpy(10.0 ** rng.uniform(-3, 4, (P, prob.n_g))).to(dev))
    rhs = torch.from_numpy(rng.normal(0, 1.0, (P, prob.n_w))).to(dev)
    # the same fp32 arrays the kernels consume
    ws = wd.to(torch.float32).T.contiguous()
    gq, jq, fq, grq = prob.alloc_outputs(P)
    prob.eval_device(ws, gq, jq, fq, grq)
    hq = prob.eval_hess_device(ws, lam.to(torch.float32).T.contiguous())
    solver = DeviceIPSolver(prob, max_problems=P)
    delta = torch.full((P,), 1e-2, dtype=torch.float64, device=dev)
    dw, d_out = solver.kkt_step(jq, hq, omega.T.contiguous(), rhs.T.contiguous(), delta)
    torch.cuda.synchronize()
    K = H + torch.einsum("prw,pr,prv->pwv", J, omega, J)
    eye = torch.eye(prob.n_w, dtype=torch.float64, device=dev)
    Lc, info = torch.linalg.cholesky_ex(K + d_out[:, None, None] * eye)
    assert int(info.abs().max()) == 0
    ref = torch.cholesky_solve(rhs[:, :, None], Lc)[:, :, 0]
    err = (dw.T - ref).abs().amax(1) / ref.abs().amax(1)
    with capsys.disabled():
        print(f"\n[{name}] banded vs dense Newton step, 64 starts: max relative difference {err.max().item():.2e}; delta out "
              f"{d_out.min().item():.1e}..{d_out.max().item():.1e}")
    # rho = 1e8 on the equality rows puts the condition number of the matrix at 1e10 and beyond: two fp64 eliminations in different
    # orders agree to ~1e-8 .. 1e-7 (measured: benchmark_1 3.5e-8, benchmark_4 1.2e-7), and both leave residuals at rounding level
    assert err.max().item() <= 1e-6
    Kd = K + d_out[:, None, None] * eye
    res_b = (torch.einsum("pwv,pv->pw", Kd, dw.T.contiguous()) - rhs).abs().amax(1)
    res_d = (torch.einsum("pwv,pv->pw", Kd, ref) - rhs).abs().amax(1)
    assert bool((res_b <= 10.0 * res_d + 1e-9 * rhs.abs().amax(1)).all()), (res_b.max().item(), res_d.max().item())


@pytest.mark.gpu
@pytest.mark.parametrize("name,P", [("benchmark_1", 64), ("benchmark_5", 64), ("benchmark_3", 64)])
def test_device_solver_matches_dense_solver_final_objectives(name, P, library, capsys):
    """64 starts solved by the device solver and by the dense torch solver over the same CUDA evaluation: the starts both
    converge end at objectives within 1e-4, and the device solver converges at least as many."""
    import torch
    from nlotrajectories_b200.solver import BatchedIPSolver, DeviceEvaluator, DeviceIPSolver
    net = so.from_npz(GOLDEN / f"sdf_{name}_relu128.npz") if name == "benchmark_3" else None
    prob = _gpu_problem(name, net)
    lb, ub = prob.bounds()
    w0 = prob.multistart_guess(P).astype(np.float64)
    dev_res = DeviceIPSolver(prob, max_problems=P, max_iter=200).solve(w0)
    ref = BatchedIPSolver(DeviceEvaluator(prob), lb, ub, max_iter=200).solve(torch.from_numpy(w0).cuda())
    a, b = dev_res.converged.numpy(), ref.converged.cpu().numpy()
    both = a & b
    df = np.abs(dev_res.f.numpy() - ref.f.cpu().numpy())
    with capsys.disabled():
        print(f"\n[{name}] device solver converged {a.sum()}/{P} (+{int(dev_res.stalled.sum())} stalled-feasible), dense solver {b.sum()}/{P}; "
              f"max |f_dev - f_dense| over the {both.sum()} common: {df[both].max() if both.any() else float('nan'):.2e}")
    assert both.sum() >= 0.8 * b.sum() and both.sum() >= 16
    assert np.median(df[both]) <= 1e-5 and (df[both] <= 1e-4).mean() >= 0.9      # a few starts may end in neighbouring local optima
    assert a.sum() >= 0.85 * b.sum()          # (the ReLU kinks make the outcome of individual starts sensitive to rounding)
    assert dev_res.violation.numpy()[a].max() <= 1e-4


@pytest.mark.gpu
def test_device_solver_benchmark_4_at_4096_starts(library, capsys):
    """BASELINE config 'benchmark_4 x 4096 multi-starts on one GPU': the whole batch solves on the device."""
    import time
    from nlotrajectories_b200.solver import DeviceIPSolver
    net = so.from_npz(GOLDEN / "sdf_benchmark_4_relu128.npz")
    prob = _gpu_problem("benchmark_4", net)
    P = 4096
    w0 = prob.multistart_guess(P).astype(np.float64)
    solver = DeviceIPSolver(prob, max_problems=P, max_iter=150)
    t0 = time.time()
    res = solver.solve(w0)
    dt = time.time() - t0
    ok = res.converged.numpy(); st = res.stalled.numpy(); v = res.violation.numpy(); f = res.f.numpy()
    usable = ok | (st & (v <= 1e-4))
    with capsys.disabled():
        print(f"\n[benchmark_4 x {P}] {dt:.1f} s, {solver.stats}; converged {ok.mean() * 100:.1f} %, stalled-feasible {st.mean() * 100:.1f} %, "
              f"best objective {f[usable].min() if usable.any() else float('nan'):.6f}")
    assert usable.mean() >= 0.3
    assert np.isfinite(f[usable]).all() and v[usable].max() <= 1e-4
