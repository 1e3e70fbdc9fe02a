"""Batched interior-point solver: the caller on top of the evaluation hot path.
CPU: the solver logic on the numpy oracle.  GPU: the same solver over the CUDA evaluation + Hessian kernels."""
import numpy as np
import pytest
import yaml

from conftest import bench_yaml
from oracle import nlp_oracle as no
from solver_util import OracleEvaluator

B1_OPT = 1.50352      # benchmark_1 optimum at the reference's tolerance (tol 1e-4, final barrier 1e-5); SLSQP: 1.50342


def _solve_cpu(P, max_iter=200):
    import torch
    from nlotrajectories_b200.solver import BatchedIPSolver
    spec = no.NlpSpec.from_yaml_dict(yaml.safe_load(open(bench_yaml("benchmark_1"))))
    lb, ub = no.bounds(spec)
    w0 = no.multistart_guess(spec, P)
    res = BatchedIPSolver(OracleEvaluator(spec), lb, ub, max_iter=max_iter).solve(torch.from_numpy(w0))
    return spec, lb, ub, res


def test_interior_point_on_the_oracle_benchmark_1():
    spec, lb, ub, res = _solve_cpu(2)
    assert bool(res.converged.all())
    f = res.f.numpy()
    assert np.all(np.abs(f - B1_OPT) < 2e-4), f
    assert np.all(res.violation.numpy() < 1e-6)
    # first-order optimality with the returned multipliers: grad f + J^T lam = 0
    g, jv = no.eval_g_jac(spec, res.w.numpy())
    _, gr = no.eval_f_grad(spec, res.w.numpy())
    rows, cols, _ = no.jac_pattern(spec)
    for i in range(2):
        r = gr[i].copy()
        np.add.at(r, cols, jv[i] * res.lam.numpy()[i, rows])
        assert np.abs(r).max() < 5e-3


@pytest.mark.gpu
def test_interior_point_on_the_gpu_path_matches_cpu_solution(library):
    import torch
    from nlotrajectories_b200.config import Config
    from nlotrajectories_b200.problem import NlpProblem
    from nlotrajectories_b200.solver import BatchedIPSolver, DeviceEvaluator
    prob = NlpProblem.from_config(Config.load(bench_yaml("benchmark_1")), None)
    lb, ub = prob.bounds()
    P = 16
    w0 = prob.multistart_guess(P).astype(np.float64)
    res = BatchedIPSolver(DeviceEvaluator(prob), lb, ub, max_iter=200).solve(torch.from_numpy(w0).cuda())
    ok = res.converged.cpu().numpy()
    f = res.f.cpu().numpy()
    assert ok.mean() >= 0.5, ok
    assert np.all(np.abs(f[ok] - B1_OPT) < 3e-4), f[ok]
    assert np.all(res.violation.cpu().numpy()[ok] < 1e-5)


@pytest.mark.gpu
def test_run_benchmark_run_returns_reference_shapes(library):
    """RunBenchmark.run() mirrors core/runner.py:9-153: (X_opt (nx, N+1), U_opt (nu, N), result, X_init, status)."""
    from nlotrajectories_b200.runner import RunBenchmark
    rb = RunBenchmark(dynamics="point_2nd", geometry="dot", x0=[0.0, 0.0, 0.0, 0.0], x_goal=[1.0, 1.0, 0.0, 0.0], N=40, dt=0.1,
                      sdf_func=[(0.5, 0.5, 0.2, 0.05)], control_bounds=((-1.0, 1.0), (-1.0, 1.0)), use_slack=True, slack_penalty=50,
                      enforce_heading=False)                       # benchmark_1_dot_circle.yaml
    X_opt, U_opt, res, X_init, status = rb.run(P=8, max_iter=200)
    assert X_opt.shape == (4, 41) and U_opt.shape == (2, 40) and X_init.shape == (41, 4)
    assert status == "success"
    np.testing.assert_allclose(X_opt[:, 0], [0, 0, 0, 0], atol=1e-5)
    np.testing.assert_allclose(X_opt[[0, 1, 3], -1], [1, 1, 0], atol=1e-5)          # terminal row skips state 2 (runner.py:54-56)
    assert np.all(np.abs(U_opt) <= 1.0 + 1e-5)
    d = np.hypot(X_opt[0] - 0.5, X_opt[1] - 0.5)
    assert d.min() >= 0.25 - 1e-4                                                     # radius + margin
    path = np.sqrt(np.diff(X_opt[0]) ** 2 + np.diff(X_opt[1]) ** 2 + 1e-8).sum()
    assert abs(path - B1_OPT) < 3e-4


def test_ackermann_with_slack_converges_from_rest_on_the_oracle():
    """benchmark_5: the linearised dynamics are inconsistent at the all-zero guess (vehicle at rest); the least-squares equality
    multipliers keep the Hessian of the Lagrangian bounded and both starts reach tol 1e-4."""
    import torch
    from nlotrajectories_b200.solver import BatchedIPSolver
    spec = no.NlpSpec.from_yaml_dict(yaml.safe_load(open(bench_yaml("benchmark_5"))))
    lb, ub = no.bounds(spec)
    w0 = no.multistart_guess(spec, 2)
    res = BatchedIPSolver(OracleEvaluator(spec), lb, ub, max_iter=200).solve(torch.from_numpy(w0))
    assert bool(res.converged.all()), res.kkt_error
    assert np.all(np.abs(res.f.numpy() - 1.6205) < 2e-3), res.f
    assert np.all(res.violation.numpy() < 1e-4)


def test_elastic_evaluator_reproduces_the_plain_solution():
    """Exact l1 penalty on the one-sided inequality rows: with a penalty above the multipliers the elastic variables end at
    zero (to the barrier's share) and the objective is the plain problem's."""
    import torch
    from nlotrajectories_b200.solver import BatchedIPSolver, ElasticEvaluator
    spec = no.NlpSpec.from_yaml_dict(yaml.safe_load(open(bench_yaml("benchmark_1"))))
    lb, ub = no.bounds(spec)
    w0 = torch.from_numpy(no.multistart_guess(spec, 2))
    ev = ElasticEvaluator(OracleEvaluator(spec), lb, ub, penalty=100.0)
    assert ev.n_w == spec.n_w + ev.m and ev.n_g == spec.n_g + ev.m and ev.m > 0
    w1 = ev.initial(w0)
    # the wrapped functions: value rows shifted by p, Jacobian / Hessian blocks consistent with finite differences
    f, grad, g, J = ev.eval(w1)
    h = 1e-6
    d = torch.from_numpy(np.random.default_rng(0).standard_normal(w1.shape))
    f2, _, g2, _ = ev.eval(w1 + h * d, want_jac=False)
    f0, _, g0, _ = ev.eval(w1 - h * d, want_jac=False)
    assert torch.allclose((f2 - f0) / (2 * h), (grad * d).sum(1), atol=1e-5)
    assert torch.allclose((g2 - g0) / (2 * h), torch.einsum("prw,pw->pr", J, d), atol=1e-5)
    res = BatchedIPSolver(ev, ev.lbg, ev.ubg, max_iter=200).solve(w1)
    assert bool(res.converged.all())
    w, p = ev.split(res.w)
    assert float(p.max()) < 1e-4
    assert np.all(np.abs(res.f.numpy() - B1_OPT) < 2e-2)


def test_solve_elastic_reports_in_original_terms():
    import torch
    from nlotrajectories_b200.solver import solve_elastic
    spec = no.NlpSpec.from_yaml_dict(yaml.safe_load(open(bench_yaml("benchmark_1"))))
    lb, ub = no.bounds(spec)
    w0 = torch.from_numpy(no.multistart_guess(spec, 2))
    res = solve_elastic(OracleEvaluator(spec), lb, ub, w0, penalty=100.0, max_iter=200)
    assert res.w.shape == (2, spec.n_w) and res.lam.shape == (2, spec.n_g)
    assert bool(res.converged.all())
    f_orig, _ = no.eval_f_grad(spec, res.w.numpy())
    np.testing.assert_allclose(res.f.numpy(), f_orig, atol=1e-9)
    assert np.all(np.abs(f_orig - B1_OPT) < 2e-3)


@pytest.mark.gpu
def test_benchmark_6_solves_on_the_gpu_path_with_the_trained_network(library):
    """The headline YAML end to end: learned ReLU SDF (weights trained by train.py at the YAML's sample count), RRT guesses lifted
    from the path, batched interior point over the tensor-core SDF kernels (values, Jacobian and the fused Hessian route)."""
    import torch
    from gpu_util import to_weights
    from oracle import sdf_oracle as so
    from conftest import GOLDEN
    from nlotrajectories_b200.config import Config
    from nlotrajectories_b200.initializer import rrt_multistart
    from nlotrajectories_b200.problem import NlpProblem
    from nlotrajectories_b200.sdf import LearnedSDF
    from nlotrajectories_b200.solver import BatchedIPSolver, DeviceEvaluator
    cfg = Config.load(bench_yaml("benchmark_6"))
    w0 = rrt_multistart(cfg, 8, lift=True, workers=1).astype(np.float64)      # in-process: this pytest process already holds a CUDA context
    model = LearnedSDF(to_weights(so.from_npz(str(GOLDEN / "sdf_benchmark_6_relu128.npz"))))
    prob = NlpProblem.from_config(cfg, model)
    lb, ub = prob.bounds()
    res = BatchedIPSolver(DeviceEvaluator(prob), lb, ub, max_iter=200).solve(torch.from_numpy(w0).cuda())
    usable = (res.converged | res.stalled).cpu().numpy() & (res.violation.cpu().numpy() <= 1e-4)
    assert usable.mean() >= 0.5, (res.kkt_error, res.violation)
    f = res.f.cpu().numpy()[usable]
    assert abs(f.min() - 9.30927) < 5e-3, f
    # the solution respects the learned SDF at every footprint point and the control bounds
    g = DeviceEvaluator(prob).eval(res.w, want_jac=False)[2].cpu().numpy()[usable]
    assert np.all(g >= lb[None] - 1e-4) and np.all(g <= ub[None] + 1e-4)
    model.close()


def test_compacting_the_working_set_does_not_change_any_start():
    """Finished starts leave the working set every 10 iterations; each start's iterates depend on that start alone."""
    import torch
    from nlotrajectories_b200.solver import BatchedIPSolver
    spec = no.NlpSpec.from_yaml_dict(yaml.safe_load(open(bench_yaml("benchmark_1"))))
    lb, ub = no.bounds(spec)
    w0 = torch.from_numpy(no.multistart_guess(spec, 6))
    a = BatchedIPSolver(OracleEvaluator(spec), lb, ub, max_iter=200, compact=True).solve(w0)
    b = BatchedIPSolver(OracleEvaluator(spec), lb, ub, max_iter=200, compact=False).solve(w0)
    assert bool(a.converged.all()) and bool(b.converged.all())
    assert len(set(a.iterations.tolist())) > 1                    # the starts do finish at different iterations
    assert torch.equal(a.iterations, b.iterations)
    np.testing.assert_allclose(a.w.numpy(), b.w.numpy(), atol=1e-9)
    np.testing.assert_allclose(a.f.numpy(), b.f.numpy(), atol=1e-12)
    np.testing.assert_allclose(a.lam.numpy(), b.lam.numpy(), atol=1e-6)
