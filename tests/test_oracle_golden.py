"""CPU: the oracle against the golden vectors produced by the reference itself (oracle/make_golden.py)."""
import numpy as np
import pytest
import yaml

from conftest import GOLDEN, bench_yaml, close
from oracle import nlp_oracle as no
from oracle import sdf_oracle as so


def test_sdf_oracle_matches_reference_torchscript(shipped_net):
    z = np.load(GOLDEN / "sdf_shipped_fourier128.npz")
    P, sbar = z["P"].astype(np.float64), z["sbar"].astype(np.float64)
    net = shipped_net.astype(np.float64)
    s, J = so.value_jac(net, P)
    assert not close(s, z["value"]).any()
    assert not close(J, z["jac"]).any()
    assert not close(so.adj1(net, P, sbar), z["adj1"]).any()
    H = so.jac_adj1(net, P, sbar)
    # second derivatives carry |W0|^2 ~ 500: the bar is relative to the output scale (max |H| ~ 35 here)
    assert np.abs(H - z["jac_adj1"]).max() <= 1e-5 * max(1.0, np.abs(z["jac_adj1"]).max())


def test_sdf_oracle_survey_appendix_e(shipped_net):
    """SURVEY.md Appendix E known-answer table (fp32 TorchScript, seed 2.0 for the Hessian)."""
    pts = np.array([[0.0, 0.0], [0.5, 0.5], [1.0, 1.0]])
    want_s = [2.99533278e-01, -1.71203554e-01, 1.84726045e-01]
    want_j = [[-9.69587803e-01, -1.83145329e-01], [-2.81356037e-01, -6.89222097e-01], [-3.42374355e-01, 8.44195664e-01]]
    want_h00 = [-1.329027e+01, 9.568398e+00, -1.703431e+01]
    net = shipped_net.astype(np.float64)
    s, J = so.value_jac(net, pts)
    H = so.jac_adj1(net, pts, np.full(3, 2.0))
    np.testing.assert_allclose(s, want_s, atol=1e-6)
    np.testing.assert_allclose(J, want_j, atol=2e-6)
    np.testing.assert_allclose(H[:, 0, 0], want_h00, atol=5e-5)


@pytest.mark.parametrize("kind", ["mlp", "siren", "fourier"])
def test_sdf_oracle_derivatives_by_finite_differences(kind):
    net = {"mlp": so.synthetic_mlp(32, 2, act=so.ACT_TANH, dtype=np.float64),
           "siren": so.synthetic_siren(32, 1, omega0=5.0, dtype=np.float64),
           "fourier": so.synthetic_fourier(32, 1, scale=2.0, act=so.ACT_SIGMOID, dtype=np.float64)}[kind]
    rng = np.random.default_rng(3)
    P = rng.uniform(-0.5, 1.5, (50, 2))
    sbar = rng.uniform(0.5, 1.5, 50)
    eps = 1e-6
    _, J = so.value_jac(net, P)
    H = so.jac_adj1(net, P, sbar)
    for d in range(2):
        e = np.zeros(2); e[d] = eps
        fd = (so.forward(net, P + e) - so.forward(net, P - e)) / (2 * eps)
        np.testing.assert_allclose(J[:, d], fd, atol=1e-6)
        fdh = (so.adj1(net, P + e, sbar) - so.adj1(net, P - e, sbar)) / (2 * eps)
        np.testing.assert_allclose(H[:, :, d], fdh, atol=1e-5)


@pytest.mark.parametrize("name", ["benchmark_1", "benchmark_2", "benchmark_3", "benchmark_4", "benchmark_5", "benchmark_6"])
def test_nlp_oracle_matches_reference_assembly(name, shipped_net):
    z = np.load(GOLDEN / f"nlp_{name}.npz")
    spec = no.NlpSpec.from_yaml_dict(yaml.safe_load(open(bench_yaml(name))))
    net = shipped_net.astype(np.float64)
    g, jv = no.eval_g_jac(spec, z["w"], lambda P: so.value_jac(net, P))
    f, gr = no.eval_f_grad(spec, z["w"])
    rows, cols, _ = no.jac_pattern(spec)
    lb, ub = no.bounds(spec)
    assert g.shape[1] == spec.n_g == z["g"].shape[1]
    assert np.array_equal(rows, z["jac_rows"]) and np.array_equal(cols, z["jac_cols"])
    np.testing.assert_allclose(g, z["g"], atol=1e-12)
    np.testing.assert_allclose(jv, z["jac_vals"], atol=1e-12)
    np.testing.assert_allclose(f, z["f"], atol=1e-12)
    np.testing.assert_allclose(gr, z["grad_f"], atol=1e-12)
    assert np.array_equal(lb, z["lbg"]) and np.array_equal(ub, z["ubg"])
    np.testing.assert_allclose(spec.n_w, z["w"].shape[1])


def test_nlp_sizes_survey_appendix_a3():
    want = {"benchmark_1": (285, 329, 690), "benchmark_3": (326, 371, 974), "benchmark_4": (646, 731, 1934),
            "benchmark_6": (727, 1057, 3225)}
    for name, (n_w, n_g, nnz) in want.items():
        spec = no.NlpSpec.from_yaml_dict(yaml.safe_load(open(bench_yaml(name))))
        assert (spec.n_w, spec.n_g, len(no.jac_pattern(spec)[0])) == (n_w, n_g, nnz)


def test_nlp_oracle_jacobian_by_finite_differences(shipped_net):
    spec = no.NlpSpec.from_yaml_dict(yaml.safe_load(open(bench_yaml("benchmark_6"))))
    net = so.synthetic_mlp(16, 1, act=so.ACT_TANH, dtype=np.float64)
    sdf = lambda P: so.value_jac(net, P)
    rng = np.random.default_rng(0)
    w = no.multistart_guess(spec, 1) + rng.normal(0, 0.05, (1, spec.n_w))
    g, jv = no.eval_g_jac(spec, w, sdf)
    rows, cols, _ = no.jac_pattern(spec)
    J = np.zeros((spec.n_g, spec.n_w)); J[rows, cols] = jv[0]
    eps = 1e-6
    for v in rng.choice(spec.n_w, 40, replace=False):
        e = np.zeros((1, spec.n_w)); e[0, v] = eps
        fd = (no.eval_g_jac(spec, w + e, sdf)[0] - no.eval_g_jac(spec, w - e, sdf)[0])[0] / (2 * eps)
        np.testing.assert_allclose(J[:, v], fd, atol=2e-6)
    f, gr = no.eval_f_grad(spec, w)
    for v in rng.choice(spec.n_w, 40, replace=False):
        e = np.zeros((1, spec.n_w)); e[0, v] = eps
        fd = (no.eval_f_grad(spec, w + e)[0] - no.eval_f_grad(spec, w - e)[0])[0] / (2 * eps)
        np.testing.assert_allclose(gr[0, v], fd, atol=2e-6)


@pytest.mark.parametrize("name", ["benchmark_1", "benchmark_2", "benchmark_3", "benchmark_4", "benchmark_5", "benchmark_6"])
def test_hessian_oracle_matches_reference_expressions(name, shipped_net):
    """Hessian of the Lagrangian: the numpy restatement against second derivatives of the reference's own recorded
    expressions (oracle/make_golden.py --hessian): identical structural pattern (upper triangle, CCS) and values."""
    z = np.load(GOLDEN / f"nlp_hess_{name}.npz")
    spec = no.NlpSpec.from_yaml_dict(yaml.safe_load(open(bench_yaml(name))))
    net = shipped_net.astype(np.float64)
    rows, cols = no.hess_pattern(spec)
    assert np.array_equal(rows, z["hess_rows"]) and np.array_equal(cols, z["hess_cols"])
    w = z["w"]
    lam = np.broadcast_to(z["lam"], (w.shape[0], spec.n_g))
    sdf_h = lambda Q: (lambda H: np.stack([H[:, 0, 0], H[:, 0, 1], H[:, 1, 1]], -1))(so.jac_adj1(net, Q, np.ones(len(Q))))
    H = no.eval_hess_lag(spec, w, float(z["sigma"]), lam, lambda Q: so.value_jac(net, Q), sdf_h)
    scale = max(1.0, np.abs(z["hess_vals"]).max())
    assert np.abs(H - z["hess_vals"]).max() <= 1e-9 * scale, np.abs(H - z["hess_vals"]).max()
