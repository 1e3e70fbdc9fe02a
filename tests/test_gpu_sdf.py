"""GPU parity: the CUDA learned-SDF path (through the C ABI) against the oracle and the golden vectors."""
import ctypes as C

import numpy as np
import pytest

from conftest import GOLDEN, REPO, close
from gpu_util import kink_mask, sample_points, to_weights
from oracle import sdf_oracle as so

pytestmark = pytest.mark.gpu
TOL = 1e-5     # |a-b| <= TOL * max(1, |b|)   (BASELINE.md parity gate, FP32)


@pytest.fixture(scope="module")
def torch_cuda(library):
    import torch
    assert torch.cuda.is_available()
    return torch


def run_device(model, torch, P, sbar=None, want_jac=True):
    x = torch.from_numpy(np.ascontiguousarray(P[:, 0])).cuda()
    y = torch.from_numpy(np.ascontiguousarray(P[:, 1])).cuda()
    sb = torch.from_numpy(sbar).cuda() if sbar is not None else None
    s, jx, jy = model.eval(x, y, sb, want_jac=want_jac)
    torch.cuda.synchronize()
    return s.cpu().numpy(), (np.stack([jx.cpu().numpy(), jy.cpu().numpy()], 1) if want_jac else None)


NETS = {
    "shipped_fourier128": None,
    "relu128": lambda: so.synthetic_mlp(128, 1, seed=0),
    "relu64": lambda: so.synthetic_mlp(64, 1, seed=1),
    "relu256": lambda: so.synthetic_mlp(256, 1, seed=2),
    "relu32x3": lambda: so.synthetic_mlp(32, 3, seed=3),
    "tanh64x2": lambda: so.synthetic_mlp(64, 2, seed=4, act=so.ACT_TANH),
    "sigmoid48": lambda: so.synthetic_mlp(48, 1, seed=5, act=so.ACT_SIGMOID),
    "leaky64": lambda: so.synthetic_mlp(64, 1, seed=6, act=so.ACT_LEAKY_RELU),
    "siren64x2": lambda: so.synthetic_siren(64, 2, omega0=30.0, seed=7),
    "fourier64_tanh": lambda: so.synthetic_fourier(64, 1, scale=3.0, seed=8, act=so.ACT_TANH),
    "relu16x0": lambda: so.synthetic_mlp(16, 0, seed=9),
    "fourier256_relu": lambda: so.synthetic_fourier(256, 1, scale=2.0, seed=10),      # H = 256 tensor path, cos features
    "fourier128_relu": lambda: so.synthetic_fourier(128, 1, scale=2.0, seed=12),
    "tanh128": lambda: so.synthetic_mlp(128, 1, seed=13, act=so.ACT_TANH),            # generic 3-pass reverse GEMM, SFU tanh
    "sigmoid128": lambda: so.synthetic_mlp(128, 1, seed=15, act=so.ACT_SIGMOID),
    "leaky128": lambda: so.synthetic_mlp(128, 1, seed=16, act=so.ACT_LEAKY_RELU),
    "siren64x1": lambda: so.synthetic_siren(64, 1, omega0=30.0, seed=17),             # (sin, sin) on the tensor path
    "fourier128_sigmoid": lambda: so.synthetic_fourier(128, 1, scale=2.0, seed=18, act=so.ACT_SIGMOID),
    "fourier64_leaky": lambda: so.synthetic_fourier(64, 1, scale=2.0, seed=19, act=so.ACT_LEAKY_RELU),
    "tanh_sigmoid_mix64": None,                                                      # placeholder replaced below: run-time (generic) instantiation
}
NETS["fourier128_leaky"] = lambda: so.synthetic_fourier(128, 1, scale=2.0, seed=17, act=so.ACT_LEAKY_RELU)   # masked reverse GEMM + column sums
NETS["tanh_leaky_mix64"] = lambda: (lambda n: so.SdfNet(n.kind, n.W0, n.b0, n.hidden, n.w_out, n.b_out, so.ACT_TANH, so.ACT_LEAKY_RELU, n.p0, n.p))(
    so.synthetic_mlp(64, 1, seed=18, act=so.ACT_TANH))                                                           # leaky hidden layer on the generic kernel
NETS["tanh_sigmoid_mix64"] = lambda: (lambda n: so.SdfNet(n.kind, n.W0, n.b0, n.hidden, n.w_out, n.b_out, so.ACT_TANH, so.ACT_SIGMOID, n.p0, n.p))(
    so.synthetic_mlp(64, 1, seed=20, act=so.ACT_TANH))


# two and three H x H matrices (core/config.py:203-212 defaults; the YAMLs' fourier / siren depth): the deep tensor-tile kernel
NETS.update({
    "relu128x2": lambda: so.synthetic_mlp(128, 2, seed=21),
    "relu64x2": lambda: so.synthetic_mlp(64, 2, seed=22),                              # ModelConfig default: 64 wide, 3 layers
    "fourier128x2_relu": lambda: so.synthetic_fourier(128, 2, scale=2.0, seed=23),     # model.type fourier, num_hidden_layers 2
    "siren128x2": lambda: so.synthetic_siren(128, 2, omega0=30.0, seed=24),            # model.type siren, num_hidden_layers 2
    "tanh128x2": lambda: so.synthetic_mlp(128, 2, seed=25, act=so.ACT_TANH),
    "leaky128x2": lambda: so.synthetic_mlp(128, 2, seed=26, act=so.ACT_LEAKY_RELU),
    "sigmoid64x2": lambda: so.synthetic_mlp(64, 2, seed=27, act=so.ACT_SIGMOID),
    "relu128x3": lambda: so.synthetic_mlp(128, 3, seed=28),
    "fourier128x3_tanh": lambda: so.synthetic_fourier(128, 3, scale=2.0, seed=29, act=so.ACT_TANH),
    "fourier64x3_relu": lambda: so.synthetic_fourier(64, 3, scale=2.0, seed=30),
    "siren64x3": lambda: so.synthetic_siren(64, 3, omega0=30.0, seed=31),
    "fourier128x2_leaky": lambda: so.synthetic_fourier(128, 2, scale=2.0, seed=32, act=so.ACT_LEAKY_RELU),
})
DEEP_TC = ["relu128x2", "relu64x2", "fourier128x2_relu", "siren128x2", "tanh128x2", "leaky128x2", "sigmoid64x2", "relu128x3", "fourier128x3_tanh",
           "fourier64x3_relu", "siren64x3", "fourier128x2_leaky", "tanh64x2", "siren64x2"]


@pytest.mark.parametrize("name", DEEP_TC)
def test_deep_networks_run_on_the_tensor_path(name, torch_cuda):
    from nlotrajectories_b200.sdf import LearnedSDF
    model = LearnedSDF(to_weights(NETS[name]()), precision="auto")
    assert model.precision == "tc3xf16"
    model.close()


@pytest.mark.parametrize("precision", ["fp32", "auto"])
@pytest.mark.parametrize("name", list(NETS))
def test_value_jacobian_adjoint_match_oracle(name, precision, shipped_net, torch_cuda):
    from nlotrajectories_b200.sdf import LearnedSDF
    net = shipped_net if NETS[name] is None else NETS[name]()
    model = LearnedSDF(to_weights(net), precision=precision)
    n = 100_003                                     # ragged: not a multiple of any tile
    P = sample_points(n, seed=11)
    sbar = np.random.default_rng(2).uniform(0.5, 1.5, n).astype(np.float32)
    n64 = net.astype(np.float64)
    s_ref, J_ref = so.value_jac(n64, P.astype(np.float64))
    s, J = run_device(model, torch_cuda, P)
    tie = kink_mask(net, P)
    assert tie.mean() < 1e-2, f"{tie.sum()} kink-adjacent points"
    # derivatives of oscillatory nets (SIREN omega0 = 30) are O(omega0^2): the bar is relative to the output scale
    jscale = max(1.0, np.abs(J_ref).max()) if net.act in (so.ACT_SIN,) else 1.0
    assert not close(s, s_ref, TOL).any(), np.abs(s - s_ref).max()
    bad = close(J / jscale, J_ref / jscale, TOL).any(axis=1) & ~tie
    assert not bad.any(), (bad.sum(), np.abs(J - J_ref)[~tie].max())
    A_ref = so.adj1(n64, P.astype(np.float64), sbar.astype(np.float64))
    _, A = run_device(model, torch_cuda, P, sbar)
    assert not (close(A / jscale, A_ref / jscale, TOL).any(axis=1) & ~tie).any()
    s_only, none = run_device(model, torch_cuda, P, want_jac=False)
    assert none is None and np.array_equal(s_only, s)
    model.close()


def test_golden_vectors_from_reference_torchscript(shipped_net, torch_cuda):
    from nlotrajectories_b200.sdf import LearnedSDF
    z = np.load(GOLDEN / "sdf_shipped_fourier128.npz")
    for precision in ("fp32", "auto"):
        model = LearnedSDF(to_weights(shipped_net), precision=precision)
        s, J = run_device(model, torch_cuda, z["P"])
        _, A = run_device(model, torch_cuda, z["P"], z["sbar"])
        assert not close(s, z["value"], TOL).any()
        assert not close(J, z["jac"], TOL).any()
        assert not close(A, z["adj1"], TOL).any()
        hxx, hxy, hyy = model.hess_host(z["P"][:, 0].copy(), z["P"][:, 1].copy(), z["sbar"])
        H = np.stack([np.stack([hxx, hxy], 1), np.stack([hxy, hyy], 1)], 1)
        assert np.abs(H - z["jac_adj1"]).max() <= TOL * max(1.0, np.abs(z["jac_adj1"]).max())
        model.close()


@pytest.mark.parametrize("name", ["tanh64x2", "siren64x2", "fourier64_tanh", "sigmoid48", "relu128",
                                  "tanh128", "sigmoid128", "siren64x1", "fourier128_sigmoid"])     # the last four: Hessian as GEMMs (sdf_tc_hess.cu)
def test_hessian_matches_oracle(name, torch_cuda):
    from nlotrajectories_b200.sdf import LearnedSDF
    net = NETS[name]()
    model = LearnedSDF(to_weights(net))
    P = sample_points(4099, seed=5)
    sbar = np.random.default_rng(3).uniform(0.5, 1.5, P.shape[0]).astype(np.float32)
    H_ref = so.jac_adj1(net.astype(np.float64), P.astype(np.float64), sbar.astype(np.float64))
    hxx, hxy, hyy = model.hess_host(P[:, 0].copy(), P[:, 1].copy(), sbar)
    scale = max(1.0, np.abs(H_ref).max())
    assert np.abs(hxx - H_ref[:, 0, 0]).max() <= 2e-5 * scale
    assert np.abs(hxy - H_ref[:, 0, 1]).max() <= 2e-5 * scale
    assert np.abs(hyy - H_ref[:, 1, 1]).max() <= 2e-5 * scale
    model.close()


@pytest.mark.parametrize("name", ["shipped_fourier128", "fourier128_relu", "fourier64_relu"])
def test_hessian_on_the_tensor_path(name, shipped_net, torch_cuda):
    """Networks with a ReLU hidden layer: the Hessian comes out of the tensor kernel's second epilogue (no extra GEMM).
    Checked against the fp64 oracle (kink-adjacent points excluded and counted) and against the FP32 K1b kernel."""
    torch = torch_cuda
    from nlotrajectories_b200.sdf import LearnedSDF
    net = shipped_net if name == "shipped_fourier128" else (NETS[name]() if name in NETS else so.synthetic_fourier(64, 1, scale=2.0, seed=14))
    model = LearnedSDF(to_weights(net))
    assert model.precision == "tc3xf16"
    n = 50_021
    P = sample_points(n, seed=6)
    sbar = np.random.default_rng(4).uniform(0.5, 1.5, n).astype(np.float32)
    x, y, sb = (torch.from_numpy(a.copy()).cuda() for a in (P[:, 0], P[:, 1], sbar))
    hxx, hxy, hyy = (t.cpu().numpy() for t in model.hess(x, y, sb))
    H_ref = so.jac_adj1(net.astype(np.float64), P.astype(np.float64), sbar.astype(np.float64))
    tie = kink_mask(net, P)
    assert tie.mean() < 2e-2
    scale = max(1.0, np.abs(H_ref).max())
    for got, ref in ((hxx, H_ref[:, 0, 0]), (hxy, H_ref[:, 0, 1]), (hyy, H_ref[:, 1, 1])):
        assert np.abs(got - ref)[~tie].max() <= 2e-5 * scale, (np.abs(got - ref)[~tie].max(), scale)
    ref32 = LearnedSDF(to_weights(net), precision="fp32")
    gxx, gxy, gyy = (t.cpu().numpy() for t in ref32.hess(x, y, sb))
    assert np.abs(gxx - hxx)[~tie].max() <= 4e-5 * scale and np.abs(gyy - hyy)[~tie].max() <= 4e-5 * scale
    # host entry point takes the same route
    h2 = model.hess_host(P[:1000, 0].copy(), P[:1000, 1].copy(), sbar[:1000])
    assert np.array_equal(h2[0], hxx[:1000]) and np.array_equal(h2[2], hyy[:1000])
    model.close(); ref32.close()


def test_edge_cases_empty_single_and_host_path(shipped_net, torch_cuda):
    torch = torch_cuda
    from nlotrajectories_b200.sdf import LearnedSDF, NNObstacle
    model = LearnedSDF(to_weights(shipped_net))
    e = torch.empty(0, device="cuda")
    s, jx, jy = model.eval(e, e)
    assert s.numel() == 0
    P = sample_points(1, seed=9)
    s1, J1 = run_device(model, torch, P)
    s_ref, J_ref = so.value_jac(shipped_net.astype(np.float64), P.astype(np.float64))
    assert not close(s1, s_ref, TOL).any() and not close(J1, J_ref, TOL).any()
    P = sample_points(777, seed=10)
    sh, jxh, jyh = model.eval_host(P[:, 0].copy(), P[:, 1].copy())
    sd, Jd = run_device(model, torch, P)
    assert np.array_equal(sh, sd) and np.array_equal(jxh, Jd[:, 0]) and np.array_equal(jyh, Jd[:, 1])
    grid_x, grid_y = np.meshgrid(np.linspace(-0.5, 1.5, 31, dtype=np.float32), np.linspace(-0.5, 1.5, 17, dtype=np.float32))
    out = NNObstacle(None, model).approximated_sdf(grid_x, grid_y)          # core/sdf/l4casadi.py:242-246
    assert out.shape == grid_x.shape
    with pytest.raises(TypeError):
        NNObstacle(None, model).approximated_sdf(1.0, 2.0)
    model.close()


@pytest.mark.parametrize("name", ["relu256", "relu64", "fourier256_relu"])
@pytest.mark.parametrize("n", [1, 127, 129, 128 * 148 + 5, 128 * 148 * 4 + 77])
def test_tensor_paths_ragged_sizes(name, n, torch_cuda):
    """Tile / group / ring bookkeeping of the tensor kernels at sizes around one tile, one tile per SM and a few per SM
    (static striding for H = 256, counter-scheduled groups for H = 64)."""
    from nlotrajectories_b200.sdf import LearnedSDF
    net = NETS[name]()
    model = LearnedSDF(to_weights(net))
    assert model.precision == "tc3xf16"
    P = sample_points(n, seed=100 + n % 97)
    s_ref, J_ref = so.value_jac(net.astype(np.float64), P.astype(np.float64))
    for _ in range(2):                                   # twice: barrier phases / counters must be reusable
        s, J = run_device(model, torch_cuda, P)
        tie = kink_mask(net, P)
        assert not close(s, s_ref, TOL).any()
        assert not (close(J, J_ref, TOL).any(axis=1) & ~tie).any()
    s_only, _ = run_device(model, torch_cuda, P, want_jac=False)
    assert np.array_equal(s_only, s)
    model.close()


@pytest.mark.parametrize("name", ["relu128", "relu64", "fourier128_relu"])
@pytest.mark.parametrize("tiles_per_group", [1, 2, 3, 5])
def test_counter_scheduled_tiles_two_ahead(name, tiles_per_group, torch_cuda):
    """The tile counter of the H = 64 / 128 kernels runs two tiles ahead of the one being computed (the first two tiles of a group
    come from its position, the rest from the counter; the next tile's coordinates are prefetched): sizes around 1, 2, 3 and 5
    tiles per group, one point short of and one point past the boundary, must be covered exactly once."""
    from nlotrajectories_b200.sdf import LearnedSDF
    net = NETS[name]()
    model = LearnedSDF(to_weights(net))
    assert model.precision == "tc3xf16"
    groups = 148 * (4 if name == "relu64" else 2)
    for n in (128 * groups * tiles_per_group - 1, 128 * groups * tiles_per_group + 1):
        P = sample_points(n, seed=7 + tiles_per_group)
        s_ref, J_ref = so.value_jac(net.astype(np.float64), P.astype(np.float64))
        s, J = run_device(model, torch_cuda, P)
        tie = kink_mask(net, P)
        assert not close(s, s_ref, TOL).any()
        assert not (close(J, J_ref, TOL).any(axis=1) & ~tie).any()
    model.close()


def test_linearity_of_adjoint_and_full_size_property(torch_cuda):
    """Size-independent properties at sweep size (2^22 points): adj1(sbar) == sbar * jac, and the value
    from a value-only launch equals the value from a value+Jacobian launch bit for bit."""
    torch = torch_cuda
    from nlotrajectories_b200.sdf import LearnedSDF
    net = so.synthetic_mlp(128, 1, seed=0)
    model = LearnedSDF(to_weights(net))
    n = 1 << 22
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand(n, device="cuda", generator=g) * 2 - 0.5
    y = torch.rand(n, device="cuda", generator=g) * 2 - 0.5
    sb = torch.rand(n, device="cuda", generator=g) + 0.5
    s, jx, jy = model.eval(x, y)
    s2, ax, ay = model.eval(x, y, sb)
    assert torch.equal(s, s2)
    assert torch.allclose(ax, sb * jx, rtol=2e-5, atol=1e-6) and torch.allclose(ay, sb * jy, rtol=2e-5, atol=1e-6)
    idx = torch.randint(0, n, (2000,), device="cuda", generator=g)
    P = torch.stack([x[idx], y[idx]], 1).cpu().numpy()
    s_ref, J_ref = so.value_jac(net.astype(np.float64), P.astype(np.float64))
    tie = kink_mask(net, P)
    assert not close(s[idx].cpu().numpy(), s_ref, TOL).any()
    J = torch.stack([jx[idx], jy[idx]], 1).cpu().numpy()
    assert not (close(J, J_ref, TOL).any(axis=1) & ~tie).any()
    model.close()


def _dptr(arr):
    return arr.ctypes.data_as(C.POINTER(C.c_double))


def test_casadi_external_abi_calls(shipped_net, library, torch_cuda):
    """Call nn_sdf / jac_nn_sdf / adj1_nn_sdf / jac_adj1_nn_sdf exactly as CasADi's external loader would
    (arg/res/iw/w/mem, doubles, column-major) and compare with the reference TorchScript goldens."""
    from nlotrajectories_b200.sdf import LearnedSDF
    L = library
    z = np.load(GOLDEN / "sdf_shipped_fourier128.npz")
    model = LearnedSDF(to_weights(shipped_net))
    model.bind_casadi(batch=1)
    DP = C.POINTER(C.c_double)
    for i in range(6):
        p = np.array(z["P"][i], np.float64); seed = np.array([z["sbar"][i]], np.float64)
        out1 = np.zeros(1); out2 = np.zeros(2); out4 = np.zeros(4)
        assert L.nn_sdf((DP * 1)(_dptr(p)), (DP * 1)(_dptr(out1)), None, None, 0) == 0
        assert L.jac_nn_sdf((DP * 2)(_dptr(p), None), (DP * 1)(_dptr(out2)), None, None, 0) == 0
        assert abs(out1[0] - z["value"][i]) <= TOL and np.abs(out2 - z["jac"][i]).max() <= TOL
        assert L.adj1_nn_sdf((DP * 3)(_dptr(p), None, _dptr(seed)), (DP * 1)(_dptr(out2)), None, None, 0) == 0
        assert np.abs(out2 - z["adj1"][i]).max() <= TOL * max(1, np.abs(z["adj1"][i]).max())
        assert L.jac_adj1_nn_sdf((DP * 4)(_dptr(p), None, _dptr(seed), None), (DP * 3)(_dptr(out4), None, None), None, None, 0) == 0
        assert np.abs(out4.reshape(2, 2) - z["jac_adj1"][i]).max() <= TOL * max(1, np.abs(z["jac_adj1"]).max())
    # NULL seed == zeros; NULL result == not requested; unsupported outputs -> non-zero (nn_sdf.cpp:93-101)
    out2 = np.ones(2); p = np.array(z["P"][0], np.float64)
    assert L.adj1_nn_sdf((DP * 3)(_dptr(p), None, None), (DP * 1)(_dptr(out2)), None, None, 0) == 0 and not out2.any()
    assert L.nn_sdf((DP * 1)(_dptr(p)), (DP * 1)(None), None, None, 0) == 0
    junk = np.zeros(4)
    assert L.jac_adj1_nn_sdf((DP * 4)(_dptr(p), None, None, None), (DP * 3)(_dptr(junk), _dptr(junk), None), None, None, 0) != 0
    assert L.jac_adj1_nn_sdf((DP * 4)(_dptr(p), None, None, None), (DP * 3)(None, None, None), None, None, 0) != 0
    # batched externals
    Pn = 64
    model.bind_casadi(batch=Pn)
    pts = np.asfortranarray(z["P"][:Pn].astype(np.float64))          # column-major P x 2: x[P] then y[P]
    flat = pts.ravel(order="F").copy(); seeds = z["sbar"][:Pn].astype(np.float64)
    o1 = np.zeros(Pn); o2 = np.zeros(2 * Pn); o4 = np.zeros(4 * Pn)
    assert L.nn_sdf_batch((DP * 1)(_dptr(flat)), (DP * 1)(_dptr(o1)), None, None, 0) == 0
    assert not close(o1, z["value"][:Pn], TOL).any()
    assert L.jac_nn_sdf_batch((DP * 2)(_dptr(flat), None), (DP * 1)(_dptr(o2)), None, None, 0) == 0
    assert not close(o2.reshape(2, Pn).T, z["jac"][:Pn], TOL).any()
    assert L.adj1_nn_sdf_batch((DP * 3)(_dptr(flat), None, _dptr(seeds)), (DP * 1)(_dptr(o2)), None, None, 0) == 0
    assert not close(o2.reshape(2, Pn).T, z["adj1"][:Pn], TOL).any()
    assert L.jac_adj1_nn_sdf_batch((DP * 4)(_dptr(flat), None, _dptr(seeds), None), (DP * 3)(_dptr(o4), None, None), None, None, 0) == 0
    Hs = z["jac_adj1"][:Pn]
    scale = TOL * max(1, np.abs(Hs).max())
    assert np.abs(o4[0:2 * Pn:2] - Hs[:, 0, 0]).max() <= scale and np.abs(o4[1:2 * Pn:2] - Hs[:, 0, 1]).max() <= scale
    assert np.abs(o4[2 * Pn::2] - Hs[:, 1, 0]).max() <= scale and np.abs(o4[2 * Pn + 1::2] - Hs[:, 1, 1]).max() <= scale
    L.nlo_casadi_bind(None)
    model.close()


def test_two_tensor_path_models_alternating_and_streams(torch_cuda):
    """The tensor path keeps one model's small vectors in __constant__ memory: alternating between two models
    (and launching on side streams) must re-upload them correctly."""
    torch = torch_cuda
    from nlotrajectories_b200.sdf import LearnedSDF
    nets = [so.synthetic_mlp(128, 1, seed=21), so.synthetic_mlp(128, 1, seed=22), so.synthetic_mlp(64, 1, seed=23)]
    models = [LearnedSDF(to_weights(n)) for n in nets]
    assert all(m.precision == "tc3xf16" for m in models)
    P = sample_points(5000, seed=3)
    refs = [so.value_jac(n.astype(np.float64), P.astype(np.float64)) for n in nets]
    ties = [kink_mask(n, P) for n in nets]
    x = torch.from_numpy(P[:, 0].copy()).cuda(); y = torch.from_numpy(P[:, 1].copy()).cuda()
    side = torch.cuda.Stream()
    for rep in range(3):
        for i in (0, 1, 2, 1, 0):
            stream = side if (rep + i) % 2 else torch.cuda.current_stream()
            with torch.cuda.stream(stream):
                s, jx, jy = models[i].eval(x, y)
            stream.synchronize()
            J = np.stack([jx.cpu().numpy(), jy.cpu().numpy()], 1)
            assert not close(s.cpu().numpy(), refs[i][0], TOL).any()
            assert not (close(J, refs[i][1], TOL).any(axis=1) & ~ties[i]).any()
    for m in models:
        m.close()


def test_two_host_threads_share_the_tensor_path(torch_cuda):
    """Two host threads, each evaluating its own model on its own stream at the same time: the check of the
    __constant__ owner, the re-upload and the launch are one critical section, and the tile counters are handed
    out atomically, so every result must still match the oracle."""
    import threading
    torch = torch_cuda
    from nlotrajectories_b200.sdf import LearnedSDF
    nets = [so.synthetic_mlp(128, 1, seed=31), so.synthetic_mlp(128, 1, seed=32)]
    models = [LearnedSDF(to_weights(n)) for n in nets]
    P = sample_points(40000, seed=5)
    refs = [so.value_jac(n.astype(np.float64), P.astype(np.float64))[0] for n in nets]
    x = torch.from_numpy(P[:, 0].copy()).cuda(); y = torch.from_numpy(P[:, 1].copy()).cuda()
    torch.cuda.synchronize()
    bad = [0, 0]

    def work(i):
        torch.cuda.set_device(0)
        st = torch.cuda.Stream()
        for _ in range(40):
            with torch.cuda.stream(st):
                s, _, _ = models[i].eval(x, y)
            st.synchronize()
            bad[i] += int(close(s.cpu().numpy(), refs[i], TOL).sum())

    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert bad == [0, 0]
    for m in models:
        m.close()
