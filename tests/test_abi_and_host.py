"""CPU: the C-ABI library loads and exports every symbol include/nlo_b200.h declares (no compute calls),
the CasADi-ABI metadata matches _l4c_generated/nn_sdf.cpp, and the host-side logic (YAML schema, weight
import, sharding, best-of selection over gloo) behaves like the reference's."""
import ctypes as C
import os
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import yaml

from conftest import BENCH, GOLDEN, REPO, bench_yaml


def header_symbols():
    text = (REPO / "include" / "nlo_b200.h").read_text()
    return sorted(set(re.findall(r"NLO_API\s+[\w\s\*]+?\b(\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol(library):
    from nlotrajectories_b200 import lib
    names = header_symbols()
    assert len(names) > 60
    out = subprocess.run(["nm", "-D", "--defined-only", str(lib.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if ln.strip()}
    missing = [n for n in names if n not in exported]
    assert not missing, missing
    assert set(names) == set(lib.SIGNATURES), set(names) ^ set(lib.SIGNATURES)


def test_casadi_abi_metadata_matches_reference_shim(library):
    """Counts and sparsities of _l4c_generated/nn_sdf.cpp:36-55, 64-65, 76-77, 88-89."""
    L = library
    assert (L.nn_sdf_n_in(), L.nn_sdf_n_out()) == (1, 1)
    assert (L.jac_nn_sdf_n_in(), L.jac_nn_sdf_n_out()) == (2, 1)
    assert (L.adj1_nn_sdf_n_in(), L.adj1_nn_sdf_n_out()) == (3, 1)
    assert (L.jac_adj1_nn_sdf_n_in(), L.jac_adj1_nn_sdf_n_out()) == (4, 3)
    assert list(L.nn_sdf_sparsity_in(0)[:3]) == [1, 2, 1]
    assert list(L.nn_sdf_sparsity_out(0)[:3]) == [1, 1, 1]
    assert not L.nn_sdf_sparsity_in(1) and not L.nn_sdf_sparsity_out(1)


def test_batched_casadi_sparsity(library):
    L = library
    assert L.nlo_casadi_set_batch(5) == 0
    assert list(L.nn_sdf_batch_sparsity_in(0)[:3]) == [5, 2, 1]
    assert list(L.nn_sdf_batch_sparsity_out(0)[:3]) == [5, 1, 1]
    sp = L.jac_nn_sdf_batch_sparsity_out(0)
    nrow, ncol = sp[0], sp[1]
    colind = [sp[2 + i] for i in range(ncol + 1)]
    rows = [sp[2 + ncol + 1 + i] for i in range(colind[-1])]
    assert (nrow, ncol) == (5, 10) and colind == list(range(11)) and rows == [0, 1, 2, 3, 4] * 2
    sp = L.jac_adj1_nn_sdf_batch_sparsity_out(0)
    assert (sp[0], sp[1]) == (10, 10)
    colind = [sp[2 + i] for i in range(11)]
    rows = [sp[13 + i] for i in range(20)]
    assert colind == list(range(0, 21, 2)) and rows[:4] == [0, 5, 1, 6] and rows[10:14] == [0, 5, 1, 6]
    assert L.nlo_casadi_set_batch(0) != 0


def test_no_gpu_means_loud_failure_not_fallback(library):
    """Without a CUDA device model creation must fail with a message (never a silent CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from nlotrajectories_b200 import lib
    from nlotrajectories_b200.sdf import LearnedSDF, SdfWeights
    w = SdfWeights.from_npz(GOLDEN / "sdf_shipped_fourier128_weights.npz")
    with pytest.raises(lib.NloError, match="no CUDA device|no usable CUDA"):
        LearnedSDF(w)
    res = (C.POINTER(C.c_double) * 1)((C.c_double * 1)())
    arg = (C.POINTER(C.c_double) * 1)((C.c_double * 2)(0.1, 0.2))
    os.environ.pop("NLO_B200_WEIGHTS", None)
    assert library.nn_sdf(arg, res, None, None, 0) != 0


def test_weight_blob_roundtrip_and_torchscript_import(tmp_path, library):
    from nlotrajectories_b200.sdf import ACT_COS_SCALE, ACT_RELU, SdfWeights
    w = SdfWeights.from_npz(GOLDEN / "sdf_shipped_fourier128_weights.npz")
    assert (w.kind, w.hidden, w.n_hidden_mats, w.act0, w.act, w.p0) == ("fourier", 128, 1, ACT_COS_SCALE, ACT_RELU, 10.0)
    assert w.blob.size == 3 * 128 + 128 * 128 + 128 + 128 + 1
    p = tmp_path / "m.nlow"
    w.save_nlow(p)
    w2 = SdfWeights.from_nlow(p)
    assert (w2.kind, w2.hidden, w2.n_hidden_mats, w2.act0, w2.act, w2.p0, w2.p) == (w.kind, 128, 1, w.act0, w.act, 10.0, 1.0)
    assert np.array_equal(w.blob, w2.blob)
    ref_pt = REPO / "oracle" / "_ref" / "nn_sdf.pt"
    if ref_pt.exists():
        w3 = SdfWeights.from_torchscript(ref_pt)
        assert np.array_equal(w3.blob, w.blob) and w3.p0 == 10.0


def test_state_dict_import_matches_torch_modules():
    """FourierMLP / SIREN layouts of core/nn_architectures.py and l4casadi's naive MLP naming."""
    import torch
    from nlotrajectories_b200.sdf import SdfWeights
    from oracle import sdf_oracle as so
    torch.manual_seed(0)
    H = 16
    sd = {"input_layer.weight": torch.randn(H, 2), "input_layer.bias": torch.randn(H),
          "hidden_layers.0.weight": torch.randn(H, H), "hidden_layers.0.bias": torch.randn(H),
          "output_layer.weight": torch.randn(1, H), "output_layer.bias": torch.randn(1)}
    w = SdfWeights.from_state_dict("mlp", sd, activation_function="ReLU")
    P = torch.rand(7, 2)
    h = torch.relu(P @ sd["input_layer.weight"].T + sd["input_layer.bias"])
    h = torch.relu(h @ sd["hidden_layers.0.weight"].T + sd["hidden_layers.0.bias"])
    want = (h @ sd["output_layer.weight"].T + sd["output_layer.bias"])[:, 0].numpy()
    b = w.blob
    net = so.SdfNet("mlp", b[:2 * H].reshape(H, 2), b[2 * H:3 * H], [(b[3 * H:3 * H + H * H].reshape(H, H), b[3 * H + H * H:4 * H + H * H])],
                    b[4 * H + H * H:5 * H + H * H], float(b[-1]), w.act0, w.act, w.p0, w.p)
    np.testing.assert_allclose(so.forward(net, P.numpy()), want, atol=1e-5)
    sdf = {"fourier.weights": torch.randn(2, H), "fourier.bias": torch.randn(H), "layers.0.weight": torch.randn(H, H),
           "layers.0.bias": torch.randn(H), "output_layer.weight": torch.randn(1, H), "output_layer.bias": torch.randn(1)}
    wf = SdfWeights.from_state_dict("fourier", sdf, activation_function="tanh", scale=1.0)
    assert (wf.act0, wf.act, wf.n_hidden_mats) == (so.ACT_COS_SCALE, so.ACT_TANH, 1)
    np.testing.assert_array_equal(wf.blob[:2 * H].reshape(H, 2), sdf["fourier.weights"].T.numpy())
    with pytest.raises(ValueError, match="Unsupported activation function"):
        SdfWeights.from_state_dict("mlp", sd, activation_function="swish")
    with pytest.raises(ValueError, match="Unsupported model type"):
        SdfWeights.from_state_dict("cnn", sd)


def test_config_parses_reference_yamls_unchanged():
    from nlotrajectories_b200.config import Config, ConfigError
    files = sorted(BENCH.glob("benchmark_*.yaml"))
    assert len(files) == 6
    for f in files:
        cfg = Config.load(f)
        assert cfg.solver.type == "ipopt" and cfg.model.type == "mlp" and cfg.model.hidden_dim == 128
        assert cfg.model.n_hidden_mats() == 1
    c6 = Config.load(bench_yaml("benchmark_6"))
    assert (c6.body.dynamic, c6.body.shape, c6.solver.N, c6.solver.use_slack, c6.solver.use_smooth) == ("ackermann_2nd", "rectangle", 80, False, True)
    assert c6.solver.initializer.mode == "rrt" and c6.solver.initializer.max_iter == 5000
    c1 = Config.load(bench_yaml("benchmark_1"))
    assert c1.circles() == [(0.5, 0.5, 0.2, 0.05, 0)]
    c5 = Config.load(bench_yaml("benchmark_5"))          # circle + three squares, analytic mode
    assert [o[4] for o in c5.circles()] == [0, 1, 1, 1] and c5.circles()[2][:4] == (1.1, 0.7, 0.2, 0.01)
    bad = yaml.safe_load(open(bench_yaml("benchmark_1")))
    bad["body"]["control_bounds"] = [-1.0, 1.0]          # the reference's configs/broken.yaml:7 shape
    with pytest.raises(ConfigError):
        Config.parse(bad)
    bad = yaml.safe_load(open(bench_yaml("benchmark_1"))); bad["solver"].pop("type")
    with pytest.raises(ConfigError, match="field required: type"):
        Config.parse(bad)
    bad = yaml.safe_load(open(bench_yaml("benchmark_1"))); bad["body"]["dynamic"] = "hovercraft"
    with pytest.raises(ConfigError):
        Config.parse(bad)


def test_shard_range_covers_batch_exactly():
    from nlotrajectories_b200.distributed import shard_range
    for total in (0, 1, 7, 4096, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from nlotrajectories_b200.distributed import init_process_group, select_best, shard_range
rank, local_rank, world = init_process_group("gloo")
total, n_w = 10, 6
lo, hi = shard_range(total, rank, world)
g = torch.Generator().manual_seed(5)
merit_all = torch.rand(total, generator=g)
merit_all[7] = -1.0                      # the winner lives on the last rank
w_all = torch.arange(total * n_w, dtype=torch.float32).reshape(total, n_w)
w_soa = w_all[lo:hi].T.contiguous()
val, idx, w_best = select_best(merit_all[lo:hi].clone(), w_soa, lo, n_w)
assert idx == 7 and abs(val + 1.0) < 1e-12, (val, idx)
assert torch.equal(w_best, w_all[7]), w_best
# an empty shard must not break the exchange
val, idx, _ = select_best(merit_all[lo:hi][:0].clone() if rank == 0 else merit_all[lo:hi].clone(), w_soa, lo, n_w)
assert idx == 7
dist.barrier()
print("ok", rank)
"""


def test_best_of_selection_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29731", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script), str(REPO)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)
