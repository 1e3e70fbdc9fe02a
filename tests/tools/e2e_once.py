"""Host-buffer (e2e) timing of one configuration of nlo_nlp_eval_host[_compact]: benchmark_6 x 65,536, pinned buffers.
The chunking of the host path is read from NLO_B200_HOST_CHUNKS / NLO_B200_HOST_LANES (once per process), so a sweep runs this
script once per setting:   NLO_B200_HOST_CHUNKS=32 NLO_B200_HOST_LANES=3 python tests/tools/e2e_once.py"""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
from gpu_util import to_weights                                  # noqa: E402
from nlotrajectories_b200.config import Config                   # noqa: E402
from nlotrajectories_b200.problem import NlpProblem              # noqa: E402
from nlotrajectories_b200.sdf import LearnedSDF                  # noqa: E402
from oracle import sdf_oracle as so                              # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
model = LearnedSDF(to_weights(so.synthetic_mlp(128, 1, seed=0)))
prob = NlpProblem.from_config(Config.load(next((REPO / "nlotrajectories_b200/benchmarks").glob("benchmark_6*.yaml"))), model)
pin = lambda shape: torch.empty(shape, dtype=torch.float32).pin_memory().numpy()
w = pin((P, prob.n_w)); w[:] = prob.multistart_guess(P)
lay = prob.compact_layout()
comp = {"g": pin((P, len(lay["g_var_rows"]))), "jac": pin((P, len(lay["jac_var_nz"]))), "f": pin((P,)), "grad_f": pin((P, len(lay["grad_var_idx"])))}
full = {"g": pin((P, prob.n_g)), "jac": pin((P, prob.nnz)), "f": pin((P,)), "grad_f": pin((P, prob.n_w))}
out = {}
for name, call in (("compact", lambda: prob.eval_host_compact(w, out=comp)), ("full", lambda: prob.eval_host(w, out=full))):
    call(); call()
    ts = []
    for _ in range(7):
        t0 = time.perf_counter(); call(); ts.append(time.perf_counter() - t0)
    out[name + "_ms_min"] = 1e3 * min(ts); out[name + "_ms_med"] = 1e3 * float(np.median(ts))
# raw copies for scale: one big device->host and host->device copy of the compact payload
nbytes = sum(v.nbytes for v in comp.values())
d = torch.empty(nbytes // 4, dtype=torch.float32, device="cuda"); h = torch.empty(nbytes // 4, dtype=torch.float32).pin_memory()
for tag, fn in (("d2h", lambda: h.copy_(d, non_blocking=True)), ("h2d", lambda: d.copy_(h, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    out[f"raw_{tag}_gbs"] = nbytes / dt / 1e9
out.update(chunks=os.environ.get("NLO_B200_HOST_CHUNKS", "16"), lanes=os.environ.get("NLO_B200_HOST_LANES", "2"), d2h_bytes_compact=nbytes)
print(json.dumps(out), flush=True)
