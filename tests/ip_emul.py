"""CPU emulation of the device interior-point solver for the tests (test infrastructure only).

tests/tools/ip_host_emul.cpp compiles the SAME per-problem bodies, table builders and iteration loop that the CUDA kernels of
nlotrajectories_b200/csrc/ip_solver.cu wrap into a small host library; the NLP evaluation is supplied by the numpy oracle
through ctypes callbacks.  This pins the solver logic and the block-tridiagonal linear algebra without a GPU."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from oracle import nlp_oracle as no
from oracle import sdf_oracle as so

HERE = Path(__file__).resolve().parent
SRC = HERE / "tools" / "ip_host_emul.cpp"
OUT = HERE / "_build" / "libip_emul.so"
CSRC = HERE.parent / "nlotrajectories_b200" / "csrc"

EVAL_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p)
HESS_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p)


def build() -> C.CDLL:
    deps = [SRC, CSRC / "ip_core.cuh", CSRC / "ip_tables.hpp"]
    if not OUT.exists() or OUT.stat().st_mtime < max(d.stat().st_mtime for d in deps):
        OUT.parent.mkdir(exist_ok=True)
        r = subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-o", str(OUT), str(SRC)],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"g++ failed on {SRC}:\n{r.stderr}")
    lib = C.CDLL(str(OUT))
    lib.ip_emul_last_error.restype = C.c_char_p
    return lib


def _view(ptr, shape, dtype=np.float32):
    n = int(np.prod(shape))
    ctype = C.c_float if dtype == np.float32 else C.c_double
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(n,)).reshape(shape)


class OracleProblem:
    """Sizes, patterns, bounds and fp32 variable-major evaluation callbacks of one NLP family on the numpy oracle."""

    def __init__(self, spec, net=None):
        self.spec = spec
        self.n_w, self.n_g = spec.n_w, spec.n_g
        rows, cols, _ = no.jac_pattern(spec)
        assert np.all(np.diff(cols) >= 0), "oracle pattern must be in compressed-column order"
        self.jrow = rows.astype(np.int32)
        self.jcolind = np.concatenate([[0], np.cumsum(np.bincount(cols, minlength=spec.n_w))]).astype(np.int32)
        hrows, hcols = no.hess_pattern(spec)
        assert np.all(np.diff(hcols) >= 0)
        self.hrow = hrows.astype(np.int32)
        self.hcolind = np.concatenate([[0], np.cumsum(np.bincount(hcols, minlength=spec.n_w))]).astype(np.int32)
        self.nnz, self.nnzh = len(self.jrow), len(self.hrow)
        self.lb, self.ub = (np.ascontiguousarray(b, np.float64) for b in no.bounds(spec))
        if net is not None:
            n64 = net.astype(np.float64)
            self.sdf = lambda Q: so.value_jac(n64, Q)
            self.sdf_h = lambda Q: (lambda H: np.stack([H[:, 0, 0], H[:, 0, 1], H[:, 1, 1]], -1))(so.jac_adj1(n64, Q, np.ones(len(Q))))
        else:
            self.sdf = self.sdf_h = None
        self.evals = 0

    @property
    def stages(self):
        s = self.spec
        return dict(N=s.N, nx=s.nx, nu=s.nu, use_slack=int(s.use_slack), n_term=len(s.terminal_idx), g_off_dyn=s.nx + len(s.terminal_idx))

    def eval_cb(self):
        def cb(w32, P, ld, g, jac, f, grad):
            try:
                w = _view(w32, (self.n_w, ld))[:, :P].T.astype(np.float64)
                self.evals += 1
                if g or jac:
                    gv, jv = no.eval_g_jac(self.spec, w, self.sdf)
                    if g:
                        _view(g, (self.n_g, ld))[:, :P] = gv.T
                    if jac:
                        _view(jac, (self.nnz, ld))[:, :P] = jv.T
                if f or grad:
                    fv, gr = no.eval_f_grad(self.spec, w)
                    if f:
                        _view(f, (ld,))[:P] = fv
                    if grad:
                        _view(grad, (self.n_w, ld))[:, :P] = gr.T
                return 0
            except Exception as exc:          # noqa: BLE001 - a Python exception must not cross the C boundary
                print("eval callback failed:", exc)
                return 1
        return EVAL_CB(cb)

    def hess_cb(self):
        def cb(w32, lam32, P, ld, hess):
            try:
                w = _view(w32, (self.n_w, ld))[:, :P].T.astype(np.float64)
                lam = _view(lam32, (self.n_g, ld))[:, :P].T.astype(np.float64)
                hv = no.eval_hess_lag(self.spec, w, np.ones(P), lam, self.sdf, self.sdf_h)
                _view(hess, (self.nnzh, ld))[:, :P] = hv.T
                return 0
            except Exception as exc:          # noqa: BLE001
                print("hess callback failed:", exc)
                return 1
        return HESS_CB(cb)


def _i32(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _f64(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def solve(op: OracleProblem, w0, tol=1e-4, max_iter=300, mu0=0.1, ls_multipliers=True, compact=True, verbose=0):
    lib = build()
    w0 = np.ascontiguousarray(w0, np.float64)
    P = w0.shape[0]
    w = np.empty_like(w0); lam = np.empty((P, op.n_g)); f = np.empty(P); viol = np.empty(P); err = np.empty(P)
    iters = np.empty(P, np.int32); status = np.empty(P, np.int32); stats = np.zeros(5, np.int32)
    ecb, hcb = op.eval_cb(), op.hess_cb()
    st = op.stages
    rc = lib.ip_emul_solve(C.c_int(op.n_w), C.c_int(op.n_g), _i32(op.jcolind), _i32(op.jrow), C.c_int(op.nnzh), _i32(op.hcolind), _i32(op.hrow),
                           _f64(op.lb), _f64(op.ub), C.c_int(st["N"]), C.c_int(st["nx"]), C.c_int(st["nu"]), C.c_int(st["use_slack"]),
                           C.c_int(st["n_term"]), C.c_int(st["g_off_dyn"]), _f64(w0), C.c_size_t(P), C.c_double(tol), C.c_int(max_iter),
                           C.c_double(mu0), C.c_int(int(ls_multipliers)), C.c_int(int(compact)), C.c_int(verbose), ecb, hcb, _f64(w), _f64(f),
                           _f64(viol), _f64(err), _i32(iters), _i32(status), _f64(lam), _i32(stats))
    if rc:
        raise RuntimeError("ip_emul_solve failed: " + lib.ip_emul_last_error().decode())
    return dict(w=w, f=f, violation=viol, kkt_error=err, iterations=iters, status=status, lam=lam,
                stats=dict(zip(("iterations", "evaluations", "hessians", "trials", "compactions"), map(int, stats))))


def kkt_step(op: OracleProblem, jac, hess, omega, rhs, delta_in):
    """(H + J^T diag(omega) J + delta I) dw = rhs per problem through the block-tridiagonal path.  Inputs problem-major:
    jac (P, nnz), hess (P, nnzh), omega (P, n_g), rhs (P, n_w), delta_in (P,).  Returns dw (P, n_w), delta_out (P,)."""
    lib = build()
    P = jac.shape[0]
    jv = np.ascontiguousarray(jac.T, np.float32); hv = np.ascontiguousarray(hess.T, np.float32)
    om = np.ascontiguousarray(omega.T, np.float64); rh = np.ascontiguousarray(rhs.T, np.float64)
    d_in = np.ascontiguousarray(delta_in, np.float64)
    dw = np.empty((op.n_w, P)); d_out = np.empty(P)
    st = op.stages
    rc = lib.ip_emul_kkt_step(C.c_int(op.n_w), C.c_int(op.n_g), _i32(op.jcolind), _i32(op.jrow), C.c_int(op.nnzh), _i32(op.hcolind), _i32(op.hrow),
                              _f64(op.lb), _f64(op.ub), C.c_int(st["N"]), C.c_int(st["nx"]), C.c_int(st["nu"]), C.c_int(st["use_slack"]),
                              C.c_int(st["n_term"]), C.c_int(st["g_off_dyn"]), C.c_size_t(P), jv.ctypes.data_as(C.POINTER(C.c_float)),
                              hv.ctypes.data_as(C.POINTER(C.c_float)), _f64(om), _f64(rh), _f64(d_in), _f64(dw), _f64(d_out))
    if rc:
        raise RuntimeError("ip_emul_kkt_step failed: " + lib.ip_emul_last_error().decode())
    return dw.T.copy(), d_out
