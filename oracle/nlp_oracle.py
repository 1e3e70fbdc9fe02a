"""CPU (numpy, fp64) restatement of the reference's NLP evaluation (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows, in the reference tree (src/nlotrajectories/...):
* ``core/runner.py:44-103``      - decision-vector order, constraint order, objective
* ``core/dynamics.py:33-148``    - the six continuous-time models f(x,u)
* ``core/geometry.py:59-144``    - footprint transform, per-knot SDF rows
* ``core/utils.py:18-33``        - soft_min (un-stabilised log-sum-exp, alpha = 10)
* ``core/sdf/casadi.py:27-45, 377-390`` - analytic circle, MultiObstacle soft-min union
* ``core/sdf/l4casadi.py:241-257``      - learned SDF call site

Layout:
  w = [vec(X) (nx*(N+1), knot-major) ; vec(U) (nu*N) ; slack (N+1, only if use_slack)]
  g = [init nx ; terminal (nx or nx-1) ; Euler defects N*nx ; slack>=0 (N+1) ; SDF rows ; control box nu*N]
  dg/dw is returned as values on a fixed structural pattern in compressed-column order.

Everything is vectorised over a batch of problems: ``w`` has shape (P, n_w).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional, Tuple

import numpy as np

ALPHA = 10.0      # core/utils.py:18
EPS_PATH = 1e-8   # core/runner.py:82

DYN_DIMS = {"point_1st": (4, 2), "point_2nd": (4, 2), "unicycle": (3, 2), "unicycle_2nd": (5, 2),
            "ackermann": (4, 2), "ackermann_2nd": (7, 2)}

# structural non-zeros of A = df/dx and B = df/du, row-major order
DYN_A = {
    "point_1st": [],
    "point_2nd": [(0, 2), (1, 3)],
    "unicycle": [(0, 2), (1, 2)],
    "unicycle_2nd": [(0, 2), (0, 3), (1, 2), (1, 3), (2, 4)],
    "ackermann": [(0, 2), (1, 2), (2, 3)],
    "ackermann_2nd": [(0, 2), (0, 4), (1, 2), (1, 4), (2, 3), (2, 4), (3, 6), (4, 3), (4, 4), (4, 6)],
}
DYN_B = {
    "point_1st": [(0, 0), (1, 1)],
    "point_2nd": [(2, 0), (3, 1)],
    "unicycle": [(0, 0), (1, 0), (2, 1)],
    "unicycle_2nd": [(3, 0), (4, 1)],
    "ackermann": [(0, 0), (1, 0), (2, 0), (3, 1)],
    "ackermann_2nd": [(4, 0), (5, 0), (6, 1)],
}


@dataclass
class NlpSpec:
    dynamics: str
    shape: str                       # "dot" | "rectangle" | "triangle"
    x0: np.ndarray
    goal: np.ndarray
    control_bounds: List[Tuple[float, float]]
    N: int
    dt: float
    use_slack: bool = False
    slack_penalty: float = 1000.0
    use_smooth: bool = False
    smooth_weight: float = 10.0
    enforce_heading: bool = True
    length: Optional[float] = None
    width: Optional[float] = None
    wheelbase: Optional[float] = None
    sdf_mode: str = "l4casadi"       # "l4casadi": learned net; "casadi": analytic soft-min union
    circles: List[Tuple[float, float, float, float]] = field(default_factory=list)  # analytic obstacles: (cx, cy, r | size, margin)
    kinds: List[int] = field(default_factory=list)                                  # 0 = circle, 1 = square (same order)

    @staticmethod
    def from_yaml_dict(d: dict) -> "NlpSpec":
        b, s = d["body"], d["solver"]
        circles, kinds = [], []
        if s["mode"] == "casadi":
            for o in d["obstacles"]:
                if o["type"] == "circle":
                    circles.append((float(o["center"][0]), float(o["center"][1]), float(o["radius"]), float(o.get("margin", 0.0))))
                    kinds.append(0)
                elif o["type"] == "square":
                    circles.append((float(o["center"][0]), float(o["center"][1]), float(o["size"]), float(o.get("margin", 0.0))))
                    kinds.append(1)
                else:
                    raise NotImplementedError("oracle covers analytic circles and squares (benchmarks 1, 2, 5)")
        return NlpSpec(
            dynamics=b["dynamic"], shape=b["shape"], x0=np.array(b["start_state"], float),
            goal=np.array(b["goal_state"], float), control_bounds=[tuple(map(float, cb)) for cb in b["control_bounds"]],
            N=int(s.get("N", 20)), dt=float(s.get("dt", 0.1)), use_slack=bool(s.get("use_slack", False)),
            slack_penalty=float(s.get("slack_penalty", 1000)), use_smooth=bool(s.get("use_smooth", False)),
            smooth_weight=float(s.get("smooth_weight", 10.0)), enforce_heading=bool(s.get("enforce_heading", True)),
            length=b.get("length"), width=b.get("width"), wheelbase=b.get("wheelbase"), sdf_mode=s["mode"], circles=circles, kinds=kinds)

    # -- sizes -----------------------------------------------------------------------------------
    @property
    def nx(self): return DYN_DIMS[self.dynamics][0]
    @property
    def nu(self): return DYN_DIMS[self.dynamics][1]
    @property
    def n_X(self): return self.nx * (self.N + 1)
    @property
    def n_U(self): return self.nu * self.N
    @property
    def n_w(self): return self.n_X + self.n_U + ((self.N + 1) if self.use_slack else 0)
    @property
    def terminal_idx(self):
        return list(range(self.nx)) if self.enforce_heading else [i for i in range(self.nx) if i != 2]

    def body_points(self) -> np.ndarray:
        if self.shape == "dot":
            return np.zeros((1, 2))
        l, w = self.length / 2, self.width / 2
        if self.shape == "rectangle":           # core/geometry.py:125-135
            return np.array([(-l, -w), (-l, w), (l, w), (l, -w)])
        if self.shape == "triangle":            # core/geometry.py:138-144
            return np.array([(l, 0.0), (-l, w), (-l, -w)])
        raise ValueError(self.shape)

    @property
    def n_body(self): return len(self.body_points())
    @property
    def sdf_rows_per_knot(self):
        if self.shape == "dot" or self.use_slack:
            return 1
        return self.n_body
    @property
    def n_g(self):
        return (self.nx + len(self.terminal_idx) + self.N * self.nx + ((self.N + 1) if self.use_slack else 0)
                + (self.N + 1) * self.sdf_rows_per_knot + self.nu * self.N)

    # -- variable indices --------------------------------------------------------------------------
    def iX(self, i, k): return k * self.nx + i
    def iU(self, i, k): return self.n_X + k * self.nu + i
    def iS(self, k): return self.n_X + self.n_U + k


# -------------------------------------------------------------------------------------------------
# dynamics  (core/dynamics.py)
# -------------------------------------------------------------------------------------------------
def dynamics_f(spec: NlpSpec, x: np.ndarray, u: np.ndarray):
    """x: (..., nx), u: (..., nu) -> f (..., nx), A values (..., nnzA), B values (..., nnzB) in DYN_A/DYN_B order."""
    name = spec.dynamics
    z = np.zeros(x.shape[:-1])
    one = np.ones(x.shape[:-1])
    if name == "point_1st":
        f = [u[..., 0], u[..., 1], z, z]
        A, B = [], [one, one]
    elif name == "point_2nd":
        f = [x[..., 2], x[..., 3], u[..., 0], u[..., 1]]
        A, B = [one, one], [one, one]
    elif name == "unicycle":
        th, v, om = x[..., 2], u[..., 0], u[..., 1]
        f = [v * np.cos(th), v * np.sin(th), om]
        A = [-v * np.sin(th), v * np.cos(th)]
        B = [np.cos(th), np.sin(th), one]
    elif name == "unicycle_2nd":
        th, v, om = x[..., 2], x[..., 3], x[..., 4]
        f = [v * np.cos(th), v * np.sin(th), om, u[..., 0], u[..., 1]]
        A = [-v * np.sin(th), np.cos(th), v * np.cos(th), np.sin(th), one]
        B = [one, one]
    elif name == "ackermann":
        L = spec.wheelbase
        th, psi, v, pd = x[..., 2], x[..., 3], u[..., 0], u[..., 1]
        t = np.tan(psi)
        f = [v * np.cos(th), v * np.sin(th), v * t / L, pd]
        A = [-v * np.sin(th), v * np.cos(th), v * (1 + t * t) / L]
        B = [np.cos(th), np.sin(th), t / L, one]
    elif name == "ackermann_2nd":
        # NOTE the reference's slot quirk: v = x[4], psi_dot = x[6] and the
        # returned derivative is (dx, dy, dtheta, dpsi, domega, dv, ddpsi)  (core/dynamics.py:131-148)
        L = spec.wheelbase
        th, psi, v, pd = x[..., 2], x[..., 3], x[..., 4], x[..., 6]
        a, al = u[..., 0], u[..., 1]
        t = np.tan(psi)
        q = 1.0 / (1.0 + psi * psi)
        f = [v * np.cos(th), v * np.sin(th), v * t / L, pd, (pd * q * v + t * a) / L, a, al]
        A = [-v * np.sin(th), np.cos(th), v * np.cos(th), np.sin(th),
             v * (1 + t * t) / L, t / L, one,
             (a * (1 + t * t) - 2 * psi * v * pd * q * q) / L, pd * q / L, v * q / L]
        B = [t / L, one, one]
    else:
        raise ValueError(name)
    stack = lambda lst: np.stack(lst, axis=-1) if lst else np.zeros(x.shape[:-1] + (0,))
    return stack(f), stack(A), stack(B)


# -------------------------------------------------------------------------------------------------
# SDF on footprint points
# -------------------------------------------------------------------------------------------------
class _Jet:
    """Value, gradient and Hessian w.r.t. (x, y), element-wise over arrays: enough arithmetic to push the reference's smooth
    square SDF (core/sdf/casadi.py:69-115) through second order."""

    def __init__(self, v, dx, dy, dxx, dxy, dyy):
        self.v, self.dx, self.dy, self.dxx, self.dxy, self.dyy = v, dx, dy, dxx, dxy, dyy

    @staticmethod
    def const(c, like):
        z = np.zeros_like(like)
        return _Jet(z + c, z, z, z, z, z)

    def __add__(self, o):
        o = o if isinstance(o, _Jet) else _Jet.const(o, self.v)
        return _Jet(self.v + o.v, self.dx + o.dx, self.dy + o.dy, self.dxx + o.dxx, self.dxy + o.dxy, self.dyy + o.dyy)

    def __sub__(self, o):
        o = o if isinstance(o, _Jet) else _Jet.const(o, self.v)
        return _Jet(self.v - o.v, self.dx - o.dx, self.dy - o.dy, self.dxx - o.dxx, self.dxy - o.dxy, self.dyy - o.dyy)

    def __mul__(self, o):
        if not isinstance(o, _Jet):
            return _Jet(self.v * o, self.dx * o, self.dy * o, self.dxx * o, self.dxy * o, self.dyy * o)
        return _Jet(self.v * o.v, self.dx * o.v + self.v * o.dx, self.dy * o.v + self.v * o.dy,
                    self.dxx * o.v + 2 * self.dx * o.dx + self.v * o.dxx,
                    self.dxy * o.v + self.dx * o.dy + self.dy * o.dx + self.v * o.dxy,
                    self.dyy * o.v + 2 * self.dy * o.dy + self.v * o.dyy)

    def sqrt(self):
        r = np.sqrt(self.v)
        f1, f2 = 0.5 / r, -0.25 / (r * self.v)
        return _Jet(r, f1 * self.dx, f1 * self.dy, f2 * self.dx * self.dx + f1 * self.dxx,
                    f2 * self.dx * self.dy + f1 * self.dxy, f2 * self.dy * self.dy + f1 * self.dyy)


def _square_jet(x, y, cx, cy, size, margin):
    """SquareObstacle.approximated_sdf (core/sdf/casadi.py:69-115): soft |.|, soft max / min with eps = 1e-6."""
    one, zero = np.ones_like(x), np.zeros_like(x)
    X = _Jet(x - cx, one, zero, zero, zero, zero)
    Y = _Jet(y - cy, zero, one, zero, zero, zero)
    half = size / 2 + margin
    d_x = (X * X + 1e-6).sqrt() - half
    d_y = (Y * Y + 1e-6).sqrt() - half
    smax = lambda a, b: (a + b + ((a - b) * (a - b) + 1e-6).sqrt()) * 0.5
    smin = lambda a, b: (a + b - ((a - b) * (a - b) + 1e-6).sqrt()) * 0.5
    z = _Jet.const(0.0, x)
    d_x_out, d_y_out = smax(d_x, z), smax(d_y, z)
    outside = (d_x_out * d_x_out + d_y_out * d_y_out).sqrt()
    inside = smin(smax(d_x, d_y), z)
    return outside + inside


def _circle_jet(x, y, cx, cy, r, margin):
    dx, dy = x - cx, y - cy
    d = np.sqrt(dx * dx + dy * dy)
    nx_, ny_ = dx / d, dy / d
    return _Jet(d - (r + margin), nx_, ny_, (1 - nx_ * nx_) / d, -nx_ * ny_ / d, (1 - ny_ * ny_) / d)


def _analytic_union(spec: NlpSpec, pts: np.ndarray):
    """MultiObstacle.approximated_sdf (core/sdf/casadi.py:385-386): soft_min over the obstacles' smooth SDFs, to second order."""
    x, y = pts[:, 0], pts[:, 1]
    kinds = spec.kinds or [0] * len(spec.circles)
    jets = [(_circle_jet if k == 0 else _square_jet)(x, y, *c) for k, c in zip(kinds, spec.circles)]
    vals = np.stack([j.v for j in jets], axis=0)
    e = np.exp(-ALPHA * vals)
    ssum = e.sum(axis=0)
    om = e / ssum
    s = -1.0 / ALPHA * np.log(ssum)                   # soft_min; identity up to rounding for 1 obstacle
    mx = sum(om[i] * jets[i].dx for i in range(len(jets))); my = sum(om[i] * jets[i].dy for i in range(len(jets)))
    hxx = sum(om[i] * (jets[i].dxx - ALPHA * jets[i].dx * jets[i].dx) for i in range(len(jets))) + ALPHA * mx * mx
    hxy = sum(om[i] * (jets[i].dxy - ALPHA * jets[i].dx * jets[i].dy) for i in range(len(jets))) + ALPHA * mx * my
    hyy = sum(om[i] * (jets[i].dyy - ALPHA * jets[i].dy * jets[i].dy) for i in range(len(jets))) + ALPHA * my * my
    return s, np.stack([mx, my], axis=-1), np.stack([hxx, hxy, hyy], axis=-1)


def circles_sdf(spec: NlpSpec, pts: np.ndarray):
    """Analytic mode (solver.mode casadi): value and gradient of the soft-min union (circles: core/sdf/casadi.py:33-41; squares:
    :69-115)."""
    s, g, _ = _analytic_union(spec, pts)
    return s, g


def _footprint(spec: NlpSpec, X: np.ndarray):
    """X: (P, N+1, nx) -> world points (P, N+1, nb, 2) and d(point)/d(theta) (same shape)."""
    bp = spec.body_points()
    x, y = X[..., 0:1], X[..., 1:2]
    if spec.shape == "dot":
        pts = np.stack([x, y], axis=-1)
        return pts, np.zeros_like(pts)
    th = X[..., 2:3]
    c, s = np.cos(th), np.sin(th)
    bx, by = bp[:, 0], bp[:, 1]
    px = x + c * bx - s * by                          # core/geometry.py:80-82
    py = y + s * bx + c * by
    dpx = -s * bx - c * by
    dpy = c * bx - s * by
    return np.stack([px, py], axis=-1), np.stack([dpx, dpy], axis=-1)


# -------------------------------------------------------------------------------------------------
# pattern
# -------------------------------------------------------------------------------------------------
def jac_pattern(spec: NlpSpec):
    """Structural (row, col) list in *emission* order plus the permutation to compressed-column order.

    Returns rows, cols (both in CCS order) and ``perm`` with ``vals_ccs = vals_emit[..., perm]``.
    """
    rows, cols = [], []
    r = 0
    for i in range(spec.nx):                                   # block 1
        rows.append(r); cols.append(spec.iX(i, 0)); r += 1
    for i in spec.terminal_idx:                                # block 2
        rows.append(r); cols.append(spec.iX(i, spec.N)); r += 1
    A, B = DYN_A[spec.dynamics], DYN_B[spec.dynamics]
    for k in range(spec.N):                                    # block 3
        for i in range(spec.nx):
            rows.append(r + i); cols.append(spec.iX(i, k + 1))     # d/dx_{k+1} = I
        for i in range(spec.nx):
            rows.append(r + i); cols.append(spec.iX(i, k))         # -(I) diagonal
        for (i, j) in A:
            if i != j:
                rows.append(r + i); cols.append(spec.iX(j, k))     # -dt*A off-diagonal
        for (i, j) in B:
            rows.append(r + i); cols.append(spec.iU(j, k))
        r += spec.nx
    if spec.use_slack:                                         # block 4
        for k in range(spec.N + 1):
            rows.append(r); cols.append(spec.iS(k)); r += 1
    for k in range(spec.N + 1):                                # block 5
        if spec.shape == "dot":
            rows += [r, r]; cols += [spec.iX(0, k), spec.iX(1, k)]; r += 1
        elif spec.use_slack:
            rows += [r, r, r, r]; cols += [spec.iX(0, k), spec.iX(1, k), spec.iX(2, k), spec.iS(k)]; r += 1
        else:
            for c in range(spec.n_body):
                rows += [r, r, r]; cols += [spec.iX(0, k), spec.iX(1, k), spec.iX(2, k)]; r += 1
    for i in range(spec.nu):                                   # block 6
        for k in range(spec.N):
            rows.append(r); cols.append(spec.iU(i, k)); r += 1
    assert r == spec.n_g, (r, spec.n_g)
    rows, cols = np.array(rows), np.array(cols)
    perm = np.lexsort((rows, cols))
    return rows[perm], cols[perm], perm


def bounds(spec: NlpSpec):
    lb, ub = [], []
    lb += list(spec.x0); ub += list(spec.x0)
    lb += [spec.goal[i] for i in spec.terminal_idx]; ub += [spec.goal[i] for i in spec.terminal_idx]
    lb += [0.0] * (spec.N * spec.nx); ub += [0.0] * (spec.N * spec.nx)
    if spec.use_slack:
        lb += [0.0] * (spec.N + 1); ub += [np.inf] * (spec.N + 1)
    n5 = (spec.N + 1) * spec.sdf_rows_per_knot
    lb += [0.0] * n5; ub += [np.inf] * n5
    for i in range(spec.nu):
        lb += [spec.control_bounds[i][0]] * spec.N; ub += [spec.control_bounds[i][1]] * spec.N
    return np.array(lb), np.array(ub)


# -------------------------------------------------------------------------------------------------
# evaluation
# -------------------------------------------------------------------------------------------------
def unpack(spec: NlpSpec, w: np.ndarray):
    P = w.shape[0]
    X = w[:, :spec.n_X].reshape(P, spec.N + 1, spec.nx)
    U = w[:, spec.n_X:spec.n_X + spec.n_U].reshape(P, spec.N, spec.nu)
    S = w[:, spec.n_X + spec.n_U:] if spec.use_slack else None
    return X, U, S


def eval_g_jac(spec: NlpSpec, w: np.ndarray, sdf: Optional[Callable] = None):
    """Returns g (P, n_g) and Jacobian values (P, nnz) in compressed-column order of ``jac_pattern``.

    ``sdf(points (n,2)) -> (s (n,), grad (n,2))`` is the learned SDF (l4casadi mode); ignored in casadi mode.
    """
    w = np.asarray(w, dtype=np.float64)
    P = w.shape[0]
    X, U, S = unpack(spec, w)
    g, jv = [], []
    g.append(X[:, 0, :]); jv.append(np.ones((P, spec.nx)))
    ti = spec.terminal_idx
    g.append(X[:, spec.N, ti]); jv.append(np.ones((P, len(ti))))
    # Euler defects  (core/runner.py:59-64)
    f, Av, Bv = dynamics_f(spec, X[:, :-1, :], U)
    r = X[:, 1:, :] - (X[:, :-1, :] + spec.dt * f)
    g.append(r.reshape(P, -1))
    A, B = DYN_A[spec.dynamics], DYN_B[spec.dynamics]
    blk = [np.ones((P, spec.N, spec.nx))]
    diag = -np.ones((P, spec.N, spec.nx))
    for n, (i, j) in enumerate(A):
        if i == j:
            diag[:, :, i] = diag[:, :, i] - spec.dt * Av[..., n]
    blk.append(diag)
    off = [-spec.dt * Av[..., n] for n, (i, j) in enumerate(A) if i != j]
    if off:
        blk.append(np.stack(off, axis=-1))
    blk.append(-spec.dt * Bv)
    jv.append(np.concatenate(blk, axis=-1).reshape(P, -1))
    if spec.use_slack:
        g.append(S); jv.append(np.ones((P, spec.N + 1)))
    # SDF rows (core/geometry.py:63-67, 107-117)
    pts, dpts = _footprint(spec, X)
    nb = pts.shape[2]
    flat = pts.reshape(-1, 2)
    if spec.sdf_mode == "casadi":
        s, gr = circles_sdf(spec, flat)
    else:
        s, gr = sdf(flat)
    s = np.asarray(s, dtype=np.float64).reshape(P, spec.N + 1, nb)
    gr = np.asarray(gr, dtype=np.float64).reshape(P, spec.N + 1, nb, 2)
    dth = gr[..., 0] * dpts[..., 0] + gr[..., 1] * dpts[..., 1]
    if spec.shape == "dot":
        g.append(s[:, :, 0]); jv.append(gr[:, :, 0, :].reshape(P, -1))
    elif spec.use_slack:
        e = np.exp(-ALPHA * s)                                   # core/utils.py:28-31 (un-stabilised)
        ssum = e.sum(axis=-1)
        m = -1.0 / ALPHA * np.log(ssum)
        om = e / ssum[..., None]
        g.append(m + S)
        row = np.stack([(om * gr[..., 0]).sum(-1), (om * gr[..., 1]).sum(-1), (om * dth).sum(-1), np.ones_like(m)], axis=-1)
        jv.append(row.reshape(P, -1))
    else:
        g.append(s.reshape(P, -1))
        row = np.stack([gr[..., 0], gr[..., 1], dth], axis=-1)
        jv.append(row.reshape(P, -1))
    g.append(U.transpose(0, 2, 1).reshape(P, -1)); jv.append(np.ones((P, spec.n_U)))
    g = np.concatenate(g, axis=1)
    jv = np.concatenate(jv, axis=1)
    _, _, perm = jac_pattern(spec)
    assert jv.shape[1] == perm.shape[0], (jv.shape, perm.shape)
    return g, jv[:, perm]


def eval_f_grad(spec: NlpSpec, w: np.ndarray):
    """Objective (core/runner.py:80-98) and its dense gradient (P, n_w)."""
    w = np.asarray(w, dtype=np.float64)
    P = w.shape[0]
    X, U, S = unpack(spec, w)
    dx = X[:, 1:, 0] - X[:, :-1, 0]
    dy = X[:, 1:, 1] - X[:, :-1, 1]
    seg = np.sqrt(dx * dx + dy * dy + EPS_PATH)
    f = seg.sum(axis=1)
    grad = np.zeros_like(w)
    gX = np.zeros((P, spec.N + 1, spec.nx))
    gX[:, 1:, 0] += dx / seg; gX[:, :-1, 0] -= dx / seg
    gX[:, 1:, 1] += dy / seg; gX[:, :-1, 1] -= dy / seg
    grad[:, :spec.n_X] = gX.reshape(P, -1)
    if spec.use_slack:
        f = f + spec.slack_penalty * (S * S).sum(axis=1)
        grad[:, spec.n_X + spec.n_U:] = 2 * spec.slack_penalty * S
    if spec.use_smooth:
        Us = U[:, :spec.N - 1, :]                      # last control unpenalised (runner.py:94-95)
        f = f + spec.smooth_weight * (Us * Us).sum(axis=(1, 2))
        gU = np.zeros_like(U)
        gU[:, :spec.N - 1, :] = 2 * spec.smooth_weight * Us
        grad[:, spec.n_X:spec.n_X + spec.n_U] = gU.reshape(P, -1)
    return f, grad


def multistart_guess(spec: NlpSpec, P: int, seed0: int = 1234) -> np.ndarray:
    """SURVEY.md section 8(d) multi-start definition: problem i uses default_rng(seed0 + i);
    X = linspace(x0, goal) (trajectory_initialization.py:54-55) with (x,y) pushed sideways by
    A*sin(pi t)*n_hat, A ~ U(-0.4, 0.4), plus N(0, 0.01^2) jitter; U = 0, slack = 0."""
    t = np.linspace(0.0, 1.0, spec.N + 1)
    base = spec.x0[None, :] + t[:, None] * (spec.goal - spec.x0)[None, :]
    d = spec.goal[:2] - spec.x0[:2]
    nrm = np.array([-d[1], d[0]]) / (np.linalg.norm(d) + 1e-12)
    w = np.zeros((P, spec.n_w))
    for i in range(P):
        rng = np.random.default_rng(seed0 + i)
        amp = rng.uniform(-0.4, 0.4)
        Xi = base.copy()
        Xi[:, :2] += amp * np.sin(np.pi * t)[:, None] * nrm[None, :] + rng.normal(0.0, 0.01, (spec.N + 1, 2))
        w[i, :spec.n_X] = Xi.reshape(-1)
    return w


# -------------------------------------------------------------------------------------------------
# Hessian of the Lagrangian  L(w) = sigma * f(w) + lam^T g(w)      
# -------------------------------------------------------------------------------------------------
# IPOPT's default is the exact Hessian (core/runner.py:113-125 sets no hessian_approximation), which CasADi
# assembles from second derivatives of the same expressions: dynamics (core/dynamics.py), footprint + soft-min
# (core/geometry.py:78-117, core/utils.py:28-31), the learned SDF's jac_adj1 (_l4c_generated/nn_sdf.cpp:88-104) and
# the objective (core/runner.py:80-98).  Returned: values of the structural non-zeros of the UPPER triangle in
# compressed-column order [upstream-memory: CasADi's nlp_hess_l is declared "triu:hess:gamma:x:x"].

# structurally non-zero second derivatives of f(x,u) w.r.t. z = (x, u), pairs (a <= b), u_j at index nx + j
DYN_H = {
    "point_1st": [], "point_2nd": [],
    "unicycle": [(2, 2), (2, 3)],
    "unicycle_2nd": [(2, 2), (2, 3)],
    "ackermann": [(2, 2), (2, 4), (3, 3), (3, 4)],
    "ackermann_2nd": [(2, 2), (2, 4), (3, 3), (3, 4), (3, 6), (3, 7), (4, 6)],
}


def dynamics_hess(spec: NlpSpec, x: np.ndarray, u: np.ndarray, lam: np.ndarray):
    """sum_i lam_i d2 f_i / dz_a dz_b for the pairs of DYN_H[spec.dynamics] (same order); x (..., nx), lam (..., nx)."""
    name = spec.dynamics
    if name in ("point_1st", "point_2nd"):
        return np.zeros(x.shape[:-1] + (0,))
    th = x[..., 2]
    c, s = np.cos(th), np.sin(th)
    if name in ("unicycle", "unicycle_2nd"):
        v = u[..., 0] if name == "unicycle" else x[..., 3]
        out = [-v * (lam[..., 0] * c + lam[..., 1] * s), -lam[..., 0] * s + lam[..., 1] * c]
    elif name == "ackermann":
        L = spec.wheelbase
        psi, v = x[..., 3], u[..., 0]
        t = np.tan(psi); sec2 = 1 + t * t
        out = [-v * (lam[..., 0] * c + lam[..., 1] * s), -lam[..., 0] * s + lam[..., 1] * c,
               lam[..., 2] * v * 2 * t * sec2 / L, lam[..., 2] * sec2 / L]
    elif name == "ackermann_2nd":
        L = spec.wheelbase
        psi, v, pd, a = x[..., 3], x[..., 4], x[..., 6], u[..., 0]
        t = np.tan(psi); sec2 = 1 + t * t
        q = 1.0 / (1.0 + psi * psi)
        dq = -2 * psi * q * q
        d2q = q * q * (8 * psi * psi * q - 2)
        l2, l4 = lam[..., 2], lam[..., 4]
        out = [-v * (lam[..., 0] * c + lam[..., 1] * s), -lam[..., 0] * s + lam[..., 1] * c,
               l2 * v * 2 * t * sec2 / L + l4 * (pd * v * d2q + a * 2 * t * sec2) / L,
               l2 * sec2 / L + l4 * pd * dq / L,
               l4 * v * dq / L, l4 * sec2 / L, l4 * q / L]
    else:
        raise ValueError(name)
    return np.stack(out, axis=-1)


def circles_hess(spec: NlpSpec, pts: np.ndarray):
    """Hessian (n, 3) = (hxx, hxy, hyy) of the analytic soft-min union."""
    return _analytic_union(spec, pts)[2]


def hess_pattern(spec: NlpSpec):
    """Structural (row, col), row <= col, of the Hessian of the Lagrangian in compressed-column order."""
    ent = set()
    nx, nu, N = spec.nx, spec.nu, spec.N
    zi = lambda a, k: spec.iX(a, k) if a < nx else spec.iU(a - nx, k)
    pose = [(0, 0), (0, 1), (1, 1)] if spec.shape == "dot" else [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)]
    for k in range(N + 1):
        for (a, b) in [(0, 0), (0, 1), (1, 1)] + pose:                 # objective + SDF rows
            ent.add((spec.iX(a, k), spec.iX(b, k)))
        if k < N:
            for (a, b) in DYN_H[spec.dynamics]:
                ent.add((zi(a, k), zi(b, k)))
            for a in (0, 1):
                for b in (0, 1):
                    ent.add((spec.iX(a, k), spec.iX(b, k + 1)))           # path length couples neighbouring knots
            if spec.use_smooth and k < N - 1:
                for j in range(nu):
                    ent.add((spec.iU(j, k), spec.iU(j, k)))
        if spec.use_slack:
            ent.add((spec.iS(k), spec.iS(k)))
    ent = sorted(ent, key=lambda rc: (rc[1], rc[0]))
    assert all(r <= c for r, c in ent)
    return np.array([r for r, _ in ent]), np.array([c for _, c in ent])


def eval_hess_lag(spec: NlpSpec, w: np.ndarray, sigma: np.ndarray, lam: np.ndarray,
                  sdf: Optional[Callable] = None, sdf_hess: Optional[Callable] = None):
    """Values (P, nnz_h) of the upper triangle of  sigma * hess f + sum_r lam_r * hess g_r  on ``hess_pattern``.

    ``sdf(points) -> (s, grad (n,2))`` and ``sdf_hess(points) -> (n,3) = (hxx, hxy, hyy)`` are the learned SDF
    (l4casadi mode); ignored in casadi mode.  sigma: (P,), lam: (P, n_g).
    """
    w = np.asarray(w, dtype=np.float64)
    P = w.shape[0]
    sigma = np.broadcast_to(np.asarray(sigma, dtype=np.float64), (P,))
    lam = np.asarray(lam, dtype=np.float64)
    X, U, S = unpack(spec, w)
    nx, nu, N = spec.nx, spec.nu, spec.N
    rows, cols = hess_pattern(spec)
    pos = {(int(r), int(c)): i for i, (r, c) in enumerate(zip(rows, cols))}
    H = np.zeros((P, len(rows)))

    def add(r, c, v):
        H[:, pos[(r, c) if r <= c else (c, r)]] += v
    zi = lambda a, k: spec.iX(a, k) if a < nx else spec.iU(a - nx, k)
    # row offsets
    off_dyn = nx + len(spec.terminal_idx)
    off_sdf = off_dyn + N * nx + ((N + 1) if spec.use_slack else 0)
    # dynamics: g = x_{k+1} - x_k - dt f  ->  -dt * sum_i lam_i hess f_i
    lam_dyn = lam[:, off_dyn:off_dyn + N * nx].reshape(P, N, nx)
    hd = dynamics_hess(spec, X[:, :-1, :], U, lam_dyn)
    for n, (a, b) in enumerate(DYN_H[spec.dynamics]):
        for k in range(N):
            add(zi(a, k), zi(b, k), -spec.dt * hd[:, k, n])
    # SDF rows
    pts, dpts = _footprint(spec, X)
    nb = pts.shape[2]
    flat = pts.reshape(-1, 2)
    if spec.sdf_mode == "casadi":
        s, gr = circles_sdf(spec, flat)
        hs = circles_hess(spec, flat)
    else:
        s, gr = sdf(flat)
        hs = sdf_hess(flat)
    s = np.asarray(s, np.float64).reshape(P, N + 1, nb)
    gr = np.asarray(gr, np.float64).reshape(P, N + 1, nb, 2)
    hs = np.asarray(hs, np.float64).reshape(P, N + 1, nb, 3)
    jx, jy = gr[..., 0], gr[..., 1]
    hxx, hxy, hyy = hs[..., 0], hs[..., 1], hs[..., 2]
    tx, ty = dpts[..., 0], dpts[..., 1]
    if spec.shape == "dot":
        ux = uy = np.zeros_like(tx)
    else:
        ux, uy = -(pts[..., 0] - X[..., 0:1]), -(pts[..., 1] - X[..., 1:2])       # d2 p / d theta2
    d3 = np.stack([jx, jy, jx * tx + jy * ty], axis=-1)                            # grad of s(p(q)) w.r.t. q = (x, y, theta)
    ax, ay = hxx * tx + hxy * ty, hxy * tx + hyy * ty
    # per-point 3x3 symmetric block, order (00, 01, 02, 11, 12, 22)
    Hb = np.stack([hxx, hxy, ax, hyy, ay, tx * ax + ty * ay + jx * ux + jy * uy], axis=-1)
    pairs = [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)]
    if spec.shape == "dot":
        lam_s = lam[:, off_sdf:off_sdf + N + 1]
        blk = lam_s[..., None] * Hb[:, :, 0, :]
    elif spec.use_slack:
        lam_s = lam[:, off_sdf:off_sdf + N + 1]
        e = np.exp(-ALPHA * s)
        om = e / e.sum(axis=-1, keepdims=True)
        mean = (om[..., None] * d3).sum(axis=2)                                    # (P, N+1, 3)
        blk = np.zeros((P, N + 1, 6))
        for n, (a, b) in enumerate(pairs):
            second = (om * d3[..., a] * d3[..., b]).sum(axis=2)
            blk[..., n] = (om * Hb[..., n]).sum(axis=2) - ALPHA * (second - mean[..., a] * mean[..., b])
        blk = lam_s[..., None] * blk
    else:
        lam_s = lam[:, off_sdf:off_sdf + (N + 1) * nb].reshape(P, N + 1, nb)
        blk = (lam_s[..., None] * Hb).sum(axis=2)
    for n, (a, b) in enumerate(pairs):
        if spec.shape == "dot" and (a == 2 or b == 2):
            continue
        for k in range(N + 1):
            add(spec.iX(a, k), spec.iX(b, k), blk[:, k, n])
    # objective (core/runner.py:80-98)
    dx, dy = X[:, 1:, 0] - X[:, :-1, 0], X[:, 1:, 1] - X[:, :-1, 1]
    r2 = dx * dx + dy * dy + EPS_PATH
    r3 = r2 * np.sqrt(r2)
    M = {(0, 0): (r2 - dx * dx) / r3, (0, 1): -dx * dy / r3, (1, 1): (r2 - dy * dy) / r3}
    for k in range(N):
        for (a, b), v in M.items():
            add(spec.iX(a, k), spec.iX(b, k), sigma * v[:, k])
            add(spec.iX(a, k + 1), spec.iX(b, k + 1), sigma * v[:, k])
        for a in (0, 1):
            for b in (0, 1):
                add(spec.iX(a, k), spec.iX(b, k + 1), -sigma * M[(min(a, b), max(a, b))][:, k])
    if spec.use_slack:
        for k in range(N + 1):
            add(spec.iS(k), spec.iS(k), sigma * 2.0 * spec.slack_penalty)
    if spec.use_smooth:
        for k in range(N - 1):
            for j in range(nu):
                add(spec.iU(j, k), spec.iU(j, k), sigma * 2.0 * spec.smooth_weight)
    return H
