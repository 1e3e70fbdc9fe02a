"""Initial state trajectories for the NLP (reference: core/trajectory_initialization.py).

* ``LinearInitializer``  - straight line in state space (trajectory_initialization.py:41-55).
* ``RRTInitializer``     - a rapidly-exploring random tree in the (x, y) plane over the exact obstacle SDF, shortcut and smoothed
  with a cubic spline, lifted to the full state with zeros (:58-236).  The reference draws from the unseeded stdlib ``random``;
  here every planner owns a seeded generator, which is what turns the planner into a multi-start family: problem i of a batch
  plans with seed ``seed0 + i`` and lands in its own homotopy class around the obstacles.
* ``DefaultInitializer`` - no initial guess (:239-249).

Host-side numpy by design: planning is sequential and runs once per start; the tree's nearest-neighbour search and the collision
checks along an edge are vectorised.  Quirks kept: the trajectory has exactly ``N`` rows (the caller passes ``solver.N + 1``,
scripts/run_benchmark.py:116); the footprint inflation is ``max_b |min(b_x, b_y)| + margin`` over the body points of a rectangle and
0 for every other footprint (:109-114); only (x, y) are planned, the other states start at zero (:228-231).
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import numpy as np


class LinearInitializer:
    def __init__(self, x0, x_goal, N: int):
        self.x0, self.x_goal, self.N = np.asarray(x0, float), np.asarray(x_goal, float), int(N)

    def get_initial_guess(self) -> np.ndarray:
        return np.linspace(self.x0, self.x_goal, self.N + 1)


class DefaultInitializer:
    def get_initial_guess(self) -> None:
        return None


class RRTInitializer:
    def __init__(self, N: int, x0, x_goal, dt: float, sdf_func: Optional[Callable], bounds, body_points: Optional[Sequence] = None,
                 rectangle: bool = True, step_size: float = 0.05, max_iter: int = 1000, margin: float = 0.01,
                 goal_sample_rate: float = 0.05, seed: int = 0):
        self.N, self.dt = int(N), float(dt)
        self.x0, self.x_goal = np.asarray(x0, float), np.asarray(x_goal, float)
        self.sdf_func = sdf_func                       # vectorised exact SDF: sdf(x[n], y[n]) -> [n]
        self.bounds = np.asarray(bounds, float)        # [[xmin, ymin], [xmax, ymax]]
        self.step_size, self.max_iter, self.margin = float(step_size), int(max_iter), float(margin)
        self.goal_sample_rate = float(goal_sample_rate)
        self.rng = np.random.default_rng(seed)
        if rectangle and body_points is not None and len(body_points):
            self.inflation = max(abs(float(np.min(b))) for b in body_points) + self.margin
        else:
            self.inflation = 0.0
        self.last_tree = None

    # ---- collision checks: points every ~step_size along the edge, all at once ------------------------------------
    def _collision_free(self, p1: np.ndarray, p2: np.ndarray) -> bool:
        if self.sdf_func is None:
            return True
        n = max(1, int(np.ceil(np.linalg.norm(p2 - p1) / self.step_size)))
        t = np.arange(n + 1)[:, None] / n
        pts = p1[None, :] + (p2 - p1)[None, :] * t
        return bool(np.all(self.sdf_func(pts[:, 0], pts[:, 1]) >= self.inflation))

    def _shortcut(self, path: np.ndarray) -> np.ndarray:
        out, i = [path[0]], 0
        while i < len(path) - 1:
            j = len(path) - 1
            while j > i + 1 and not self._collision_free(path[i], path[j]):
                j -= 1
            out.append(path[j])
            i = j
        return np.array(out)

    @staticmethod
    def _insert_intermediate_points(points: np.ndarray, max_angle_deg: float = 60.0) -> np.ndarray:
        out = [points[0]]
        for i in range(1, len(points) - 1):
            v1, v2 = points[i] - points[i - 1], points[i + 1] - points[i]
            c = np.dot(v1, v2) / (np.linalg.norm(v1) * np.linalg.norm(v2))
            if np.degrees(np.arccos(np.clip(c, -1.0, 1.0))) > max_angle_deg:
                out.append((points[i] + points[i - 1]) / 2)
            out.append(points[i])
        out.append(points[-1])
        return np.array(out)

    @staticmethod
    def _spline(points: np.ndarray, num: int) -> np.ndarray:
        if len(points) <= 2:
            return np.linspace(points[0], points[-1], num)
        from scipy.interpolate import CubicSpline
        s = np.linspace(0.0, 1.0, len(points))
        s_new = np.linspace(0.0, 1.0, num)
        return np.stack([CubicSpline(s, points[:, 0])(s_new), CubicSpline(s, points[:, 1])(s_new)], axis=1)

    def plan(self) -> np.ndarray:
        """Raw tree path (M, 2) from start to goal; RuntimeError after max_iter like the reference (:215-216)."""
        start, end = self.x0[:2].copy(), self.x_goal[:2].copy()
        pos = np.empty((self.max_iter + 2, 2)); parent = np.full(self.max_iter + 2, -1, np.int64)
        pos[0] = start
        n_nodes, final = 1, -1
        lo, hi = self.bounds[0], self.bounds[1]
        for _ in range(self.max_iter):
            ref = end if self.rng.random() < self.goal_sample_rate else self.rng.uniform(lo, hi)
            d = pos[:n_nodes] - ref
            near = int(np.argmin(np.einsum("ij,ij->i", d, d)))
            direction = ref - pos[near]
            norm = np.linalg.norm(direction)
            if norm == 0.0:
                continue
            new = pos[near] + direction / norm * self.step_size
            if not self._collision_free(pos[near], new):
                continue
            pos[n_nodes] = new; parent[n_nodes] = near; n_nodes += 1
            if np.linalg.norm(new - end) < self.step_size:
                pos[n_nodes] = end; parent[n_nodes] = n_nodes - 1; final = n_nodes; n_nodes += 1
                break
        if final < 0:
            raise RuntimeError("RRT failed to find a path within max_iter.")
        self.last_tree = (pos[:n_nodes].copy(), parent[:n_nodes].copy())
        idx, node = [], final
        while node >= 0:
            idx.append(node); node = parent[node]
        return pos[idx[::-1]]

    def get_initial_guess(self) -> np.ndarray:
        path = self._shortcut(self._insert_intermediate_points(self.plan()))
        traj = np.zeros((self.N, self.x0.shape[0]))
        traj[:, 0:2] = self._spline(path, self.N)
        return traj


def lift_path(X: np.ndarray, dynamic: str, dt: float, wheelbase: Optional[float] = None, control_bounds=None):
    """Fill the states the planner leaves at zero from the geometry of its (x, y) path, so that the forward-Euler defects of
    the position rows vanish: heading = direction of p[k+1] - p[k], speed = |p[k+1] - p[k]| / dt, turn rate / steering angle
    from the heading differences.  Not in the reference (its guess keeps zeros, core/trajectory_initialization.py:228-231); used
    by the batched solve, whose interior point has no restoration phase to recover from a kinematically inconsistent start.
    Returns (X, U) with U (N, nu) the controls the dynamics read directly (zeros elsewhere); the caller restores the pinned start state."""
    X = X.copy()
    n = X.shape[0]
    d = np.diff(X[:, :2], axis=0)
    d = np.vstack([d, d[-1:]])
    speed = np.linalg.norm(d, axis=1) / dt
    th = np.unwrap(np.arctan2(d[:, 1], d[:, 0]))
    om = np.append(np.diff(th), 0.0) / dt
    nu = {"point_1st": 2, "point_2nd": 2, "unicycle": 2, "unicycle_2nd": 2, "ackermann": 2, "ackermann_2nd": 2}[dynamic]
    U = np.zeros((n - 1, nu))
    if dynamic == "point_1st":
        U[:, 0:2] = d[:-1] / dt
    elif dynamic == "point_2nd":
        X[:, 2:4] = d / dt
    elif dynamic == "unicycle":
        X[:, 2] = th
        U[:, 0], U[:, 1] = speed[:-1], om[:-1]
    elif dynamic == "unicycle_2nd":
        X[:, 2], X[:, 3], X[:, 4] = th, speed, om
    else:
        L = float(wheelbase)
        psi = np.arctan(L * om / np.maximum(speed, 1e-6))
        X[:, 2], X[:, 3] = th, psi
        if dynamic == "ackermann":
            U[:, 0], U[:, 1] = speed[:-1], (np.append(np.diff(psi), 0.0) / dt)[:-1]
        else:
            X[:, 4] = speed                                  # the slot the reference's model reads as v (core/dynamics.py:131-148)
            X[:, 6] = np.append(np.diff(psi), 0.0) / dt
    if control_bounds is not None:
        for i, (lo, hi) in enumerate(control_bounds):
            U[:, i] = np.clip(U[:, i], 0.95 * lo, 0.95 * hi)
    return X, U


_PLAN_STATE = None      # (cfg, sdf, body, bounds, lift): set before the worker pool forks, read by the workers


def _plan_one(seed: int):
    cfg, sdf, body, bounds, lift = _PLAN_STATE
    from .problem import DYN_DIMS
    b, s = cfg.body, cfg.solver
    ini = s.initializer
    N = s.N
    nu = DYN_DIMS[b.dynamic][1]
    planner = RRTInitializer(N + 1, b.start_state, b.goal_state, s.dt, sdf, bounds, body_points=body, rectangle=b.shape == "rectangle",
                             step_size=ini.step_size, max_iter=ini.max_iter, margin=ini.margin, seed=seed)
    try:
        X = planner.get_initial_guess()
    except RuntimeError:
        X = np.linspace(np.asarray(b.start_state, float), np.asarray(b.goal_state, float), N + 1)
    U = np.zeros((N, nu))
    if lift:
        X, U = lift_path(X, b.dynamic, s.dt, getattr(b, "wheelbase", None), b.control_bounds)
        X[0] = np.asarray(b.start_state, float)              # pinned by the NLP (core/runner.py:60)
    return X, U


def rrt_multistart(cfg, P: int, first: int = 0, seed0: int = 1234, lift: bool = False, workers: Optional[int] = None) -> np.ndarray:
    """(P, n_w) initial decision vectors for a benchmark config: start i plans its own RRT path with seed ``seed0 + first + i``
    over the exact SDF of the YAML's obstacles (X from the planner, U and slack zero: core/runner.py:106-108).  Starts whose planner
    fails within ``max_iter`` fall back to the straight line.  ``lift=True`` fills heading / speed / steering from the path (``lift_path``).
    The planners are independent: they run on a pool of forked host processes (``workers``, default one per host core up to 32;
    numpy only - the children never touch CUDA), which is what keeps planning below the batched solve time."""
    import os
    global _PLAN_STATE
    from .problem import DYN_DIMS
    from .train import scene_sdf
    b, s = cfg.body, cfg.solver
    nx, nu = DYN_DIMS[b.dynamic]
    N = s.N
    n_X, n_U = nx * (N + 1), nu * N
    n_w = n_X + n_U + ((N + 1) if s.use_slack else 0)
    if b.shape == "rectangle":
        hl, hw = 0.5 * b.length, 0.5 * b.width
        body = [(-hl, -hw), (-hl, hw), (hl, hw), (hl, -hw)]                  # core/geometry.py:125-135
    else:
        body = None
    ini = s.initializer
    bounds = ini.rrt_bounds if ini.rrt_bounds is not None else [[-0.5, -0.5], [1.5, 1.5]]
    _PLAN_STATE = (cfg, scene_sdf(cfg), body, bounds, lift)
    seeds = [seed0 + first + i for i in range(P)]
    if workers is None:
        workers = min(32, os.cpu_count() or 1)
    workers = min(workers, P)
    if workers > 1 and P >= 4:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(workers) as pool:
            plans = pool.map(_plan_one, seeds, chunksize=max(1, P // (4 * workers)))
    else:
        plans = [_plan_one(sd) for sd in seeds]
    w = np.zeros((P, n_w), np.float32)
    for i, (X, U) in enumerate(plans):
        w[i, :n_X] = X.reshape(-1)
        w[i, n_X:n_X + n_U] = U.reshape(-1)
    return w
