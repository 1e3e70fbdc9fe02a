"""Batched RRT on the device: the tree search of ``core/trajectory_initialization.py:175-216`` for P planners at
once.  On a CUDA device the search is ONE kernel of the library (``csrc/rrt_kernels.cu``, C ABI ``nlo_rrt_paths``): a warp per
planner runs its whole tree search against the exact obstacle SDF and extracts its path.  ``batched_rrt_trees`` is the same algorithm,
draw for draw, as lock-step torch tensor operations - one sampled point, one nearest-node search, one steer and one collision check
per planner per iteration: the readable form, the CPU test double, and what the kernel is checked against.  The exact obstacle SDF
(circles, squares, polygons, elliptical half-rings) is evaluated on the same device.  Path post-processing (shortcut, corner splitting, cubic spline, optional lifting) stays with ``initializer.py`` on
the host, where it is a few milliseconds per path.

The host planners of ``initializer.rrt_multistart`` (a pool of forked processes) are the better tool for tens of starts; this one
is for thousands: its cost per iteration is a handful of launches whatever P is.

Random numbers come from a counter hash of (seed of the start, iteration, draw), so start i depends on ``seed0 + first + i`` only -
shards of a batch plan the same trees as the whole batch - and not on torch's generator state.
"""
from __future__ import annotations

from typing import Optional

import numpy as np


# ---- exact SDFs as torch functions of (..., 2) points (same definitions as train.py, which follows core/sdf/casadi.py) ---------
def _polygon_sdf(pts, poly, margin):
    import torch
    a = poly
    b = torch.roll(poly, -1, dims=0)
    px, py = pts[..., 0:1], pts[..., 1:2]
    ex, ey = b[:, 0] - a[:, 0], b[:, 1] - a[:, 1]
    wx, wy = px - a[:, 0], py - a[:, 1]
    t = torch.clamp((wx * ex + wy * ey) / torch.clamp(ex * ex + ey * ey, min=1e-300), 0.0, 1.0)
    d = torch.sqrt(((wx - t * ex) ** 2 + (wy - t * ey) ** 2).amin(dim=-1))
    cond = (a[:, 1] > py) != (b[:, 1] > py)
    xint = a[:, 0] + (py - a[:, 1]) * ex / torch.where(ey == 0, torch.ones_like(ey), ey)
    inside = ((cond & (px < xint)).sum(dim=-1) % 2) == 1
    return torch.where(inside, -d, d) - margin


def torch_scene_sdf(cfg, device, dtype=None):
    """Exact SDF of the YAML's obstacle list (union = min) as ``f(points (..., 2)) -> (...)`` on ``device``."""
    import torch
    from .train import elliptic_ring_points
    dtype = dtype or torch.float64
    fns = []
    for ob in cfg.obstacles:
        p = ob.params
        m = float(p.get("margin", 0.0))
        if ob.type == "circle":
            c = torch.tensor(p["center"], dtype=dtype, device=device); r = float(p["radius"]) + m
            fns.append(lambda q, c=c, r=r: torch.linalg.vector_norm(q - c, dim=-1) - r)
        elif ob.type == "square":
            c = torch.tensor(p["center"], dtype=dtype, device=device); half = float(p["size"]) / 2 + m

            def sq(q, c=c, half=half):
                d = (q - c).abs() - half
                return torch.linalg.vector_norm(torch.clamp(d, min=0.0), dim=-1) + torch.clamp(d.amax(dim=-1), max=0.0)
            fns.append(sq)
        elif ob.type in ("polygon", "trapezoid", "elliptical_ring"):
            if ob.type == "elliptical_ring":
                pts = elliptic_ring_points(p["center"], p["semi_axes"], p["width"], p.get("angle", np.pi), p.get("num_arc_points", 15),
                                           p.get("rotation", 0.0))
            else:
                pts = p["points"]
            poly = torch.tensor(np.asarray(pts, float), dtype=dtype, device=device)
            fns.append(lambda q, poly=poly, m=m: _polygon_sdf(q, poly, m))
        else:
            raise NotImplementedError(f"exact SDF of obstacle type {ob.type!r}")
    return lambda q: torch.stack([f(q) for f in fns], dim=0).amin(dim=0)


# ---- counter-hash uniforms ---------------------------------------------------------------------------------------------------
_M31 = (1 << 31) - 1


def _uniform(seed, it: int, k: int):
    """(P,) uniforms in [0, 1) from int64 seeds: three rounds of a multiply / xor-shift mix kept below 2^62 (no int64 overflow)."""
    x = (seed * 1103515245 + (it * 40503 + k * 9973 + 12345)) & _M31
    for mult, add in ((1664525, 1013904223), (22695477, 1), (1103515245, 12345)):
        x = (x ^ (x >> 15)) & _M31
        x = (x * mult + add) & _M31
    x = (x ^ (x >> 13)) & _M31
    return x.double() / float(1 << 31)


def batched_rrt_trees(sdf, start, goal, bounds, P: int, seeds, step_size: float, max_iter: int, inflation: float,
                      goal_sample_rate: float = 0.05, device=None, check_every: int = 64):
    """Grow P trees from ``start`` towards ``goal`` (2-vectors).  Returns host arrays ``pos (P, M, 2)``, ``parent (P, M)``,
    ``final (P,)`` (index of the goal node, -1 where the planner ran out of iterations)."""
    import torch
    dev = device
    f64 = torch.float64
    M = max_iter + 3                                                  # start + one node per iteration + goal + a scratch slot
    pos = torch.zeros((P, M, 2), dtype=f64, device=dev)
    parent = torch.full((P, M), -1, dtype=torch.int64, device=dev)
    start_t = torch.tensor(np.asarray(start, float)[:2], dtype=f64, device=dev)
    goal_t = torch.tensor(np.asarray(goal, float)[:2], dtype=f64, device=dev)
    lo = torch.tensor(np.asarray(bounds, float)[0], dtype=f64, device=dev)
    hi = torch.tensor(np.asarray(bounds, float)[1], dtype=f64, device=dev)
    pos[:, 0] = start_t
    n_nodes = torch.ones(P, dtype=torch.int64, device=dev)
    final = torch.full((P,), -1, dtype=torch.int64, device=dev)
    seeds_t = torch.as_tensor(np.asarray(seeds, np.int64), device=dev)
    rows = torch.arange(P, device=dev)
    tt = torch.tensor([0.0, 0.5, 1.0], dtype=f64, device=dev)[None, :, None]
    for it in range(max_iter):
        active = final < 0
        if it % check_every == 0 and it > 0 and not bool(active.any()):
            break
        to_goal = _uniform(seeds_t, it, 0) < goal_sample_rate
        sample = lo + (hi - lo) * torch.stack([_uniform(seeds_t, it, 1), _uniform(seeds_t, it, 2)], dim=1)
        ref = torch.where(to_goal[:, None], goal_t[None, :], sample)
        n_max = min(M - 1, it + 2)                                    # no tree has more nodes than this yet
        d2 = ((pos[:, :n_max] - ref[:, None, :]) ** 2).sum(-1)
        d2 = torch.where(torch.arange(n_max, device=dev)[None, :] < n_nodes[:, None], d2, torch.full_like(d2, float("inf")))
        near = d2.argmin(dim=1)
        p_near = pos[rows, near]
        direction = ref - p_near
        norm = torch.linalg.vector_norm(direction, dim=1)
        new = p_near + direction / torch.clamp(norm, min=1e-300)[:, None] * step_size
        edge = p_near[:, None, :] + (new - p_near)[:, None, :] * tt     # both ends and the midpoint of a step-long edge
        free = (sdf(edge) >= inflation).all(dim=1)
        ok = active & free & (norm > 0)
        slot = torch.where(ok, n_nodes, torch.full_like(n_nodes, M - 1))   # rejected planners write to the scratch slot M - 1
        pos[rows, slot] = torch.where(ok[:, None], new, pos[rows, slot])
        parent[rows, slot] = torch.where(ok, near, parent[rows, slot])
        n_nodes = n_nodes + ok.long()
        reached = ok & (torch.linalg.vector_norm(new - goal_t[None, :], dim=1) < step_size)
        gslot = torch.where(reached, n_nodes, torch.full_like(n_nodes, M - 1))
        pos[rows, gslot] = torch.where(reached[:, None], goal_t[None, :].expand(P, 2), pos[rows, gslot])
        parent[rows, gslot] = torch.where(reached, n_nodes - 1, parent[rows, gslot])
        final = torch.where(reached, n_nodes, final)
        n_nodes = n_nodes + reached.long()
    return pos.cpu().numpy(), parent.cpu().numpy(), final.cpu().numpy()


def scene_obstacles(cfg):
    """The YAML's obstacle list as the ``nlo_rrt_obstacle`` records + vertex array of ``nlo_rrt_paths``."""
    from . import lib as _lib
    from .train import elliptic_ring_points
    obs, verts = [], []
    for ob in cfg.obstacles:
        p = ob.params
        m = float(p.get("margin", 0.0))
        if ob.type == "circle":
            obs.append((0, 0, 0, p["center"][0], p["center"][1], float(p["radius"]), m))
        elif ob.type == "square":
            obs.append((1, 0, 0, p["center"][0], p["center"][1], float(p["size"]), m))
        elif ob.type in ("polygon", "trapezoid", "elliptical_ring"):
            pts = (elliptic_ring_points(p["center"], p["semi_axes"], p["width"], p.get("angle", np.pi), p.get("num_arc_points", 15),
                                        p.get("rotation", 0.0)) if ob.type == "elliptical_ring" else p["points"])
            obs.append((2, len(verts), len(pts), 0.0, 0.0, 0.0, m))
            verts.extend([(float(x), float(y)) for x, y in pts])
        else:
            raise NotImplementedError(f"exact SDF of obstacle type {ob.type!r}")
    arr = (_lib.RrtObstacle * len(obs))()
    for i, (kind, first, n, cx, cy, size, m) in enumerate(obs):
        arr[i] = _lib.RrtObstacle(kind, first, n, 0, float(cx), float(cy), float(size), float(m))
    return arr, np.ascontiguousarray(np.asarray(verts, np.float64).reshape(-1, 2))


def cuda_rrt_paths(cfg, seeds, start, goal, bounds, step_size: float, max_iter: int, inflation: float, goal_sample_rate: float = 0.05,
                   device_index: int = 0, max_path: int = 512, postprocess: bool = False):
    """The tree search as one CUDA kernel (``nlo_rrt_paths``).  Returns a list of (n_i, 2) fp64 paths root first (None where the planner
    ran out of iterations).  ``postprocess``: the kernel also inserts the corner midpoints and shortcuts the path (what
    ``RRTInitializer._shortcut(_insert_intermediate_points(path))`` does on the host)."""
    import ctypes as C
    from . import lib as _lib
    L = _lib.load()
    _lib.require_gpu()
    obs, verts = scene_obstacles(cfg)
    seeds = np.ascontiguousarray(seeds, np.int64)
    P = len(seeds)
    f2 = lambda v: np.ascontiguousarray(np.asarray(v, np.float64)[:2])
    s2, g2, lo, hi = f2(start), f2(goal), f2(np.asarray(bounds, float)[0]), f2(np.asarray(bounds, float)[1])
    path = np.empty((P, max_path, 2)); plen = np.empty(P, np.int32)
    _lib.check(L.nlo_rrt_paths(obs, len(obs), verts.ctypes.data if len(verts) else None, len(verts), s2.ctypes.data, g2.ctypes.data,
                               lo.ctypes.data, hi.ctypes.data, seeds.ctypes.data, P, float(step_size), int(max_iter), float(inflation),
                               float(goal_sample_rate), int(max_path), int(bool(postprocess)), int(device_index), path.ctypes.data,
                               plen.ctypes.data))
    if (plen == -2).any():
        raise RuntimeError(f"RRT path longer than {max_path} nodes")
    return [path[i, :plen[i]].copy() if plen[i] > 0 else None for i in range(P)]


_POST_STATE = None      # (host planner, cfg, lift, nx, nu): set before the post-processing pool forks


def _post_one(arg):
    """(tree path or None when the planner failed, whether it is shortcut already) -> (X, U) like initializer._plan_one."""
    path, shortcut_done = arg
    from .initializer import lift_path
    host, cfg, lift, nx, nu = _POST_STATE
    b, s = cfg.body, cfg.solver
    N = s.N
    if path is None:
        X = np.linspace(np.asarray(b.start_state, float), np.asarray(b.goal_state, float), N + 1)
    else:
        if not shortcut_done:
            path = host._shortcut(host._insert_intermediate_points(path))
        X = np.zeros((N + 1, nx))
        X[:, 0:2] = host._spline(path, N + 1)
    U = np.zeros((N, nu))
    if lift:
        X, U = lift_path(X, b.dynamic, s.dt, getattr(b, "wheelbase", None), b.control_bounds)
        X[0] = np.asarray(b.start_state, float)
    return X, U


def _host_planner(cfg):
    from .initializer import RRTInitializer
    from .train import scene_sdf
    b, s = cfg.body, cfg.solver
    ini = s.initializer
    bounds = ini.rrt_bounds if ini.rrt_bounds is not None else [[-0.5, -0.5], [1.5, 1.5]]
    if b.shape == "rectangle":
        hl, hw = 0.5 * b.length, 0.5 * b.width
        body = [(-hl, -hw), (-hl, hw), (hl, hw), (hl, -hw)]
    else:
        body = None
    host = RRTInitializer(s.N + 1, b.start_state, b.goal_state, s.dt, scene_sdf(cfg), bounds, body_points=body, rectangle=b.shape == "rectangle",
                          step_size=ini.step_size, max_iter=ini.max_iter, margin=ini.margin)
    return host, bounds


def make_post_pool(cfg, lift: bool = False, workers: Optional[int] = None):
    """Fork the post-processing workers of ``rrt_multistart_device`` (numpy only).  Call it BEFORE the process creates its CUDA
    context or NCCL threads - forking a process that holds either is unsupported - and hand the pool to ``rrt_multistart_device``."""
    import multiprocessing as mp
    import os
    global _POST_STATE
    from .problem import DYN_DIMS
    nx, nu = DYN_DIMS[cfg.body.dynamic]
    _POST_STATE = (_host_planner(cfg)[0], cfg, lift, nx, nu)
    return mp.get_context("fork").Pool(workers or min(32, os.cpu_count() or 1))


def rrt_multistart_device(cfg, P: int, first: int = 0, seed0: int = 1234, lift: bool = False, device=None,
                          workers: Optional[int] = None, pool=None) -> np.ndarray:
    """``initializer.rrt_multistart`` with the tree search batched on ``device`` (default: CUDA when available, else CPU).  The raw
    tree paths are shortcut, split at sharp corners, splined and (optionally) lifted on the host exactly like the host planner's:
    on ``pool`` (``make_post_pool``, forked before CUDA was initialised) when given; on a pool forked here when this process has
    no CUDA context yet; in this process otherwise (a process that holds a CUDA context must not fork)."""
    import os
    import torch
    global _POST_STATE
    from .problem import DYN_DIMS
    if device is None:
        device = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")
    b, s = cfg.body, cfg.solver
    nx, nu = DYN_DIMS[b.dynamic]
    N = s.N
    n_X, n_U = nx * (N + 1), nu * N
    n_w = n_X + n_U + ((N + 1) if s.use_slack else 0)
    ini = s.initializer
    may_fork = pool is None and not torch.cuda.is_initialized()      # decided before the trees run on the device
    host, bounds = _host_planner(cfg)
    seeds = [seed0 + first + i for i in range(P)]
    shortcut_done = False
    if torch.device(device).type == "cuda":                   # the search as one kernel of the library: a warp per planner
        dev_index = torch.device(device).index
        paths = cuda_rrt_paths(cfg, seeds, b.start_state, b.goal_state, bounds, ini.step_size, ini.max_iter, host.inflation,
                               device_index=dev_index if dev_index is not None else torch.cuda.current_device(), postprocess=True)
        shortcut_done = True
    else:
        pos, parent, final = batched_rrt_trees(torch_scene_sdf(cfg, device), b.start_state, b.goal_state, bounds, P, seeds, ini.step_size,
                                               ini.max_iter, host.inflation, device=device)
        paths = []
        for i in range(P):
            if final[i] < 0:
                paths.append(None)
                continue
            idx, node = [], int(final[i])
            while node >= 0:
                idx.append(node); node = int(parent[i, node])
            paths.append(pos[i, idx[::-1]])
    if workers is None:
        workers = min(32, os.cpu_count() or 1)
    workers = min(workers, P)
    if pool is not None:
        plans = pool.map(_post_one, [(pth, shortcut_done) for pth in paths], chunksize=max(1, P // (4 * workers)))
    elif may_fork and str(device) == "cpu" and workers > 1 and P >= 8:
        import multiprocessing as mp
        _POST_STATE = (host, cfg, lift, nx, nu)
        with mp.get_context("fork").Pool(workers) as own:
            plans = own.map(_post_one, [(pth, shortcut_done) for pth in paths], chunksize=max(1, P // (4 * workers)))
    else:
        _POST_STATE = (host, cfg, lift, nx, nu)
        plans = [_post_one((pth, shortcut_done)) for pth in paths]
    w = np.zeros((P, n_w), np.float32)
    for i, (X, U) in enumerate(plans):
        w[i, :n_X] = X.reshape(-1)
        w[i, n_X:n_X + n_U] = U.reshape(-1)
    return w
