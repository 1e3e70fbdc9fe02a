"""Benchmark YAML schema.  Parses the reference's ``benchmarks/*.yaml`` unchanged so that
``run-benchmark --config <yaml>`` stays drop-in (reference: core/config.py:26-222).  New knobs
(batch size, GPUs, seed, weights) are additive and live on the command line, not in the YAML.

Plain dataclasses instead of pydantic models; field names, defaults and the accepted literals are
the reference's.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from pathlib import Path
from typing import Any, List, Optional, Tuple

import numpy as np
import yaml

DYNAMICS = ("point_1st", "point_2nd", "unicycle", "unicycle_2nd", "ackermann", "ackermann_2nd")   # core/dynamics.py:7-13
SHAPES = ("dot", "rectangle", "triangle")                                                           # core/geometry.py:17-20
GOAL_MODES = ("center", "any_point")
OBSTACLE_TYPES = ("circle", "square", "polygon", "elliptical_ring", "trapezoid", "discr_s")
MODEL_TYPES = ("mlp", "fourier", "siren")
INIT_MODES = ("default", "linear", "rrt")


class ConfigError(ValueError):
    pass


def _req(d: dict, key: str, where: str):
    if key not in d:
        raise ConfigError(f"{where}: field required: {key}")
    return d[key]


def _lit(v, allowed, where):
    if v not in allowed:
        raise ConfigError(f"{where}: {v!r} is not one of {allowed}")
    return v


@dataclass
class BodyConfig:                       # core/config.py:26-50
    shape: str
    dynamic: str
    start_state: List[float]
    goal_state: List[float]
    control_bounds: List[Tuple[float, float]]
    goal_mode: str = "center"
    length: Optional[float] = None
    width: Optional[float] = None
    wheelbase: Optional[float] = None

    @staticmethod
    def parse(d: dict) -> "BodyConfig":
        cb = _req(d, "control_bounds", "body")
        try:
            bounds = [(float(lo), float(hi)) for lo, hi in cb]
        except (TypeError, ValueError):
            raise ConfigError("body.control_bounds: expected a list of [min, max] pairs") from None
        return BodyConfig(
            shape=_lit(_req(d, "shape", "body"), SHAPES, "body.shape"),
            dynamic=_lit(_req(d, "dynamic", "body"), DYNAMICS, "body.dynamic"),
            start_state=[float(v) for v in _req(d, "start_state", "body")],
            goal_state=[float(v) for v in _req(d, "goal_state", "body")],
            control_bounds=bounds,
            goal_mode=_lit(d.get("goal_mode", "center"), GOAL_MODES, "body.goal_mode"),
            length=d.get("length"), width=d.get("width"), wheelbase=d.get("wheelbase"))


@dataclass
class ObstacleConfig:                   # core/config.py:53-149 (discriminated on ``type``)
    type: str
    params: dict

    @staticmethod
    def parse(d: dict) -> "ObstacleConfig":
        t = _lit(_req(d, "type", "obstacles[]"), OBSTACLE_TYPES, "obstacles[].type")
        need = {"circle": ("center", "radius"), "square": ("center", "size"), "polygon": ("points",),
                "elliptical_ring": ("center", "semi_axes", "width"), "trapezoid": ("points",),
                "discr_s": ("center", "semi_axes", "width")}[t]
        for k in need:
            _req(d, k, f"obstacles[{t}]")
        p = {k: v for k, v in d.items() if k != "type"}
        p.setdefault("margin", 0.0)
        return ObstacleConfig(t, p)


@dataclass
class InitializerConfig:                # core/config.py:152-185
    mode: str = "linear"
    rrt_bounds: Optional[List[List[float]]] = None
    step_size: float = 0.05
    max_iter: int = 1000
    margin: float = 0.01

    @staticmethod
    def parse(lst) -> "InitializerConfig":
        if not lst:
            return InitializerConfig()
        d = lst[0]                                        # ``choice`` = first entry (config.py:183-185)
        mode = _lit(d.get("mode"), INIT_MODES, "solver.initializer[].mode")
        if mode == "rrt" and "rrt_bounds" not in d:
            raise ConfigError("solver.initializer[rrt]: field required: rrt_bounds")
        return InitializerConfig(mode, d.get("rrt_bounds"), float(d.get("step_size", 0.05)), int(d.get("max_iter", 1000)),
                                 float(d.get("margin", 0.01)))


@dataclass
class SolverConfig:                     # core/config.py:188-200
    mode: str
    type: str
    N: int = 20
    dt: float = 0.1
    use_slack: bool = False
    slack_penalty: Optional[float] = 1000.0
    use_smooth: bool = False
    smooth_weight: float = 10.0
    enforce_heading: bool = True
    initializer: InitializerConfig = field(default_factory=InitializerConfig)

    @staticmethod
    def parse(d: dict) -> "SolverConfig":
        N = int(d.get("N", 20))
        if N < 1:
            raise ConfigError("solver.N: must be >= 1")
        return SolverConfig(
            mode=_lit(_req(d, "mode", "solver"), ("casadi", "l4casadi"), "solver.mode"),
            type=_lit(_req(d, "type", "solver"), ("ipopt", "sqpmethod"), "solver.type"),
            N=N, dt=float(d.get("dt", 0.1)), use_slack=bool(d.get("use_slack", False)),
            slack_penalty=d.get("slack_penalty", 1000.0), use_smooth=bool(d.get("use_smooth", False)),
            smooth_weight=float(d.get("smooth_weight", 10.0)), enforce_heading=bool(d.get("enforce_heading", True)),
            initializer=InitializerConfig.parse(d.get("initializer")))


@dataclass
class ModelConfig:                      # core/config.py:203-212
    type: str = "mlp"
    hidden_dim: int = 64
    num_hidden_layers: int = 3
    activation_function: str = "ReLU"
    omega_0: float = 30.0
    n_samples: int = 200_000
    boundary_fraction: float = 0.3
    surface_loss_weight: float = 1.0
    eikonal_loss_weight: float = 1.0

    @staticmethod
    def parse(d: dict) -> "ModelConfig":
        m = ModelConfig(**{k: d[k] for k in ModelConfig.__dataclass_fields__ if k in d})
        _lit(m.type, MODEL_TYPES, "model.type")
        if m.hidden_dim < 1 or m.num_hidden_layers < 1:
            raise ConfigError("model: hidden_dim and num_hidden_layers must be >= 1")
        return m

    def n_hidden_mats(self) -> int:
        """H x H matrices of the network the reference builds (scripts/run_benchmark.py:64-83):
        mlp: l4c.naive.MultiLayerPerceptron(2,H,1,L) has L-1; fourier/siren get num_layers = L+2, i.e. L."""
        return self.num_hidden_layers - 1 if self.type == "mlp" else self.num_hidden_layers


@dataclass
class Config:                           # core/config.py:215-222
    body: BodyConfig
    obstacles: List[ObstacleConfig]
    solver: SolverConfig
    model: ModelConfig
    raw: dict = field(default_factory=dict, repr=False)

    @staticmethod
    def parse(d: dict) -> "Config":
        return Config(BodyConfig.parse(_req(d, "body", "config")), [ObstacleConfig.parse(o) for o in _req(d, "obstacles", "config")],
                      SolverConfig.parse(_req(d, "solver", "config")), ModelConfig.parse(_req(d, "model", "config")), raw=d)

    @staticmethod
    def load(path) -> "Config":
        with open(path, "r") as fh:
            return Config.parse(yaml.safe_load(fh))

    def circles(self) -> List[Tuple[float, float, float, float, int]]:
        """Analytic obstacles for ``solver.mode: casadi`` as (cx, cy, radius | size, margin, kind): circles (kind 0,
        core/sdf/casadi.py:27-45) and squares (kind 1, :48-115), combined by the soft-min union (:385-386).  Polygons and
        rings appear in the shipped benchmarks only behind a learned SDF (mode l4casadi)."""
        out = []
        for o in self.obstacles:
            c = o.params["center"] if o.type in ("circle", "square") else None
            if o.type == "circle":
                out.append((float(c[0]), float(c[1]), float(o.params["radius"]), float(o.params.get("margin", 0.0)), 0))
            elif o.type == "square":
                out.append((float(c[0]), float(c[1]), float(o.params["size"]), float(o.params.get("margin", 0.0)), 1))
            else:
                raise NotImplementedError(f"solver.mode casadi with obstacle type {o.type!r} is outside the CUDA hot path")
        return out
