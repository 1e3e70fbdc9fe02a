"""nlotrajectories_b200 - B200-native (sm_100a) NLP-evaluation hot path of pyMSE/NLOTrajectories.

Only what the path needs: ``csrc/`` (hand-written CUDA + the C ABI of include/nlo_b200.h) and the thin
Python host that mirrors the reference's interface for this path.  Importing the package does not
load CUDA; the first compute call does and raises if the library or a GPU is missing (no CPU fallback).
"""
from .config import Config, ConfigError  # noqa: F401
from .lib import NloError  # noqa: F401

__all__ = ["Config", "ConfigError", "NloError", "SdfWeights", "LearnedSDF", "NNObstacle", "NlpProblem", "RunBenchmark"]


def __getattr__(name):
    if name in ("SdfWeights", "LearnedSDF", "NNObstacle"):
        from . import sdf
        return getattr(sdf, name)
    if name == "NlpProblem":
        from .problem import NlpProblem
        return NlpProblem
    if name == "RunBenchmark":
        from .runner import RunBenchmark
        return RunBenchmark
    raise AttributeError(name)
