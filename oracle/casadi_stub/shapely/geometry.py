"""Stand-in for shapely.geometry: the reference only needs it for exact polygon SDFs
(core/sdf/casadi.py:135-148), which are training-target generators outside the hot path."""


class Point:  # pragma: no cover - never evaluated on the hot path
    def __init__(self, *a):
        raise NotImplementedError("shapely is not available; exact polygon SDF is out of scope")


class Polygon:  # pragma: no cover
    def __init__(self, *a, **k):
        self.args = a
