class _Patch:
    def __init__(self, *a, **k):
        pass


class Circle(_Patch):
    pass


class Polygon(_Patch):
    pass


class Rectangle(_Patch):
    pass
