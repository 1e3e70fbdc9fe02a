"""Stand-in for l4casadi (third-party, unpinned; reference README.md:24).  TEST INFRASTRUCTURE.

``L4CasADi(model)`` called on a symbolic (N,2) MX returns an (N,1) MX of opaque sympy function
applications ``sdf(x, y)`` whose derivatives are the opaque functions ``sdf_dx`` / ``sdf_dy``:
exactly the role the ``nn_sdf`` / ``jac_nn_sdf`` externals play in the reference's MX graph
(_l4c_generated/nn_sdf.cpp:57-70).  make_golden.py binds them to the numpy SDF oracle when it
lambdifies.
"""
import sympy as sp

from . import naive  # noqa: F401


class sdf_dxx(sp.Function):
    nargs = 2


class sdf_dxy(sp.Function):
    nargs = 2


class sdf_dyy(sp.Function):
    nargs = 2


class sdf_dx(sp.Function):
    nargs = 2

    def fdiff(self, argindex=1):
        return sdf_dxx(*self.args) if argindex == 1 else sdf_dxy(*self.args)


class sdf_dy(sp.Function):
    nargs = 2

    def fdiff(self, argindex=1):
        return sdf_dxy(*self.args) if argindex == 1 else sdf_dyy(*self.args)


class sdf(sp.Function):
    nargs = 2

    def fdiff(self, argindex=1):
        return sdf_dx(*self.args) if argindex == 1 else sdf_dy(*self.args)


class L4CasADi:
    def __init__(self, model=None, device="cpu", name="nn_sdf", **kw):
        self.model = model
        self.device = device
        self.name = name

    def __call__(self, coords):
        import casadi as ca
        n = coords.size1()
        return ca.MX(sp.Matrix(n, 1, [sdf(coords.m[i, 0], coords.m[i, 1]) for i in range(n)]))
