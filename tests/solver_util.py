"""CPU evaluator for the batched interior-point solver: the numpy oracle behind the solver's evaluator interface
(test infrastructure - lets the solver logic be exercised without a GPU and gives the CPU reference solve)."""
import numpy as np
import torch

from oracle import nlp_oracle as no
from oracle import sdf_oracle as so


class OracleEvaluator:
    def __init__(self, spec, net=None):
        self.spec = spec
        self.n_w, self.n_g = spec.n_w, spec.n_g
        self.rows, self.cols, _ = no.jac_pattern(spec)
        self.hrows, self.hcols = no.hess_pattern(spec)
        self.evals = 0
        if net is not None:
            n64 = net.astype(np.float64)
            self.sdf = lambda Q: so.value_jac(n64, Q)
            self.sdf_h = lambda Q: (lambda H: np.stack([H[:, 0, 0], H[:, 0, 1], H[:, 1, 1]], -1))(so.jac_adj1(n64, Q, np.ones(len(Q))))
        else:
            self.sdf = self.sdf_h = None

    def eval(self, w, want_jac=True):
        wn = w.numpy()
        P = wn.shape[0]
        g, jv = no.eval_g_jac(self.spec, wn, self.sdf)
        f, gr = no.eval_f_grad(self.spec, wn)
        self.evals += 1
        if not want_jac:
            return torch.from_numpy(f), None, torch.from_numpy(g), None
        J = np.zeros((P, self.n_g, self.n_w))
        J[:, self.rows, self.cols] = jv
        return torch.from_numpy(f), torch.from_numpy(gr), torch.from_numpy(g), torch.from_numpy(J)

    def hess(self, w, sigma, lam):
        hv = no.eval_hess_lag(self.spec, w.numpy(), sigma.numpy(), lam.numpy(), self.sdf, self.sdf_h)
        P = hv.shape[0]
        H = np.zeros((P, self.n_w, self.n_w))
        H[:, self.hrows, self.hcols] = hv
        H[:, self.hcols, self.hrows] = hv
        return torch.from_numpy(H)
