"""Oracle restatement (TEST INFRASTRUCTURE) of the fully Bayesian GP kernel ``Sober`` hands to ``recombination`` when
``fbgp=True`` (SOBER/_sober.py:63-65): ``FullyBayesianGP.marginal_predictive_covariance``
(SOBER/FBGP/_fully_Bayesian_gp.py:354-371).  SURVEY.md 8(f) row 4.

    mu_q(x)   = posterior mean of the GP with the q-th distilled hypersample       (batch_predict, :304-322 -- gpytorch)
    E(x)      = sum_q w_q mu_q(x)
    cov(x, y) = W sum_q w_q (mu_q(x) - E(x)) (mu_q(y) - E(y)),   W = 1 / (1 - sum_q w_q^2)

a rank-n_qd kernel of per-point means.  ``batch_predict`` needs gpytorch (absent here): the stand-in below predicts with
the gpytorch-free exact-GP restatements of oracle/kernels.py, one per hypersample.

A finding about the reference, kept as is: ``Ex = self.w_qd @ mu_x`` is a 1-D @ N-D matmul, so the method only accepts
2-D inputs.  ``SOBER/_rchq.py:124`` calls the kernel with the candidates reshaped to (E, S, d); ``w_qd @ mu_y`` then
raises (``Expected size for first two dimensions of batch2 tensor ...``) unless n_qd == E by accident -- the reference
cannot run this kernel through its own loop branch (N > 2 b), only through the direct branch.  ``strict=True`` keeps
that behaviour; ``strict=False`` is the same formula written to broadcast, the oracle for the product's generic path
(which evaluates the callable on 2-D tiles and therefore works).
"""
import torch


class FullyBayesianGP:
    def __init__(self, models, w_qd, strict=True):
        self.models = list(models)
        self.w_qd = torch.as_tensor(w_qd, dtype=torch.float64)
        self.strict = strict
        self.is_fbgp = True

    def batch_predict(self, x_test):
        """(n_qd, ...) means and variances; stand-in for SOBER/FBGP/_fully_Bayesian_gp.py:304-322."""
        flat = x_test.reshape(-1, x_test.shape[-1])
        mu, var = [], []
        for model in self.models:
            dist = model(flat)
            mu.append(dist.mean.reshape(x_test.shape[:-1]))
            var.append(dist.variance.clamp_min(0).reshape(x_test.shape[:-1]))
        return torch.stack(mu), torch.stack(var)

    def marginal_predictive_mean(self, x_test):                       # :346-352
        mu_batch, _ = self.batch_predict(x_test)
        return self._avg(mu_batch)

    def _avg(self, mu):
        if self.strict:
            return self.w_qd @ mu                                     # 1-D @ N-D, as written in the reference
        return torch.tensordot(self.w_qd.to(mu), mu, dims=([0], [0]))

    def marginal_predictive_covariance(self, x_test, y_test):         # :354-371
        mu_x, _ = self.batch_predict(x_test)
        mu_y, _ = self.batch_predict(y_test)
        Ex = self._avg(mu_x)
        Ey = self._avg(mu_y)
        W = 1 / (1 - self.w_qd.pow(2).sum())
        if self.strict or y_test.dim() == 2:
            return W * (self.w_qd.to(mu_x).unsqueeze(1) * (mu_x - Ex.unsqueeze(0))).T @ (mu_y - Ey.unsqueeze(0))
        # (L, q) x (q, E, S) -> (E, L, S), the layout SOBER/_rchq.py:124 expects from a kernel
        left = (self.w_qd.to(mu_x).unsqueeze(1) * (mu_x - Ex.unsqueeze(0))).T
        return W * torch.einsum("lq,qes->els", left, mu_y - Ey.unsqueeze(0))
