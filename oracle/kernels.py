"""gpytorch-free kernel objects for the oracle and the tests (TEST INFRASTRUCTURE, see oracle/__init__.py).

The reference never evaluates a kernel itself: ``SOBER/_rchq.py:35,78,124,131,156`` call an opaque
``kernel(x, y)`` callable which in the examples is ``SOBER._kernel.Kernel`` wrapping a gpytorch model
(``SOBER/_kernel.py:16-30`` -> ``SOBER/_gp.py:281-295`` -> ``model.covar_module.forward``).  gpytorch is an
un-vendored dependency (``requirements.txt:2`` ``gpytorch==1.10``; ``pyproject.toml:27`` ``>=1.11``) that is not
installed in this image, so the arithmetic of ``ScaleKernel``, ``RBFKernel``, ``MaternKernel`` and of
``gpytorch.kernels.kernel.sq_dist / dist`` is RESTATED here from the library's published algorithm:

* ``sq_dist``: subtract ``x1.mean(-2)`` from both inputs, form ``[-2 x1, |x1|^2, 1] @ [x2, 1, |x2|^2]^T``,
  zero the diagonal when ``x1`` equals ``x2``, ``clamp_min(0)``;  ``dist = sqrt(clamp_min(sq_dist, 1e-30))``.
* RBF: ``exp(-sq_dist(x1/l, x2/l) / 2)``.
* Matern: centre both inputs by the mean of ``x1``, divide by ``l``, ``r = dist``;
  nu=1/2: ``exp(-r)``; nu=3/2: ``(1+sqrt3 r) exp(-sqrt3 r)``; nu=5/2: ``(1+sqrt5 r+5/3 r^2) exp(-sqrt5 r)``.
* Scale: multiply by ``outputscale``.

PARITY UNPINNED for this file's gpytorch part: there is no gpytorch here to generate fixtures from.  The
Tanimoto similarity restates ``SOBER/_drug_modelling.py:15-25,36-38`` which IS in the reference and is pure
torch; ``tests/golden/make_golden.py`` pins it against the reference source when ``/root/reference`` exists.

The classes deliberately expose the same attribute surface the product's introspector reads from real
gpytorch objects (class names ``ScaleKernel`` / ``RBFKernel`` / ``MaternKernel`` / ``TanimotoKernel``,
``.base_kernel``, ``.lengthscale`` of shape (1, 1) or (1, d), ``.outputscale``, ``.nu``, ``.forward``), and
``GPModel`` exposes ``covar_module``, ``train_inputs`` and a prediction-strategy ``covar_cache`` the way
``SOBER/_gp.py:255-278`` consumes them.
"""
import math

import torch


# --------------------------------------------------------------------------------------------------
# distances (gpytorch.kernels.kernel.sq_dist / dist, restated)
# --------------------------------------------------------------------------------------------------
def _same_points(a, b):
    return a.shape == b.shape and bool(torch.equal(a, b))


def expanded_sq_dist(a, b, same):
    shift = a.mean(-2, keepdim=True)
    a = a - shift
    a_sq = a.pow(2).sum(dim=-1, keepdim=True)
    one_a = torch.ones_like(a_sq)
    if same:
        b, b_sq, one_b = a, a_sq, one_a
    else:
        b = b - shift
        b_sq = b.pow(2).sum(dim=-1, keepdim=True)
        one_b = torch.ones_like(b_sq)
    left = torch.cat([-2.0 * a, a_sq, one_a], dim=-1)
    right = torch.cat([b, one_b, b_sq], dim=-1)
    out = left.matmul(right.transpose(-2, -1))
    if same:
        out.diagonal(dim1=-2, dim2=-1).fill_(0)
    return out.clamp_min_(0)


def expanded_dist(a, b, same):
    return expanded_sq_dist(a, b, same).clamp_min_(1e-30).sqrt_()


# --------------------------------------------------------------------------------------------------
# kernel objects (duck-typed like gpytorch's)
# --------------------------------------------------------------------------------------------------
class _Base:
    has_lengthscale = True

    def __init__(self, lengthscale=None):
        if lengthscale is not None:
            ls = torch.as_tensor(lengthscale, dtype=torch.float64).reshape(1, -1)
            self.lengthscale = ls

    def to(self, device):
        if getattr(self, "lengthscale", None) is not None:
            self.lengthscale = self.lengthscale.to(device)
        return self

    def __call__(self, x1, x2):
        return self.forward(x1, x2)


class RBFKernel(_Base):
    def forward(self, x1, x2, **_):
        ls = self.lengthscale.to(x1)
        a, b = x1.div(ls), x2.div(ls)
        return expanded_sq_dist(a, b, _same_points(a, b)).div_(-2).exp_()


class MaternKernel(_Base):
    def __init__(self, nu=2.5, lengthscale=None):
        super().__init__(lengthscale)
        if nu not in (0.5, 1.5, 2.5):
            raise RuntimeError("nu expected to be 0.5, 1.5, or 2.5")
        self.nu = nu

    def forward(self, x1, x2, **_):
        ls = self.lengthscale.to(x1)
        centre = x1.reshape(-1, x1.size(-1)).mean(0)[(None,) * (x1.dim() - 1)]
        a, b = (x1 - centre).div(ls), (x2 - centre).div(ls)
        r = expanded_dist(a, b, _same_points(a, b))
        decay = torch.exp(-math.sqrt(self.nu * 2) * r)
        if self.nu == 0.5:
            poly = 1
        elif self.nu == 1.5:
            poly = (math.sqrt(3) * r).add(1)
        else:
            poly = (math.sqrt(5) * r).add(1).add(5.0 / 3.0 * r ** 2)
        return poly * decay


class TanimotoKernel(_Base):
    """SOBER/_drug_modelling.py:15-25 (similarity) and :36-38 (clamp at zero)."""
    has_lengthscale = False
    eps = 1e-6

    def __init__(self):
        super().__init__(None)

    def forward(self, x1, x2, **_):
        cross = torch.matmul(x1, torch.transpose(x2, -1, -2))
        n1 = torch.sum(x1 ** 2, dim=-1, keepdims=True)
        n2 = torch.sum(x2 ** 2, dim=-1, keepdims=True)
        sim = (cross + self.eps) / (self.eps + n1 + torch.transpose(n2, -1, -2) - cross)
        sim.clamp_min_(0)
        return sim


class ScaleKernel(_Base):
    has_lengthscale = False

    def __init__(self, base_kernel, outputscale=1.0):
        super().__init__(None)
        self.base_kernel = base_kernel
        self.outputscale = torch.as_tensor(outputscale, dtype=torch.float64)

    def to(self, device):
        self.base_kernel.to(device)
        self.outputscale = self.outputscale.to(device)
        return self

    def forward(self, x1, x2, **_):
        inner = self.base_kernel.forward(x1, x2)
        s = self.outputscale.to(inner)
        return inner.mul(s.view(*s.shape, 1, 1))


# --------------------------------------------------------------------------------------------------
# a GP-model stand-in carrying what SOBER/_gp.py:255-295 reads
# --------------------------------------------------------------------------------------------------
class _Strategy:
    def __init__(self, covar_cache, mean_cache=None):
        self.covar_cache = covar_cache
        self.mean_cache = mean_cache


class _Dist:
    """What ``model(x)`` / ``likelihood(model(x))`` return as far as SOBER/_pi.py:15 looks: ``.loc``, ``.mean``,
    ``.variance``."""

    def __init__(self, loc, variance):
        self.loc = self.mean = loc
        self.variance = variance


class _Noise:
    def __init__(self, noise):
        self.noise = noise

    def __call__(self, dist):                      # Gaussian likelihood: adds the observation noise
        return _Dist(dist.loc, dist.variance + self.noise)

    def eval(self):
        return self


class _ConstMean:
    def __init__(self, constant):
        self.constant = constant


class GPModel:
    """Minimal exact-GP stand-in: ``covar_module``, ``train_inputs``, ``prediction_strategy.covar_cache``.

    ``covar_cache`` is a root ``S`` with ``S S^T = (K_obs + noise I)^-1`` -- what gpytorch's exact prediction
    strategy exposes and ``get_cov_cache`` (``SOBER/_gp.py:255-278``) multiplies out.
    """

    def __init__(self, covar_module, train_x, train_y=None, noise=1e-4, mean_constant=0.0):
        self.covar_module = covar_module
        self.train_inputs = (train_x,)
        self.train_targets = train_y
        nz = torch.as_tensor([noise], dtype=train_x.dtype, device=train_x.device)
        self.likelihood = _Noise(nz)
        self.mean_module = _ConstMean(torch.as_tensor(mean_constant, dtype=train_x.dtype, device=train_x.device))
        k_obs = covar_module.forward(train_x, train_x)
        k_obs = 0.5 * (k_obs + k_obs.T) + noise * torch.eye(len(train_x), dtype=train_x.dtype, device=train_x.device)
        chol = torch.linalg.cholesky(k_obs)
        eye = torch.eye(len(train_x), dtype=train_x.dtype, device=train_x.device)
        root = torch.linalg.solve_triangular(chol, eye, upper=False).T  # root @ root.T = k_obs^-1
        mean_cache = None
        if train_y is not None:
            mean_cache = torch.cholesky_solve((train_y - self.mean_module.constant).reshape(-1, 1), chol).reshape(-1)
        self.prediction_strategy = _Strategy(root, mean_cache)

    def eval(self):
        return self

    def __call__(self, x):
        """Exact posterior at ``x`` (mean ``c + k alpha``, variance ``k(x,x) - k W k``, no observation noise)."""
        k_xo = self.covar_module.forward(x, self.train_inputs[0])
        root = self.prediction_strategy.covar_cache
        mean = self.mean_module.constant + k_xo @ self.prediction_strategy.mean_cache
        var = torch.diagonal(self.covar_module.forward(x, x)) - ((k_xo @ root) ** 2).sum(-1)
        return _Dist(mean, var)


def covariance_cache(model):
    """``get_cov_cache`` of SOBER/_gp.py:255-278 (the try-branch: the cache exists)."""
    x_obs = model.train_inputs[0]
    root = model.prediction_strategy.covar_cache
    return root @ root.T, x_obs, model.likelihood.noise


def predictive_covariance(x, y, model):
    """SOBER/_gp.py:281-295:  k(x,y) - k(x,X) W k(X,y)."""
    w, x_obs, _ = covariance_cache(model)
    k_xy = model.covar_module.forward(x, y)
    k_xo = model.covar_module.forward(x, x_obs)
    k_oy = model.covar_module.forward(x_obs, y)
    return k_xy - k_xo @ w @ k_oy


def predictive_mean(x, model):
    """Posterior mean ``c + k(x, X) alpha`` -- what ``predict_mean`` (SOBER/_gp.py:240-253) returns."""
    x_obs = model.train_inputs[0]
    return model.mean_module.constant + model.covar_module.forward(x, x_obs) @ model.prediction_strategy.mean_cache


class Kernel:
    """Mirror of ``SOBER._kernel.Kernel`` (SOBER/_kernel.py:4-47): same constructor, same three modes."""

    def __init__(self, model, mode="predictive_covariance"):
        self.model = model
        self.mode = mode

    def __call__(self, x, y):
        if self.mode == "predictive_covariance":
            return predictive_covariance(x, y, self.model)
        if self.mode == "weighted_predictive_covariance":
            m_x = predictive_mean(x, self.model)
            m_y = predictive_mean(y, self.model)
            cov = predictive_covariance(x, y, self.model)
            if m_x.dim() == 1 and m_y.dim() == 1:
                return m_x.unsqueeze(1) * cov * m_y.unsqueeze(0)
            return m_x.unsqueeze(1) * cov * m_y.unsqueeze(1)
        if self.mode == "kernel":
            return self.model.covar_module.forward(x, y)
        raise ValueError(
            'mode should be from ["predictive_covariance", "weighted_predictive_covariance", "kernel"]')


# convenience constructors used by tests / bench -------------------------------------------------
def make_kernel(family, lengthscale=1.0, outputscale=1.0, nu=2.5):
    if family == "rbf":
        base = RBFKernel(lengthscale)
    elif family == "matern":
        base = MaternKernel(nu, lengthscale)
    elif family == "tanimoto":
        base = TanimotoKernel()
    else:
        raise ValueError(family)
    return ScaleKernel(base, outputscale)


class BareModel:
    """Model stand-in for ``Kernel(model, mode="kernel")`` when there are no observations."""

    def __init__(self, covar_module):
        self.covar_module = covar_module
