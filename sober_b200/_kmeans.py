"""Lloyd's k-means as the reference runs it to pick the Nystrom landmarks -- SURVEY.md §8(f) row 2: ``KMeans`` of
``SOBER/_weights.py:100-126`` behind ``kmeans_resampling`` (``:95-97``), called by ``sampling_recombination`` for
continuous domains (``SOBER/_sampler.py:316-317``) to produce the ``pts_nys`` that ``recombination`` receives.

Same algorithm, same outputs ``(cl, c)``: centroids initialised with the first K points, ``Niter`` rounds of
[assign every point to its nearest centroid (first minimum on ties) | centroid = mean of its points], empty clusters
becoming NaN centroids exactly like ``c /= Ncl`` does there.  The reference's E step materialises the (N, K, D)
difference tensor -- 48 GB at the benchmark's N = 1e6, K = 1000, D = 6, which is why it cannot produce landmarks at
that scale; here it is ``sober_kmeans_assign`` (``csrc/kmeans.cu``): one point per thread in registers, centroids
streamed through shared memory, only the labels written.  The M step is two torch calls on the device (``index_add_``
and ``bincount``: N x D atomics, their summation order is not reproducible to the last bit, like on the reference's
own CUDA path).
"""
import torch

_MAX_D = 16


def kmeans(x, K=10, Niter=10, ops=None):
    """-> (cl (N,) int64 labels of the LAST assignment, c (K, D) centroids after the last update), float64, on the device."""
    if ops is None:
        from ._rchq import _ops
        ops = _ops()
    x = ops.f64(x)
    if x.dim() != 2:
        raise ValueError("x must be (N, D)")
    N, D = x.shape
    if N < K:
        raise ValueError("k-means needs at least K points (the reference's c.view(1, K, D) fails likewise)")
    c = x[:K, :].clone()                                   # SOBER/_weights.py:103
    cl = torch.zeros(N, dtype=torch.int64, device=x.device)
    for _ in range(Niter):
        if D <= _MAX_D and hasattr(ops, "kmeans_assign"):
            cl = ops.kmeans_assign(x, c)                   # E step, :113-115
        else:                                              # wide rows: chunked torch (never (N, K, D) at once)
            cl = torch.cat([((x[s:s + 8192, None, :] - c[None]) ** 2).sum(-1).argmin(1) for s in range(0, N, 8192)])
        c = torch.zeros_like(c)
        c.index_add_(0, cl, x)                             # :119-120
        c /= torch.bincount(cl, minlength=K).to(c.dtype).view(K, 1)    # :123-124
    return cl, c
