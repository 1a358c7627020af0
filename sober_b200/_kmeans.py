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


def kmeans(x, K=10, Niter=10, ops=None, comm=None):
    """-> (cl (N,) int64 labels of the LAST assignment, c (K, D) centroids after the last update), float64, on the device.

    ``comm`` (a ``Sharded`` communicator): ``x`` is this rank's contiguous block of the rows (rank order = row order).
    The E step is local; one all-reduce of the (K, D) sums and (K,) counts per iteration; every rank gets the same
    centroids and the labels of its own rows."""
    if ops is None:
        from ._rchq import _ops
        ops = _ops()
    x = ops.f64(x)
    if x.dim() != 2:
        raise ValueError("x must be (N, D)")
    N, D = x.shape
    if comm is not None and comm.world > 1:
        return _kmeans_sharded(x, K, Niter, ops, comm)
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


def _assign(ops, x, c):
    if x.shape[1] <= _MAX_D and hasattr(ops, "kmeans_assign"):
        return ops.kmeans_assign(x, c)
    return torch.cat([((x[s:s + 8192, None, :] - c[None]) ** 2).sum(-1).argmin(1) for s in range(0, x.shape[0], 8192)]
                     + [torch.zeros(0, dtype=torch.int64, device=x.device)])


def _kmeans_sharded(x, K, Niter, ops, comm):
    N, D = x.shape
    counts = comm.all_gather_ints(N, x.device)
    row0 = sum(counts[:comm.rank])
    if sum(counts) < K:
        raise ValueError("k-means needs at least K points (the reference's c.view(1, K, D) fails likewise)")
    # initial centroids = the first K rows of the GLOBAL order: every rank contributes the ones it holds
    c = torch.zeros((K, D), dtype=torch.float64, device=x.device)
    lo, hi = max(row0, 0), min(row0 + N, K)
    if hi > lo:
        c[lo:hi] = x[lo - row0:hi - row0]
    c = comm.all_reduce(c)
    cl = torch.zeros(N, dtype=torch.int64, device=x.device)
    for _ in range(Niter):
        cl = _assign(ops, x, c) if N > 0 else cl
        packed = torch.zeros((K, D + 1), dtype=torch.float64, device=x.device)
        if N > 0:
            packed[:, :D].index_add_(0, cl, x)
            packed[:, D] = torch.bincount(cl, minlength=K).to(torch.float64)
        packed = comm.all_reduce(packed)
        c = packed[:, :D] / packed[:, D:D + 1]
    return cl, c.contiguous()
