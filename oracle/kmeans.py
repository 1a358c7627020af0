"""TEST INFRASTRUCTURE ONLY -- CPU restatement of ``KMeans`` (``SOBER/_weights.py:100-126``), the SURVEY.md §8(f) row
"Nystrom-point selection".  Same torch ops in the same order, except that the (N, K, D) broadcast of the E step is
evaluated in chunks of points (each row of D_ij is computed exactly as there, so labels and centroids are
bit-identical; ``tests/test_kmeans.py`` checks that against fixtures produced by the unmodified reference function,
``tests/golden/make_golden_kmeans.py``)."""
import torch


def kmeans(x, K=10, Niter=10, chunk=4096):
    N, D = x.shape
    c = x[:K, :].clone()                                                    # :103
    c_j = c.view(1, K, D)                                                   # :106 (a view: follows the in-place updates)
    cl = None
    for _ in range(Niter):
        parts = []
        for s in range(0, N, chunk):
            x_i = x[s:s + chunk].reshape(-1, 1, D)
            D_ij = ((x_i - c_j) ** 2).sum(-1)                               # :114
            parts.append(D_ij.argmin(dim=1).long().view(-1))                # :115
        cl = torch.cat(parts)
        c.zero_()                                                           # :119
        c.scatter_add_(0, cl[:, None].repeat(1, D), x)                      # :120
        Ncl = torch.bincount(cl, minlength=K).type_as(c).view(K, 1)         # :123
        c /= Ncl                                                            # :124
    return cl, c
