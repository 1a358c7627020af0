#!/usr/bin/env python
"""Golden fixtures for k-means: runs the UNMODIFIED ``KMeans`` of the reference's ``SOBER/_weights.py`` (loaded by file
path: it imports nothing but torch).  Build container only.

    python tests/golden/make_golden_kmeans.py"""
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SOBER_REFERENCE", "/root/reference")


def load_reference_kmeans(root=REF):
    spec = importlib.util.spec_from_file_location("_ref_weights", os.path.join(root, "SOBER", "_weights.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.KMeans


def make(name, seed, n, d, k, niter, clustered):
    KMeans = load_reference_kmeans()
    g = torch.Generator().manual_seed(seed)
    if clustered:      # mixture of tight blobs (+ duplicates of the first rows: an EMPTY cluster -> NaN centroid)
        centres = torch.rand(7, d, dtype=torch.float64, generator=g)
        x = centres[torch.randint(0, 7, (n,), generator=g)] + 0.02 * torch.randn(n, d, dtype=torch.float64, generator=g)
        x[1] = x[0]
    else:
        x = torch.rand(n, d, dtype=torch.float64, generator=g)
    cl, c = KMeans(x.clone(), k, niter)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), x=x.numpy(), K=k, Niter=niter, cl=cl.numpy(), c=c.numpy())
    print(name, "N", n, "D", d, "K", k, "NaN centroids", int(torch.isnan(c).any(1).sum()),
          "largest cluster", int(torch.bincount(cl, minlength=k).max()))


if __name__ == "__main__":
    make("kmeans_6d_uniform", 1, 20000, 6, 100, 10, False)
    make("kmeans_2d_uniform", 2, 5000, 2, 37, 10, False)
    make("kmeans_3d_blobs_empty_cluster", 3, 3000, 3, 16, 10, True)
    make("kmeans_12d_uniform", 4, 4000, 12, 50, 5, False)
