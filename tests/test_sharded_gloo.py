"""N > 1 host logic: candidates row-sharded over a world_size-2 gloo group on the CPU (TorchOps test double for
the kernels).  The sharded run must select the same points as the single-process run."""
import os
import warnings

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from _cases import Case


def _worker(rank, world, port, name, split, out):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from _cpu_ops import TorchOps
    from sober_b200 import Recombiner, Sharded, configure
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        case = Case(name)
        lo, hi = (0, split) if rank == 0 else (split, len(case.X))
        mu = None if case.mu is None else case.mu[lo:hi].clone()
        with warnings.catch_warnings(), configure(mode="parity") as opts:
            warnings.simplefilter("ignore")
            torch.manual_seed(7)
            idx, w = Recombiner(TorchOps(), comm=Sharded(), opts=opts).run(
                case.X[lo:hi].clone(), case.Z, case.b, case.kernel(), init_weights=mu, calc_obj=case.objective)
        out[rank] = (idx.clone(), w.clone(), None if mu is None else mu.clone())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,split", [("matern6d_rest", 1234), ("predcov_matern6d", 3000),
                                        ("direct_branch", 17), ("tanimoto256", 5),
                                        # calc_obj branch (SOBER/_rchq.py:67-69,138-150,169-196) with sharded candidates:
                                        # the per-group objective sums ride in the packed all-reduce; split = 100 puts
                                        # the position-indexed objective lookup of :89 across both shards
                                        ("objective_matern4d", 2000), ("objective_matern4d", 100),
                                        # a rank whose weights are all zero (no alive rows on a shard)
                                        ("matern6d_rest", 0)])
def test_two_rank_shard_equals_single_process(name, split):
    port = 29500 + (os.getpid() + hash(name)) % 2000
    manager = mp.Manager()
    out = manager.dict()
    mp.spawn(_worker, args=(2, port, name, split, out), nprocs=2, join=True)
    case = Case(name)
    (i0, w0, m0), (i1, w1, m1) = out[0], out[1]
    assert torch.equal(i0, i1) and torch.equal(w0, w1)          # replicated result
    assert torch.equal(i0, case.idx)                             # == reference fixture
    assert float((w0 - case.w).abs().max()) < 1e-9
    if m0 is not None:                                           # each rank's weight shard holds its part
        merged = torch.cat([m0, m1])
        assert float((merged - torch.from_numpy(case.raw["mu_after"])).abs().max()) < 1e-9


# ------------------------------------------------------------------------------------------------------------
# k-means landmark selection: rows sharded, one all-reduce of sums and counts per iteration
# ------------------------------------------------------------------------------------------------------------
def _kmeans_worker(rank, world, port, split, out):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from _cpu_ops import TorchOps
    from sober_b200 import Sharded
    from sober_b200._kmeans import kmeans
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        x = torch.rand(6000, 4, dtype=torch.float64, generator=torch.Generator().manual_seed(11))
        lo, hi = (0, split) if rank == 0 else (split, len(x))
        cl, c = kmeans(x[lo:hi].clone(), K=40, Niter=6, ops=TorchOps(), comm=Sharded())
        out[rank] = (cl.clone(), c.clone())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("split", [3000, 17, 0])
def test_two_rank_kmeans_equals_single_process(split):
    """split = 17: the first K = 40 rows (the initial centroids) straddle the two ranks; split = 0: an empty shard."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from _cpu_ops import TorchOps
    from sober_b200._kmeans import kmeans
    port = 31500 + (os.getpid() + split) % 2000
    manager = mp.Manager()
    out = manager.dict()
    mp.spawn(_kmeans_worker, args=(2, port, split, out), nprocs=2, join=True)
    x = torch.rand(6000, 4, dtype=torch.float64, generator=torch.Generator().manual_seed(11))
    cl, c = kmeans(x, K=40, Niter=6, ops=TorchOps())
    (cl0, c0), (cl1, c1) = out[0], out[1]
    assert torch.equal(c0, c1)
    assert float((c0 - c).abs().max()) < 1e-12
    assert torch.equal(torch.cat([cl0, cl1]), cl)


# ------------------------------------------------------------------------------------------------------------
def _projector_worker(rank, world, port, S, dim, out):
    from sober_b200 import Sharded, _car
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        g = torch.Generator().manual_seed(S + dim)
        a = torch.randn(S, dim, dtype=torch.float64, generator=g)
        a[:, 0] = 1.0
        scaled = a / a.norm(dim=0, keepdim=True)
        rows, delta = _car.projector_rows_sharded(Sharded(), scaled)
        out[rank] = (rows.clone(), delta.clone())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("S,dim,world", [(90, 31, 2), (64, 33, 3), (50, 49, 2)])
def test_sharded_projector_null_space(S, dim, world):
    """``_car.projector_rows_sharded`` (the multi-GPU split of the replicated Caratheodory step's null space): every
    rank ends with the SAME bits, the rows span the null space of the design, and they equal the single-rank formulas."""
    from sober_b200 import _car
    from sober_b200._rchq import SingleProcess
    port = 29500 + (os.getpid() + 7 * S + dim) % 2000
    manager = mp.Manager()
    out = manager.dict()
    mp.spawn(_projector_worker, args=(world, port, S, dim, out), nprocs=world, join=True)
    g = torch.Generator().manual_seed(S + dim)
    a = torch.randn(S, dim, dtype=torch.float64, generator=g)
    a[:, 0] = 1.0
    scaled = a / a.norm(dim=0, keepdim=True)
    rows1, delta1 = _car.projector_rows_sharded(SingleProcess(), scaled)
    for r in range(1, world):
        assert torch.equal(out[0][0], out[r][0]) and torch.equal(out[0][1], out[r][1])
    rows, delta = out[0]
    assert rows.shape == (S - dim, S)
    assert float((rows - rows1).abs().max()) < 1e-12 and float((delta - delta1).abs().max()) < 1e-12
    assert float((rows @ scaled).abs().max()) < 1e-12                   # null space of the design
    # the trailing columns of the orthogonal projector I - Q Q^T
    q, _ = torch.linalg.qr(scaled)
    want = (torch.eye(S, dtype=torch.float64) - q @ q.T)[dim:, :]
    assert float((rows - want).abs().max()) < 1e-11
