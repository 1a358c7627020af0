"""Row-sharded recombination over NCCL on >= 2 GPUs of one box: same points as the single-GPU run.
Skipped on a single-GPU box (the CPU suite covers the same host logic over gloo)."""
import os
import warnings

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, name, out):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch.distributed as dist
    from _cases import Case
    import sober_b200
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        case = Case(name, dev)
        n = len(case.X)
        cut = [0] + [int(n * (r + 1) / world) + (7 if r + 1 < world else 0) for r in range(world)]
        cut[-1] = n
        lo, hi = cut[rank], cut[rank + 1]
        mu = None if case.mu is None else case.mu[lo:hi].clone()
        sober_b200.enable_sharding()
        with warnings.catch_warnings(), sober_b200.configure(mode="parity"):
            warnings.simplefilter("ignore")
            rec = sober_b200.Recombiner(sober_b200._rchq._ops(), comm=sober_b200.Sharded(), basis=case.U)
            idx, w = rec.run(case.X[lo:hi].contiguous(), case.Z, case.b, case.kernel(), init_weights=mu)
        out[rank] = (idx.cpu(), w.cpu(), None if mu is None else mu.cpu(), lo, hi)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["matern6d_rest", "predcov_matern6d", "tanimoto256"])
def test_nccl_shards_equal_single_gpu(name):
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(world, 4)
    import torch.multiprocessing as mp
    from _cases import Case
    import sober_b200
    # single-GPU run with the same (fixture) Nystrom basis; null space from this device's cuSOLVER SVD in both runs
    case = Case(name, torch.device("cuda", 0))
    mu1 = None if case.mu is None else case.mu.clone()
    with warnings.catch_warnings(), sober_b200.configure(mode="parity"):
        warnings.simplefilter("ignore")
        rec = sober_b200.Recombiner(sober_b200._rchq._ops(), basis=case.U)
        idx1, w1 = rec.run(case.X, case.Z, case.b, case.kernel(), init_weights=mu1)
    manager = mp.Manager()
    out = manager.dict()
    mp.spawn(_worker, args=(world, 29650 + os.getpid() % 300, name, out), nprocs=world, join=True)
    for r in range(world):
        idx, w, mu, lo, hi = out[r]
        assert torch.equal(idx, idx1.cpu())
        assert float((w - w1.cpu()).abs().max()) < 1e-9
        if mu is not None:
            assert float((mu - mu1.cpu()[lo:hi]).abs().max()) < 1e-9


def _fast_worker(rank, world, port, name, out):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch.distributed as dist
    from _cases import Case
    import sober_b200
    from sober_b200 import _nystrom
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        case = Case(name, dev)
        n = len(case.X)
        cut = [int(n * r / world) for r in range(world)] + [n]
        lo, hi = cut[rank], cut[rank + 1]
        mu = None if case.mu is None else case.mu[lo:hi].clone()
        sober_b200.enable_sharding()
        _nystrom._injected_test_matrix = torch.randn(case.Z.shape[0], case.b - 1, dtype=torch.float64,
                                                     generator=torch.Generator().manual_seed(5)).to(dev)
        # car_shard_min = 0: the projector null space of every Caratheodory call is split over the ranks
        with warnings.catch_warnings(), sober_b200.configure(mode="fast", car_shard_min=0):
            warnings.simplefilter("ignore")
            idx, w = sober_b200.recombination(case.X[lo:hi].contiguous(), case.Z, case.b, case.kernel(), None, None,
                                              init_weights=mu)
        out[rank] = (idx.cpu(), w.cpu())
    finally:
        _nystrom._injected_test_matrix = None
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["matern6d_rest", "rbf_ard5d"])
def test_nccl_sharded_projector_equals_single_gpu_fast_mode(name):
    """Fast mode over NCCL with the null space of the replicated Caratheodory step split over the ranks
    (_car.projector_rows_sharded): every rank returns the single-GPU selection."""
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(world, 4)
    import torch.multiprocessing as mp
    from _cases import Case
    import sober_b200
    from sober_b200 import _nystrom
    case = Case(name, torch.device("cuda", 0))
    mu1 = None if case.mu is None else case.mu.clone()
    _nystrom._injected_test_matrix = torch.randn(case.Z.shape[0], case.b - 1, dtype=torch.float64,
                                                 generator=torch.Generator().manual_seed(5)).to(case.X.device)
    try:
        with warnings.catch_warnings(), sober_b200.configure(mode="fast"):
            warnings.simplefilter("ignore")
            idx1, w1 = sober_b200.recombination(case.X, case.Z, case.b, case.kernel(), None, None, init_weights=mu1)
    finally:
        _nystrom._injected_test_matrix = None
    manager = mp.Manager()
    out = manager.dict()
    mp.spawn(_fast_worker, args=(world, 29350 + os.getpid() % 300, name, out), nprocs=world, join=True)
    for r in range(world):
        idx, w = out[r]
        assert torch.equal(idx, idx1.cpu())
        assert float((w - w1.cpu()).abs().max()) < 1e-9
