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
                case.X[lo:hi].clone(), case.Z, case.b, case.kernel(), init_weights=mu)
        out[rank] = (idx.clone(), w.clone(), None if mu is None else mu.clone())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,split", [("matern6d_rest", 1234), ("predcov_matern6d", 3000),
                                        ("direct_branch", 17), ("tanimoto256", 5)])
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
