"""sober_b200 -- B200-native (sm_100a) implementation of SOBER's RCHQ batch-selection hot path.

    from sober_b200 import recombination          # same signature / contract as SOBER._rchq.recombination
    import sober_b200; sober_b200.install()        # rebind it inside an imported SOBER package

Importing this package does not need a GPU; calling ``recombination`` does (there is no CPU fallback).
"""
from ._rchq import Recombiner, Sharded, SingleProcess, recombination, set_communicator
from ._settings import configure, options
from ._install import install, uninstall
from ._wkde import wkde_pdf
from ._kmeans import kmeans
from ._predict import gp_posterior, pi_lfi, predict

__all__ = ["recombination", "install", "uninstall", "configure", "options", "Recombiner", "Sharded",
           "SingleProcess", "set_communicator", "enable_sharding", "wkde_pdf", "kmeans", "gp_posterior", "pi_lfi", "predict"]
__version__ = "0.1.0"


def enable_sharding(group=None):
    """Row-shard the candidates over an initialised ``torch.distributed`` group: afterwards ``recombination``
    expects this rank's contiguous block of ``pts_rec`` / ``init_weights`` rows (rank order = row order) and
    returns GLOBAL indices on every rank."""
    set_communicator(Sharded(group))
