"""lmp2sqt_b200 -- B200-native space-time correlation path of infinitekh/lmp2sqt.

Only what the hot path needs: ``csrc/`` (sm_100a kernels + C-ABI + host C code), the ctypes
bindings, a Python mirror of the reference's call seam, the synthetic-dump generator and the
``torch.distributed`` plumbing for the particle-sharded multi-GPU path.
"""
from .engine import (Deck, MultiSpacetimeCorr, Results, SpacetimeCorr, SqtError, SynthSource, comm_unique_id,  # noqa: F401
                     measure_dfma_peak, run_trajectories, shard_bounds)
from ._lib import SQT_FLUSH, SQT_OK  # noqa: F401
from .analysis import Analysis  # noqa: F401

__all__ = ["Analysis", "Deck", "MultiSpacetimeCorr", "Results", "SpacetimeCorr", "SqtError", "SynthSource", "comm_unique_id", "measure_dfma_peak",
           "run_trajectories", "shard_bounds", "SQT_FLUSH", "SQT_OK"]
