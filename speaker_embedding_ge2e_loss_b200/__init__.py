"""B200-native GE2E loss: drop-in for gkv856/speaker_embedding_GE2E_loss's ``GE2ELoss``.

Host layer only (module, torch.library op, ctypes binding of include/ge2e_b200.h); all
arithmetic happens in the hand-written sm_100a kernels under ``csrc/``.
"""
from ._lib import GE2ELibraryError, lib  # noqa: F401
from .batches import SpectrogramBank  # noqa: F401
from .evaluation import EERResult, eer_sweep, evaluate_eer, threshold_counts  # noqa: F401
from .loss import GE2ELoss  # noqa: F401
from .ops import ge2e_loss  # noqa: F401
from .plan import GE2EHostFeed, GE2EPlan, ShardedGE2EHostFeed, ShardedGE2EPlan  # noqa: F401
from .tail import ProjectionL2Norm, project_normalize  # noqa: F401
from .sharded import shard_bounds, sharded_ge2e_loss  # noqa: F401

__all__ = ["GE2ELoss", "GE2EPlan", "GE2EHostFeed", "ShardedGE2EHostFeed", "ShardedGE2EPlan", "ge2e_loss", "eer_sweep", "evaluate_eer", "threshold_counts", "EERResult", "SpectrogramBank", "ProjectionL2Norm", "project_normalize", "sharded_ge2e_loss", "shard_bounds", "GE2ELibraryError", "lib"]
