"""B200-native (sm_100a) training + retrieval hot path of the two-tower recommender with the adaptive
mimic mechanism.  Drop-in for the reference's `src.models` API (see models.py) plus the fused step /
evaluation hooks for `src.pipelines.training` (see hooks.py).  Everything numeric runs in libttam.so
(include/ttam.h); there is no CPU fallback."""
from . import _lib
from ._lib import TtamError, build, lib  # noqa: F401
from .models import (  # noqa: F401
    AdaptiveMimicMechanism,
    FeatureFusionGate,
    TowerEncoder,
    TwoTowerModel,
    build_feature_encoder,
    build_id_embedding,
    build_tower_encoder,
)
from .engine import FusedEngine  # noqa: F401
from .sharded import ShardedEngine, ShardedFlatIPIndex, build_sharded_engine  # noqa: F401
from . import functional, hooks, retrieval, sharding  # noqa: F401

__all__ = ["AdaptiveMimicMechanism", "FeatureFusionGate", "TowerEncoder", "TwoTowerModel", "build_feature_encoder",
           "build_id_embedding", "build_tower_encoder", "FusedEngine", "TtamError", "build", "lib", "functional",
           "hooks", "retrieval", "sharding", "ShardedEngine", "ShardedFlatIPIndex", "build_sharded_engine"]
