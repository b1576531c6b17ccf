"""sug_b200 — B200-native (sm_100a) implementation of the SUG point-cloud encoder hot path:
DGCNN kNN / EdgeConv, PointNet shared-MLP + max-pool, and the multi-kernel Gaussian MMD, behind
the reference's own Python API.  See DESIGN.md and INTEGRATION.md."""
from . import _lib  # noqa: F401

__all__ = ["ops", "model_utils", "point_utils", "Model", "model_pointnet", "mmd", "compat"]
