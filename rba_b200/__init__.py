"""rba_b200 — B200-native (sm_100a) inference path for RbA: Mask2Former forward + Rejected-by-All score.

Everything compute runs in the in-tree CUDA library (rba_b200/lib/librba_b200.so, C ABI in
include/rba_b200.h).  Importing this package does not need a GPU; calling into it does, and fails loudly
otherwise — there is no CPU fallback."""
from . import config, ops  # noqa: F401
from ._lib import LIB_PATH, RbaError, launch_count  # noqa: F401
from .config import ModelConfig, model_config_from_cfg, model_config_from_yaml  # noqa: F401
from .engine import Engine  # noqa: F401
from .metrics import OODEvaluator, StreamingOODMetrics, evaluate_ood  # noqa: F401
from .modeling import MaskFormer, build_model  # noqa: F401
from .pipeline import PinnedBatcher, ScoreStream  # noqa: F401

__version__ = "0.1.0"
