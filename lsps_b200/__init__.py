"""lsps_b200 -- B200-native (sm_100a) training step of masabdi/LSPS behind the reference's trainer API.

Class names are importable with `from lsps_b200 import *` so that the reference's name-based selection
(`exec("trainer=%s(config.hyperparameters)")`, src/depth_train.py:99-102) picks them straight from the YAML.
The compute path is the in-tree CUDA library (csrc/liblsps_b200.so); importing fails loudly if it cannot be loaded.
"""
from . import _lib  # noqa: F401  (loads / builds the CUDA library; raises ImportError when impossible)
from .trainer import LSPSTrainerB200
from .nets import SharedResGenB200, SharedResXGenB200, SharedDisB200, poseVAEB200, MappingB200
from .data import SyntheticHandDataset, synthetic_batch
from .config import NetConfig, load_hyperparameters
from .evaluation import PoseEvaluator, NYU_RESTRICTED_JOINTS
from .augment import CropAugmenter

LSPSTrainer = LSPSTrainerB200  # the reference's own name selects the B200 trainer as well

__all__ = ["CropAugmenter", "LSPSTrainerB200", "LSPSTrainer", "SharedResGenB200", "SharedResXGenB200", "SharedDisB200",
           "poseVAEB200", "MappingB200",
           "SyntheticHandDataset", "synthetic_batch", "NetConfig", "load_hyperparameters", "PoseEvaluator",
           "NYU_RESTRICTED_JOINTS"]
