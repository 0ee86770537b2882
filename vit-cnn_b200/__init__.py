"""vitcnn_b200: B200-native implementation of the ViT-CNN hybrid's data-parallel hot path
(per-pixel HSI + LiDAR patch extraction and the hybrid's forward / backward) behind the
reference toolkit's own get_model('ViT-CNN') / forward(hsi, lidar) interface."""
from . import _lib  # noqa: F401
from .model import ViTCNN  # noqa: F401
from .model_utils import get_model, test, val  # noqa: F401
from .scene import predict_scene_host  # noqa: F401
from .train import Trainer, train  # noqa: F401
from .datasets import MultiModalX  # noqa: F401

__all__ = ["ViTCNN", "get_model", "test", "val", "train", "Trainer", "MultiModalX", "predict_scene_host"]
