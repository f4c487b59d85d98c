"""cartnet_b200 -- B200-native implementation of CartNet's message-passing hot path.

Public surface (mirrors the reference's `models/cartnet.py` and `dataset/utils.py`):
    CartNet, Encoder, CartNet_layer, Cholesky_head, Scalar_head, radius_graph_pbc
Importing the package does not need a GPU; calling any operator does, and raises if the
CUDA library (cartnet_b200/libcartnet_b200.so) is missing -- there is no CPU fallback.
"""
from . import augment  # noqa: F401
from .batch import CrystalBatch, DeferredScalars, DevicePrefetcher, collate  # noqa: F401
from .cartnet import CartNet, CartNet_layer, Cholesky_head, Encoder, Scalar_head  # noqa: F401
from .device_dataset import DeviceDataset, DeviceLoader  # noqa: F401
from .functional import compute_loss  # noqa: F401
from .radius_graph import build_graph, radius_graph_pbc  # noqa: F401

__all__ = ["CartNet", "Encoder", "CartNet_layer", "Cholesky_head", "Scalar_head", "radius_graph_pbc",
           "build_graph", "compute_loss", "CrystalBatch", "DevicePrefetcher", "DeferredScalars", "collate", "DeviceDataset", "DeviceLoader"]
