"""TEST INFRASTRUCTURE ONLY. Imports the UNMODIFIED reference hot-path modules from
/root/reference through the stand-ins in oracle/_shim. Only usable in the authoring
container (the GPU box has no /root/reference); used by tests/golden/make_golden.py to
generate tests/golden/*.npz and by tests that pin oracle/cartnet_oracle.py."""
import importlib
import os
import sys

REF = os.environ.get("CARTNET_REFERENCE", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_shim")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "models", "cartnet.py"))


def load():
    """Returns (ref_cartnet_module, ref_model_utils_module, ref_dataset_utils_module, cfg)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF)
    for p in (_SHIM, REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    cartnet = importlib.import_module("models.cartnet")
    mutils = importlib.import_module("models.utils")
    dutils = importlib.import_module("dataset.utils")
    from torch_geometric.graphgym.config import cfg
    return cartnet, mutils, dutils, cfg
