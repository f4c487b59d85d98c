"""TEST-ONLY stand-in for the handful of torch_geometric symbols the reference's
hot-path files import (PyG is not installable offline). Used solely by
oracle/ref_loader.py to import /root/reference unmodified when generating
golden vectors. Never imported by the product package."""
