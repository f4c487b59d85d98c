from types import SimpleNamespace

# the hot path reads only cfg.radius (cartnet.py:201) and cfg.invariant (cartnet.py:156)
cfg = SimpleNamespace(radius=5.0, invariant=False)
