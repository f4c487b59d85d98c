"""CPU: the C-ABI library builds, loads, and exports exactly the symbols include/cartnet_b200.h declares
(no compute calls -- there is no GPU here)."""
import os
import re
import subprocess

from cartnet_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "cartnet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(cartnet_[a-z0-9_]+)\s*\(", src))


def test_library_is_built():
    assert os.path.isfile(_lib.LIB_PATH), "run __graft_entry__.build() first"


def test_header_and_binding_agree():
    assert _declared() == set(_lib.SIGNATURES), (_declared() ^ set(_lib.SIGNATURES))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\sT\s+(cartnet_[a-z0-9_]+)", out))
    assert _declared() <= exported, _declared() - exported
    assert lib.cartnet_version() >= 100
    assert lib.cartnet_last_error() is not None


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """No CPU fallback: without the built shared library every operator raises."""
    import pytest
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libcartnet_b200.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()
