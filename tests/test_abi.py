"""CPU test: the C-ABI library loads and exports every symbol include/b2me.h declares.
No compute call is made here (no GPU in the CPU test tier)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in os.listdir(os.path.join(ROOT, "include")):
        if not h.endswith(".h"):
            continue
        txt = open(os.path.join(ROOT, "include", h)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        names |= set(re.findall(r"\b(b2(?:me|fr|fp|tq|dbk)_\w+)\s*\(", txt))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    so = os.path.join(ROOT, "h264_b200", "libb2me.so")
    if not os.path.exists(so):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(so)
    syms = declared_symbols()
    assert len(syms) >= 15
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_product_does_not_import_oracle():
    """The product package must never route through oracle/ (no CPU fallback)."""
    for dp, _, fs in os.walk(os.path.join(ROOT, "h264_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liborc" not in txt, f


def test_missing_extension_fails_loudly(monkeypatch, tmp_path):
    from h264_b200 import api
    monkeypatch.setattr(api, "_lib", None)
    monkeypatch.setattr(api, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(api.B2Error):
        api.lib()
