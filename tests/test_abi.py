"""The C-ABI library loads and exports every symbol that include/irc_b200.h declares (no compute: runs without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "irc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(irc_[A-Za-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import irc_b200  # noqa: F401
    from irc_b200 import _native
    lib = _native.lib()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} is declared in include/irc_b200.h but not exported by libirc_sm100.so"
    assert sorted(_native.EXPORTS) == syms, set(_native.EXPORTS) ^ set(syms)
    assert lib.irc_version() >= 100


def test_struct_sizes_match_header():
    """ctypes mirrors of the argument structs must have the C layout (spot-check through sizeof arithmetic)."""
    from irc_b200 import _native as n
    assert ctypes.sizeof(n.CView) == 40
    assert ctypes.sizeof(n.TapArgs) == 4 * (2 + 128 + 8)
    assert ctypes.sizeof(n.ConvGemmArgs) % 8 == 0 and ctypes.sizeof(n.TnGemmArgs) % 8 == 0


def test_no_fallback_without_library(tmp_path, monkeypatch):
    """The product must fail loudly when the CUDA library is missing."""
    from irc_b200 import _native as n
    monkeypatch.setattr(n, "_lib", None)
    monkeypatch.setattr(n, "LIB_PATH", str(tmp_path / "nope.so"))
    try:
        n.lib()
        raise AssertionError("expected IrcError")
    except n.IrcError as e:
        assert "no CPU or PyTorch fallback" in str(e)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "infrared-colorization-with-resnet-generator-and-patchgan_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            txt = open(os.path.join(pkg, f)).read()
            assert "irc_oracle" not in txt and "ref_backend" not in txt, f
