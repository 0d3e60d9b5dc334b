"""The C-ABI shared library loads on a CPU-only box and exports every function include/wae_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "wae_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(wae_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    lib = ctypes.CDLL(os.path.join(ROOT, "wavesandeigenvalues.jl_b200", "libwae_b200.so"))
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/wae_b200.h but not exported"


def test_python_binding_covers_the_header():
    from wae_b200 import _lib
    declared = set(_declared()) - set(_lib.HOST_DIAGNOSTICS)
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_no_gpu_means_loud_failure():
    """No CPU fallback: creating a context without a CUDA device must raise (skipped on the GPU box)."""
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from wae_b200 import _lib
    with pytest.raises(_lib.WaeError):
        _lib.Context(0)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "wavesandeigenvalues.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
