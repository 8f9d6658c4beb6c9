"""The C-ABI library loads without a GPU and exports every symbol include/risltc_cuda.h declares; entry points that need a
device fail loudly (non-zero + message) instead of falling back to anything."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "risltc_cuda.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(risltc_cuda_\w+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    from risltc_b200 import api
    lib = api.lib()
    names = declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f"{name} is declared in include/risltc_cuda.h but not exported"


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from risltc_b200 import api
    with pytest.raises(api.RisltcError, match="no CUDA device"):
        api.Device(0)
    assert b"no CPU fallback" in api.lib().risltc_cuda_last_error()
    # null device handles are rejected, not dereferenced
    assert api.lib().risltc_cuda_render_frame(None, None, C.c_uint32(0)) != 0


def test_product_code_does_not_import_the_oracle():
    """oracle/ is test infrastructure: nothing under risltc_b200/ or include/ may reference it."""
    for path in list((ROOT / "risltc_b200").rglob("*.py")) + list((ROOT / "risltc_b200").rglob("*.c*")) + list((ROOT / "risltc_b200").rglob("*.h")):
        text = path.read_text(errors="ignore")
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), path
        assert "liborc" not in text and "orc_" not in text, path
