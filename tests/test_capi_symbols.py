"""CPU suite: the C-ABI library builds for sm_100a, loads, exports every symbol include/tce_b200.h declares, and
refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def built_lib():
    from tinychatengine_b200.build import build

    return build()


def declared_symbols():
    text = (ROOT / "include" / "tce_b200.h").read_text()
    return sorted(set(re.findall(r"TCE_API\s+[\w\s\*]+?\b(tce_\w+)\s*\(", text)))


def test_header_declares_something():
    syms = declared_symbols()
    assert "tce_w4a16_gemv" in syms and "tce_w8a8_matmul" in syms and "tce_attn_decode" in syms and len(syms) >= 20


def test_library_exports_every_declared_symbol(built_lib):
    L = C.CDLL(str(built_lib))
    for s in declared_symbols():
        assert hasattr(L, s), f"{s} declared in include/tce_b200.h but not exported by libtce_b200.so"


def test_python_binding_covers_header(built_lib):
    from tinychatengine_b200 import _lib

    assert sorted(_lib.SIGNATURES) == declared_symbols()
    _lib.lib()


def test_sass_is_blackwell_native(built_lib):
    import subprocess

    sass = subprocess.run(["cuobjdump", "-sass", str(built_lib)], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "UBLKCP" in sass, "TMA bulk copies missing from the W4A16 GEMV / attention kernels"
    assert "IMMA" in sass or "HMMA" in sass, "mma.sync path missing from the W4A16 GEMV"


@pytest.mark.skipif(torch.cuda.is_available(), reason="this checks the no-GPU behaviour")
def test_no_cpu_fallback(built_lib):
    from tinychatengine_b200 import _lib
    from tinychatengine_b200.runtime import Context

    L = _lib.lib()
    h = C.c_void_p()
    rc = L.tce_ctx_create(0, C.byref(h))
    assert rc == -3 and b"no CPU fallback" in L.tce_last_error()
    with pytest.raises(_lib.TceError):
        Context()


def test_zeros_width_matches_reference_rule(built_lib):
    from oracle import quant
    from tinychatengine_b200 import _lib, formats

    L = _lib.lib()
    for ic in (128, 1024, 1152, 4096, 5120, 11008, 13824, 14336):
        assert L.tce_zeros_width(ic, 128) == quant.calculate_zeros_width(ic, 128) == formats.zeros_width(ic, 128)
    assert L.tce_zeros_width(11008, 128) == 11 and L.tce_zeros_width(4096, 64) == 8
