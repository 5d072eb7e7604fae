"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/opsg_b200.h declares with
the arity the ctypes binding assumes, and refuses to compute without an sm_100 device (no CPU fallback)."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

from openpsg_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "opsg_b200.h").read_text()


def _declared():
    text = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|size_t|const char\*)\s+(opsg_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args == "void" else len([a for a in args.split(",") if a.strip()])
    return out


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    decl = _declared()
    assert len(decl) >= 20
    for name, nargs in decl.items():
        assert hasattr(lib, name), f"{name} declared in include/opsg_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} missing from the ctypes binding"
        assert len(_lib.SIGNATURES[name]) == nargs, f"{name}: header has {nargs} args, binding {len(_lib.SIGNATURES[name])}"
    assert set(_lib.SIGNATURES) == set(decl)
    assert lib.opsg_version() >= 100


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device behaviour")
def test_compute_entry_points_fail_loudly_without_a_gpu():
    lib = _lib.load()
    assert lib.opsg_device_check() == _lib.OPSG_E_NO_DEVICE
    rc = lib.opsg_gemm_bf16(None, 8, None, 8, None, 8, 8, 8, 8, None, 0, None, 0, 0, 0, 1, None)
    assert rc == _lib.OPSG_E_NO_DEVICE
    assert "no CPU fallback" in _lib.last_error()
    from openpsg_b200 import ops
    with pytest.raises(ValueError):
        ops.gemm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device behaviour")
def test_head_refuses_cpu_inference():
    from openpsg_b200 import synth
    from tests.helpers import build_product_head
    head = build_product_head()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        head(synth.make_image_inputs(synth.WORKLOADS["cfg1"], 0))
