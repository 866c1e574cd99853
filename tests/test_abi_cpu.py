"""The C-ABI library loads on a CPU-only box and exports every symbol include/b200em.h declares (no compute calls)."""
import ctypes
import os

from torch_em_b200 import _lib


def test_library_exports_every_declared_symbol():
    protos = _lib.parse_header()
    assert len(protos) >= 30
    assert os.path.exists(_lib.LIB_PATH), "build with `python torch-em_b200/build.py`"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [name for name in protos if not hasattr(lib, name)]
    assert not missing, missing
    loaded = _lib.load()
    assert loaded.b200em_abi_version() == 1
    assert loaded.b200em_conv3d_umma_supported(32, 32, 3, 3, 3) == 1
    assert loaded.b200em_conv3d_umma_supported(1, 32, 3, 3, 3) == 0
    assert loaded.b200em_conv3d_wgrad_umma_supported(64, 48, 3, 3, 3) == 1


def test_header_cites_reference_lines():
    src = open(_lib.HEADER).read()
    for ref in ("unet.py:429-438", "unet.py:391-406", "loss/dice.py:34-93", "transform/label.py:248-327", "unet.py:456"):
        assert ref in src, ref
