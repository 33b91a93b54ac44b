"""The C-ABI library loads on a CPU-only box and exports every symbol include/genre_b200.h declares
(no compute calls here)."""
import ctypes
import os
import re

import pytest

from conftest import REPO
from genre_shapehd_b200 import _lib

HEADER = os.path.join(REPO, "include", "genre_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(genre_b200_\w+)\s*\(", text)))


def test_header_declares_what_binding_binds():
    assert declared_symbols() == _lib.EXPORTED_SYMBOLS


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "build the library first (__graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), "missing export: " + name


def test_version_and_error_string():
    lib = _lib.load()
    assert lib.genre_b200_version() >= 1000
    # argument errors are reported without touching the device
    rc = lib.genre_b200_nnd_forward(None, None, 1, 1, 1, None, None, None, None, None)
    assert rc == -1
    assert b"null" in lib.genre_b200_last_error()
    assert lib.genre_b200_voxelize_workspace_bytes(32, 256 * 256, 128) > 32 * 256 * 256 * 20


def test_workspace_too_small_is_rejected():
    lib = _lib.load()
    dummy = ctypes.c_void_p(256)  # never dereferenced: the size check comes first
    rc = lib.genre_b200_cam_bp_forward(dummy, 1, 1, 16, 16, 256, 256, 16, 1, dummy, 1, 1, dummy, 1, 1, dummy, None, 16,
                                       0, dummy, 16, None)
    assert rc == -2


def test_cpu_tensors_are_refused():
    import torch
    from nndistance.functions.nnd import nndistance
    from toolbox.cam_bp.cam_bp.modules.camera_backprojection_module import Camera_back_projection_layer
    with pytest.raises((RuntimeError, AssertionError)):
        nndistance(torch.zeros(1, 4, 3), torch.zeros(1, 5, 3))
    with pytest.raises((RuntimeError, AssertionError)):
        Camera_back_projection_layer()(torch.zeros(1, 1, 8, 8))
