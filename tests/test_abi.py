"""CPU-side checks of the boundary: the C-ABI library loads without a GPU and exports
every symbol include/golfer_b200.h declares; the product fails loudly with no device."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import golfer_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "golfer_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(gs_[a-z_]+)\s*\(", hdr)))


def test_header_symbols_are_all_exported():
    assert os.path.exists(golfer_b200.library_path()), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(golfer_b200.library_path())
    declared = _declared_symbols()
    assert set(declared) == set(golfer_b200.host.ABI_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in golfer_b200.h but not exported"
    assert lib.gs_abi_version() == 1


def test_config_struct_matches_header_layout():
    # 4 + 8 + 2 + 8 + 4 int32 fields
    assert ctypes.sizeof(golfer_b200.host._GsConfig) == 4 * (4 + 8 + 2 + 8 + 4)


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback_without_device():
    L = golfer_b200.load_library()
    h = ctypes.c_void_p()
    rc = L.gs_create(ctypes.byref(h), 0, None, None, 0, 0, 0)
    assert rc == -2                                   # GS_ERR_NO_DEVICE
    assert b"no CPU fallback" in L.gs_last_error()
    with pytest.raises(golfer_b200.GolferError):
        golfer_b200.segment(torch.zeros(1, 4, 17, 3))
    with pytest.raises(golfer_b200.GolferError):
        golfer_b200.align(np.zeros((4, 17, 2), np.float32), np.zeros((4, 17, 2), np.float32))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "computer-vision-system-for-analyzing-golfer-action_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "oracle/_build" not in src and "liboracle" not in src, f
