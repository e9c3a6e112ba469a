"""CPU tier: the C-ABI library builds for sm_100a, loads, and exports every symbol that
include/b2aruco.h declares (no compute calls here: there is no GPU in this tier)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def so_path():
    from aruco_slam_b200 import _lib
    return _lib.build()


def test_header_symbols_exported(so_path):
    from aruco_slam_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "b2aruco.h")).read()
    declared = sorted(set(re.findall(r"^\s*(?:const char \*|int\s+|void \*|void\s+)(b2a_[a-z_0-9]+)\s*\(", hdr, re.M)))
    assert declared, "no declarations found"
    L = ctypes.CDLL(so_path)
    for name in declared:
        assert hasattr(L, name), name
    assert sorted(_lib.SYMBOLS) == declared


def test_predefined_dictionary_tables_without_gpu(so_path):
    """b2a_get_predefined_dictionary is host-only data: same bytes as the package tables."""
    import numpy as np
    from aruco_slam_b200 import aruco, dictionaries as D
    for did in range(22):
        a, b = aruco.library_dictionary(did), D.getPredefinedDictionary(did)
        assert np.array_equal(a.table, b.table) and a.max_correction_bits == b.max_correction_bits


def test_create_fails_loudly_without_gpu(so_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from aruco_slam_b200 import aruco, dictionaries as D
    with pytest.raises(aruco.B2AError):
        aruco.ArucoDetector(D.getPredefinedDictionary(0), max_shape=(64, 64))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "aruco_slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower(), f
