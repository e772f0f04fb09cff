"""CPU test: the C-ABI library builds for sm_100a, loads, and exports every declared symbol."""
import ctypes
import os
import re

from gaussreg_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "gaussreg_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(gr_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_header_symbols():
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 7
    for name in names:
        assert hasattr(lib, name), name
    # the Python loader knows every declared symbol, and nothing else
    assert sorted(_lib.exported_symbols()) == names


def test_workspace_queries_and_version():
    L = _lib.lib()
    assert b"sm_100a" in L.gr_version()
    assert L.gr_grid_subsample_workspace_size(60000, 2) > 60000 * 12
    assert L.gr_radius_neighbors_workspace_size(60000, 60000, 2) > 60000 * 16
    assert L.gr_launch_count() == 0
