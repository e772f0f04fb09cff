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
    # T1 table size query (no device work): angle nodes for sigma_a = 15 are [0, 12] at step 1/8 (+ guard), distance nodes
    # [0, 1024] at step 1/2, (f, h f') resp. (f, h f', h^2 f'') rows of hidden_dim channels
    n = L.gr_structure_embedding_table_floats(256, 15.0)
    assert n == (99 * 2 + 2049 * 3) * 256
    assert L.gr_structure_embedding_table_floats(250, 15.0) == 0 and L.gr_structure_embedding_table_floats(256, 0.0) == 0


def test_module_key_cache_follows_parameter_changes():
    """Host logic: the native parameter structs are keyed on (version, storage) of a CACHED flat tensor list (walking the
    module tree costs 0.5 ms per call); the list must be dropped whenever Parameters can have been replaced."""
    import torch
    import torch.nn as nn

    from gaussreg_b200 import ops
    m = nn.Sequential(nn.Linear(4, 4), nn.GroupNorm(2, 4))
    k0 = ops._module_key(m)
    assert ops._module_key(m) == k0 and len(k0) == 4
    with torch.no_grad():
        m[0].weight.add_(1.0)  # in-place edit: version counter
    k1 = ops._module_key(m)
    assert k1 != k0
    sd = {k: v.clone() + 1 for k, v in m.state_dict().items()}
    m.load_state_dict(sd, assign=True)  # replaces the Parameter objects: the post-hook drops the cached list
    k2 = ops._module_key(m)
    assert k2 != k1 and all(p.data_ptr() in [d for _, d in k2] for p in m.parameters())
    m[0].weight = nn.Parameter(torch.zeros(4, 4))  # manual surgery needs the documented invalidation
    ops.invalidate_weight_caches(m)
    assert m[0].weight.data_ptr() in [d for _, d in ops._module_key(m)]


def test_stacked_rows_detects_adjacent_views():
    """Host logic: row-wise ops run once on both clouds when they are adjacent views of one matrix."""
    import torch

    from gaussreg_b200 import ops
    x = torch.arange(40, dtype=torch.float32).reshape(10, 4)
    both = ops.stacked_rows(x[:3], x[3:])
    assert both is not None and both.shape == (10, 4) and both.data_ptr() == x.data_ptr() and torch.equal(both, x)
    assert ops.stacked_rows(x[:3].clone(), x[3:]) is None      # different storage
    assert ops.stacked_rows(x[:3], x[4:]) is None              # a gap between them
    assert ops.stacked_rows(x[:3, :2], x[3:, :2]) is None      # not contiguous
    assert ops.stacked_rows(x[:3], x[3:].double()) is None     # dtype


def test_argument_checks_need_no_device():
    """The C ABI rejects malformed calls before touching CUDA (GR_ERR_BAD_ARG = -1): checked here without a GPU."""
    L = _lib.lib()
    assert L.gr_grid_subsample_chain(None, None, 2, 100, None, 4, None, None, None, None, 0, None, None) == -1
    assert L.gr_grid_subsample_chain(None, None, 0, 100, None, 0, None, None, None, None, 0, None, None) == -1
    # angle_k outside 1..3, hidden_dim not a multiple of 64, sigma_a <= 0
    for k, C, sa in ((0, 256, 15.0), (4, 256, 15.0), (3, 200, 15.0), (3, 256, 0.0)):
        assert L.gr_structure_embedding_tabulated(None, None, 10, k, None, sa, None, C, None, None, None, None, None, None) == -1
    # zero rows: nothing to do
    assert L.gr_structure_embedding_tabulated(None, None, 0, 3, None, 15.0, None, 256, None, None, None, None, None, None) == 0
    assert L.gr_structure_embedding_build_table(None, 256, None, None, None, None, 15.0, None, None) == -1
