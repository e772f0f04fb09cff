"""CPU: the algorithm behind csrc/embedding_tab.cu (Hermite tables of proj(sinusoid(.))) against the fp32 restatement of the
reference (oracle/network.structure_embedding, itself pinned to the reference's goldens by test_oracle_network.py) and
against the exact fp64 function.  The GPU kernel is compared with the same oracle in tests/test_network_gpu.py."""
import numpy as np
import torch

from gaussreg_b200.config import make_cfg
from gaussreg_b200.model import create_model
from oracle import embedding_table as oet
from oracle import network as onet


def _weights():
    torch.manual_seed(0)
    sd = create_model(make_cfg()).state_dict()
    p = "transformer.embedding"
    return sd, (sd[p + ".embedding.div_term"].numpy(), sd[p + ".proj_d.weight"].numpy(), sd[p + ".proj_d.bias"].numpy(),
                sd[p + ".proj_a.weight"].numpy(), sd[p + ".proj_a.bias"].numpy())


def _rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def test_table_sizes_match_the_library():
    from gaussreg_b200 import _lib
    for sigma_a in (15.0, 7.5, 30.0):
        rows = oet.nodes_a(sigma_a) * 2 + (oet.X_MAX_D * oet.INV_H_D + 1) * 3
        assert _lib.lib().gr_structure_embedding_table_floats(256, sigma_a) == rows * 256, sigma_a
    assert oet.nodes_a(15.0) == 99


def test_hermite_tables_reproduce_the_scalar_functions():
    """Interpolation error far below the fp32 evaluation error of the reference itself."""
    _, (div, Wd, bd, Wa, ba) = _weights()
    ta, td = oet.build_tables(div, Wd, bd, Wa, ba, 15.0)
    rng = np.random.default_rng(0)
    xa = (rng.random(4000) * 12.0).astype(np.float32)
    xd = (rng.random(4000) * 63.9).astype(np.float32)
    ea, ed = oet.exact(xa, div, Wa, ba), oet.exact(xd, div, Wd, bd)
    assert _rel(oet.hermite3(ta, xa), ea) < 2e-7
    assert _rel(oet.hermite5(td, xd), ed) < 2e-7
    # the fp32 reference path: sin / cos of an fp32-rounded phase, fp32 product
    ref_a = onet.sinusoidal_embedding(torch.from_numpy(xa), torch.from_numpy(div)) @ torch.from_numpy(Wa).T + torch.from_numpy(ba)
    assert _rel(ref_a.numpy(), ea) > _rel(oet.hermite3(ta, xa), ea)  # the table is the more accurate of the two
    # node values are exact, and the end points of both ranges are inside the tables
    assert np.array_equal(oet.hermite3(ta, np.float32([0.0, 1.0, 12.0])), ta[[0, 8, 96], 0])
    assert np.array_equal(oet.hermite5(td, np.float32([0.0, 0.5, 63.5])), td[[0, 1, 127], 0])


def test_tabulated_embedding_matches_the_reference_restatement():
    sd, (div, Wd, bd, Wa, ba) = _weights()
    g = torch.Generator().manual_seed(1)
    pts = (torch.rand(97, 3, generator=g) - 0.5) * torch.tensor([4.0, 3.0, 2.5])
    want = onet.structure_embedding(sd, pts, 0.2, 15, 3).numpy()
    d_idx, a_idx, _ = onet.embedding_indices(pts, 0.2, 15, 3)
    got = oet.structure_embedding(d_idx.numpy(), a_idx.numpy(), oet.build_tables(div, Wd, bd, Wa, ba, 15.0))
    assert got.shape == want.shape
    assert _rel(got, want) < 1e-6 and float(np.abs(got - want).max()) < 2e-5
