"""CPU tests: oracle/network.py (torch fp32 restatement) pinned against golden vectors produced by the
UNMODIFIED reference modules (tests/golden/make_network_golden.py), plus the proof that
gaussreg_b200.model.create_model rebuilds the reference's seeded weights."""
import os

import numpy as np
import pytest
import torch

from oracle import network as onet
from tests.helpers import GOLDEN_DIR, golden_spec, oracle_data, rel_l2, seeded_model

ROW_STRIDE = {"encoder1_2": 97, "encoder2_3": 61, "encoder3_3": 31, "encoder4_3": 13, "encoder5_3": 7,
              "decoder4": 13, "decoder3": 31, "decoder2": 61}


def test_seeded_weights_match_reference_checksums():
    g = np.load(os.path.join(GOLDEN_DIR, "weights_checksum.npz"))
    sd = seeded_model(0).state_dict()
    names = list(g["names"])
    assert sorted(sd) == names and len(names) == 316
    for i, k in enumerate(names):
        assert repr(tuple(sd[k].shape)) == str(g["shapes"][i])
        s = np.array([float(sd[k].double().sum()), float(sd[k].double().abs().sum())])
        assert np.array_equal(s, g["sums"][i]), k


@pytest.mark.parametrize("case", ["room5k", "textured3k"])
def test_oracle_forward_matches_reference_golden(case):
    gold = np.load(os.path.join(GOLDEN_DIR, f"network_golden_{case}.npz"))
    data = oracle_data(golden_spec(gold))
    assert np.array_equal(np.stack([l.numpy() for l in data["lengths"]]), gold["lengths"])
    widths = [t.shape[1] for t in data["neighbors"]] + [t.shape[1] for t in data["subsampling"]] + [t.shape[1] for t in data["upsampling"]]
    assert widths == gold["widths"].tolist()
    sd = seeded_model(0).state_dict()
    taps = {}
    with torch.no_grad():
        out = onet.forward(sd, data, taps=taps)
    # float stages: same ATen kernels, so agreement is at rounding level; tolerance 1e-5 rel-L2
    for name, stride in ROW_STRIDE.items():
        assert rel_l2(taps[name][::stride], gold["bb/" + name]) < 1e-5, name
    n_ref = taps["ref_embeddings"].shape[0]
    n_src = taps["src_embeddings"].shape[0]
    assert rel_l2(taps["ref_embeddings"][[0, n_ref // 2]], gold["emb/ref_rows"]) < 1e-5
    assert rel_l2(taps["src_embeddings"][[1, n_src - 1]], gold["emb/src_rows"]) < 1e-5
    assert rel_l2(out["ref_feats_c"], gold["ref_feats_c"]) < 1e-4
    assert rel_l2(out["src_feats_c"], gold["src_feats_c"]) < 1e-4
    # discrete selections: the 256 superpoint pairs must be the same SET; two entries whose scores differ by
    # ~1 ulp may swap places (observed gap 2.8e-7 relative), which permutes patches but not the result
    got = set(zip(out["ref_node_corr_indices"].tolist(), out["src_node_corr_indices"].tolist()))
    want = set(zip(gold["ref_node_corr_indices"].tolist(), gold["src_node_corr_indices"].tolist()))
    assert got == want
    same_order = np.array_equal(out["ref_node_corr_indices"].numpy(), gold["ref_node_corr_indices"]) and np.array_equal(
        out["src_node_corr_indices"].numpy(), gold["src_node_corr_indices"])
    assert out["corr_scores"].shape[0] == gold["corr_scores"].shape[0]
    if same_order:
        ms = out["matching_scores"]
        assert rel_l2(ms[[0, ms.shape[0] // 2, ms.shape[0] - 1]], gold["matching_scores_sample"]) < 1e-4
        assert rel_l2(out["ref_corr_points"], gold["ref_corr_points"]) < 1e-6
        assert rel_l2(out["corr_scores"], gold["corr_scores"]) < 1e-3
    # north_star tolerance on the transform: 1e-4 Frobenius
    assert float(np.linalg.norm(out["estimated_transform"].numpy() - gold["estimated_transform"])) < 1e-4


def test_config_matches_reference_config():
    """gaussreg_b200.config.make_cfg() carries exactly the model-related values of the reference's config.py
    (golden produced by importing the unmodified reference, tests/golden/make_config_golden.py)."""
    import json
    from gaussreg_b200.config import make_cfg
    gold = json.load(open(os.path.join(GOLDEN_DIR, "config_golden.json")))
    cfg = make_cfg()
    for section, values in gold.items():
        for key, want in values.items():
            assert key in cfg[section], (section, key)
            got = cfg[section][key]
            if isinstance(want, float):
                assert float(got) == want, (section, key, got, want)
            else:
                assert got == want, (section, key, got, want)
