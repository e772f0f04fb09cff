"""Shared test helpers (TEST INFRASTRUCTURE): seeded weights, oracle-side data pyramid."""
import ast
import os

import numpy as np
import torch

from gaussreg_b200.config import make_cfg, NEIGHBOR_LIMITS
from gaussreg_b200.model import create_model
from gaussreg_b200.synthetic import make_pair_inputs
from oracle import neighbors as on

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def seeded_model(seed=0):
    """The reference's `create_model(make_cfg())` weights under torch/numpy seed `seed` (CPU, eval)."""
    torch.manual_seed(seed)
    np.random.seed(seed)
    m = create_model(make_cfg())
    m.eval()
    return m


def oracle_data(spec, limits=NEIGHBOR_LIMITS):
    """Pair inputs + the CPU neighbour pyramid (reference build when it travelled, else the C port)."""
    d = make_pair_inputs(**spec)
    n_ref, n_src = d["ref_points"].shape[0], d["src_points"].shape[0]
    pts = np.concatenate([d["ref_points"], d["src_points"]]).astype(np.float32)
    lens = np.array([n_ref, n_src], np.int64)
    impl = on.ref() if on.have_ref() else on.port()
    cfg = make_cfg()
    pyr = on.precompute_data_stack_mode(impl, pts, lens, cfg.backbone.num_stages, cfg.backbone.init_voxel_size,
                                        cfg.backbone.init_radius, limits)
    data = {k: [torch.from_numpy(np.ascontiguousarray(a)) for a in v] for k, v in pyr.items()}
    data["features"] = torch.from_numpy(np.concatenate([d["ref_feats"], d["src_feats"]]).astype(np.float32))
    data["transform"] = torch.from_numpy(d["transform"])
    return data


def golden_spec(gold):
    return ast.literal_eval(str(gold["spec"]))


def rel_l2(a, b):
    a = torch.as_tensor(a).double().flatten()
    b = torch.as_tensor(b).double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def to_cuda(data):
    out = {}
    for k, v in data.items():
        if isinstance(v, list):
            out[k] = [t.cuda() for t in v]
        elif isinstance(v, torch.Tensor):
            out[k] = v.cuda()
        else:
            out[k] = v
    return out
