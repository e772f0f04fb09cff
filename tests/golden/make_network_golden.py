"""Generates tests/golden/network_golden_*.npz by running the UNMODIFIED Python reference
(/root/reference, imported through tests/golden/ref_harness.py) on seeded synthetic pairs, CPU fp32.

    python tests/golden/make_network_golden.py

Stored per case: the stage sizes, row-subsampled backbone / transformer activations, the discrete
selections (superpoint correspondences), Sinkhorn samples, the LGR correspondences and transform, and
per-tensor checksums of the reference's seeded weights (so that the product can prove it rebuilt the
same weights without shipping 114 MB).
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_harness as rh  # noqa: E402
from gaussreg_b200.synthetic import make_pair_inputs  # noqa: E402

CASES = {
    # BASELINE.json configs[0]: 5k-point pair, reference CPU path
    "room5k": dict(seed=0, n_points=5000),
    # well-conditioned pair: same texture seen by both clouds, small motion
    "textured3k": dict(seed=5, n_points=3000, textured=True, angle=0.1, translation=(0.1, -0.05, 0.05)),
}
LIMITS = [89, 30, 43, 49, 49]
ROW_STRIDE = {"encoder1_2": 97, "encoder2_3": 61, "encoder3_3": 31, "encoder4_3": 13, "encoder5_3": 7,
              "decoder4": 13, "decoder3": 31, "decoder2": 61}


def run_case(net, cfg, name, spec):
    from geotransformer.utils.data import registration_collate_fn_stack_mode

    d = make_pair_inputs(**spec)
    dd = {k: d[k] for k in ("ref_points", "src_points", "ref_feats", "src_feats")}
    data = registration_collate_fn_stack_mode([dd], cfg.backbone.num_stages, cfg.backbone.init_voxel_size,
                                              cfg.backbone.init_radius, LIMITS)
    # mirrors demo.py:148 to_cuda(): makes the [:, :limit] slices contiguous
    for key in ("neighbors", "subsampling", "upsampling"):
        data[key] = [t.contiguous() for t in data[key]]
    taps = {}
    hooks = []
    for bname in ROW_STRIDE:
        mod = getattr(net.backbone, bname)
        hooks.append(mod.register_forward_hook(lambda m, i, o, bname=bname: taps.__setitem__(bname, o.detach())))
    hooks.append(net.transformer.embedding.register_forward_hook(
        lambda m, i, o: taps.setdefault("embeddings", []).append(o.detach())))
    hooks.append(net.fine_matching.register_forward_hook(lambda m, i, o: taps.__setitem__("lgr", o)))
    hooks.append(net.coarse_matching.register_forward_hook(lambda m, i, o: taps.__setitem__("coarse", o)))
    import model as ref_model
    ref_model.registration_with_ransac_from_correspondences = lambda *a, **k: taps["lgr"][3].numpy()
    t = time.time()
    with torch.no_grad():
        out = net(data)
    print(name, "reference forward %.1fs" % (time.time() - t))
    for h in hooks:
        h.remove()
    g = {}
    g["spec"] = np.array(repr(spec))
    g["lengths"] = np.stack([l.numpy() for l in data["lengths"]])
    g["widths"] = np.array([t.shape[1] for t in data["neighbors"]] + [t.shape[1] for t in data["subsampling"]]
                           + [t.shape[1] for t in data["upsampling"]])
    for bname, stride in ROW_STRIDE.items():
        g["bb/" + bname] = taps[bname][::stride].numpy()
    emb_ref, emb_src = taps["embeddings"]
    g["emb/ref_rows"] = emb_ref[0, [0, emb_ref.shape[1] // 2]].numpy()
    g["emb/src_rows"] = emb_src[0, [1, emb_src.shape[1] - 1]].numpy()
    g["ref_feats_c"] = out["ref_feats_c"].numpy()
    g["src_feats_c"] = out["src_feats_c"].numpy()
    g["ref_node_corr_indices"] = out["ref_node_corr_indices"].numpy()
    g["src_node_corr_indices"] = out["src_node_corr_indices"].numpy()
    g["node_corr_scores"] = taps["coarse"][2].numpy()
    ms = out["matching_scores"]
    g["matching_scores_sample"] = ms[[0, ms.shape[0] // 2, ms.shape[0] - 1]].numpy()
    g["matching_scores_rowsum"] = ms.exp().sum(2).numpy()
    g["ref_corr_points"] = out["ref_corr_points"].numpy()
    g["src_corr_points"] = out["src_corr_points"].numpy()
    g["corr_scores"] = out["corr_scores"].numpy()
    g["estimated_transform"] = taps["lgr"][3].numpy()
    g["gt_transform"] = d["transform"]
    return g, data, out


def main():
    net, cfg = rh.create_reference_model(0)
    sd = net.state_dict()
    names = sorted(sd)
    sums = np.array([[float(sd[k].double().sum()), float(sd[k].double().abs().sum())] for k in names])
    np.savez_compressed(os.path.join(HERE, "weights_checksum.npz"), names=np.array(names), sums=sums,
                        shapes=np.array([repr(tuple(sd[k].shape)) for k in names]))
    for name, spec in CASES.items():
        g, _, _ = run_case(net, cfg, name, spec)
        path = os.path.join(HERE, f"network_golden_{name}.npz")
        np.savez_compressed(path, **g)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB; C =", g["corr_scores"].shape[0])
        print(np.round(g["estimated_transform"], 4))
        print(np.round(g["gt_transform"], 4))


if __name__ == "__main__":
    main()
