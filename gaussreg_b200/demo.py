"""Command-line equivalent of experiments/geotransformer.gaussian_splatting.indoor/demo.py on the B200 path:

    python -m gaussreg_b200.demo --ref_file A/point_cloud.ply --src_file B/point_cloud.ply \
        --weights weights/coarse_registration.pth.tar --output_path demo_outputs

Same arguments, same outputs (`estimated_transform.npz` with the transform between the ORIGINAL clouds,
`point_cloud_{ref,src,src_org}.ply`), no plyfile / open3d / fpsample: the 3DGS files are parsed by
gaussians.read_gaussian_ply, the preparation (demo.py:30-124) runs in csrc/gaussians.cu, the forward is the
gaussreg_b200 model.  Differences, by design: `estimated_transform` is the LocalGlobalRegistration result (the
reference's Open3D RANSAC post-step, model.py:209-215, is third-party and randomised), clouds with more than
`--num_sample` surviving Gaussians are rejected instead of FPS-subsampled (fpsample is third-party), and normals
are not estimated for the written point clouds.
"""
import argparse
import os

import numpy as np
import torch

from . import gaussians
from .config import make_cfg, NEIGHBOR_LIMITS
from .data import registration_collate_fn_stack_mode
from .model import create_model


def make_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument("--src_file", default="scene_name/B/output/point_cloud/iteration_30000/point_cloud.ply")
    parser.add_argument("--ref_file", default="scene_name/A/output/point_cloud/iteration_30000/point_cloud.ply")
    parser.add_argument("--output_path", default="demo_outputs")
    parser.add_argument("--weights", default="weights/coarse_registration.pth.tar",
                        help="checkpoint with a 'model' state_dict; 'random:<seed>' for a seeded random init")
    parser.add_argument("--num_sample", type=int, default=30000)
    return parser


def load_model(cfg, weights):
    if weights.startswith("random:"):
        seed = int(weights.split(":", 1)[1])
        torch.manual_seed(seed)
        np.random.seed(seed)
        return create_model(cfg).cuda().eval()
    model = create_model(cfg).cuda()
    state_dict = torch.load(weights, map_location="cuda")
    model.load_state_dict(state_dict["model"])  # strict, demo.py:143-144
    return model.eval()


def run(args):
    cfg = make_cfg()
    data_dict = gaussians.load_data(args.ref_file, args.src_file, args.num_sample)
    ref_color = data_dict["ref_feats"][:, 1:].cpu().numpy()
    src_color = data_dict["src_feats"][:, 1:].cpu().numpy()
    host = {k: data_dict[k] for k in ("ref_adjust_scale", "src_adjust_scale", "ref_center", "src_center")}
    tensors = {k: v for k, v in data_dict.items() if k not in host}
    model = load_model(cfg, args.weights)
    data = registration_collate_fn_stack_mode([tensors], cfg.backbone.num_stages, cfg.backbone.init_voxel_size,
                                              cfg.backbone.init_radius, NEIGHBOR_LIMITS)
    out = model(data)
    estimated_transform = out["estimated_transform"].cpu().numpy()
    ref_points, src_points = out["ref_points"].cpu().numpy(), out["src_points"].cpu().numpy()
    T = gaussians.unnormalize_transform(estimated_transform, host["ref_adjust_scale"], host["src_adjust_scale"],
                                        host["ref_center"], host["src_center"])
    ref_org = ref_points / host["ref_adjust_scale"] + host["ref_center"]
    src_org = src_points / host["src_adjust_scale"] + host["src_center"]
    os.makedirs(args.output_path, exist_ok=True)
    gaussians.write_point_cloud_ply(os.path.join(args.output_path, "point_cloud_src_org.ply"), src_org, src_color / 255)
    src_moved = src_org @ T[:3, :3].T.astype(np.float64) + T[:3, 3].astype(np.float64)
    gaussians.write_point_cloud_ply(os.path.join(args.output_path, "point_cloud_ref.ply"), ref_org, ref_color / 255)
    gaussians.write_point_cloud_ply(os.path.join(args.output_path, "point_cloud_src.ply"), src_moved, src_color / 255)
    path = gaussians.save_estimated_transform(args.output_path, T)
    return T, path


def main():
    T, path = run(make_parser().parse_args())
    print(T)
    print("wrote", path)


if __name__ == "__main__":
    main()
