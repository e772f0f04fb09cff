"""Configuration of the coarse-registration model: the values of
experiments/geotransformer.gaussian_splatting.indoor/config.py:79-125 and demo.py:136, without the
reference's import-time directory creation (config.py:26-31)."""


class _Cfg(dict):
    __getattr__ = dict.__getitem__

    def __setattr__(self, k, v):
        self[k] = v


NEIGHBOR_LIMITS = [89, 30, 43, 49, 49]  # demo.py:136 / test.py:129


def make_cfg():
    c = _Cfg()
    c.seed = 7351
    b = c.backbone = _Cfg()
    b.num_stages = 5
    b.init_voxel_size = 0.025
    b.kernel_size = 15
    b.base_radius = 2.5
    b.base_sigma = 2.0
    b.init_radius = b.base_radius * b.init_voxel_size
    b.init_sigma = b.base_sigma * b.init_voxel_size
    b.group_norm = 32
    b.input_dim = 4
    b.init_dim = 64
    b.output_dim = 256
    m = c.model = _Cfg()
    m.ground_truth_matching_radius = 0.05
    m.num_points_in_patch = 128
    m.num_sinkhorn_iterations = 100
    cm = c.coarse_matching = _Cfg()
    cm.num_targets = 128
    cm.overlap_threshold = 0.1
    cm.num_correspondences = 256
    cm.dual_normalization = True
    g = c.geotransformer = _Cfg()
    g.input_dim = 2048
    g.hidden_dim = 256
    g.output_dim = 256
    g.num_heads = 4
    g.blocks = ["self", "cross", "self", "cross", "self", "cross"]
    g.sigma_d = 0.2
    g.sigma_a = 15
    g.angle_k = 3
    g.reduction_a = "max"
    f = c.fine_matching = _Cfg()
    f.topk = 3
    f.acceptance_radius = 0.1
    f.mutual = True
    f.confidence_threshold = 0.05
    f.use_dustbin = False
    f.use_global_score = False
    f.correspondence_threshold = 3
    f.correspondence_limit = None
    f.num_refinement_steps = 5
    return c
