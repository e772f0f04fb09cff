"""TEST INFRASTRUCTURE: golden neighbour limits from the UNMODIFIED reference `calibrate_neighbors_stack_mode`
(geotransformer/utils/data.py:192-217) driven by the reference collate + the reference's own C++ ops (oracle/_ref),
on a small synthetic dataset.

    python tests/golden/make_calibration_golden.py     # writes tests/golden/calibration_golden.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

SPEC = dict(n_pairs=3, n_points=2500, num_stages=4, voxel_size=0.025, search_radius=0.0625, keep_ratio=0.8, sample_threshold=2000)


def dataset(spec=SPEC):
    from gaussreg_b200.synthetic import make_pair_inputs
    keys = ("ref_points", "src_points", "ref_feats", "src_feats")
    return [{k: make_pair_inputs(50 + i, spec["n_points"])[k] for k in keys} for i in range(spec["n_pairs"])]


def main():
    import ref_harness
    ref_harness.install()
    from geotransformer.utils import data as ref_data
    s = SPEC
    limits = ref_data.calibrate_neighbors_stack_mode(dataset(), ref_data.registration_collate_fn_stack_mode, s["num_stages"],
                                                     s["voxel_size"], s["search_radius"], s["keep_ratio"], s["sample_threshold"])
    print("reference neighbour limits:", limits)
    np.savez(os.path.join(HERE, "calibration_golden.npz"), neighbor_limits=np.asarray(limits), spec=repr(SPEC))


if __name__ == "__main__":
    main()
