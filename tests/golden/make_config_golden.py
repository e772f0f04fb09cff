"""TEST INFRASTRUCTURE: the model-related subtrees of the UNMODIFIED reference's `make_cfg()`
(experiments/geotransformer.gaussian_splatting.indoor/config.py) as JSON, for the CPU config-parity test.

    python tests/golden/make_config_golden.py     # writes tests/golden/config_golden.json
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
KEYS = ("backbone", "model", "coarse_matching", "geotransformer", "fine_matching")


def main():
    import ref_harness
    ref_harness.install()
    import config as ref_config
    cfg = ref_config.make_cfg()
    out = {k: dict(cfg[k]) for k in KEYS}
    with open(os.path.join(HERE, "config_golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(json.dumps(out)[:400])


if __name__ == "__main__":
    main()
