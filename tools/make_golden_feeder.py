"""Generates tests/golden/feeder_small.pt: decoded-sample inputs (seeded, dpft_b200.feeder.synthetic_raw_batch) pushed through
the UNMODIFIED reference's per-sample dataset steps (KRadarDataset.scale_radar_data / _add_transformations / _add_projections /
_add_shape / resize_image, dataset.py:143-169) and torch's default_collate (loader.py:27).  Build container only.

  python tools/make_golden_feeder.py
"""
import os
import sys

import torch
from torch.utils.data import default_collate

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, ".."))

import reference_shim  # noqa: E402
from dpft_b200 import feeder  # noqa: E402

CASES = [
    dict(name="fusion_resize_int", inputs=["camera_mono", "radar_bev", "radar_front"], batch=2, seed=41, image_size=24, scale=True,
         sizes={"camera_mono": (36, 64, 3), "radar_bev": (8, 16, 6), "radar_front": (6, 16, 6)}),
    dict(name="camera_resize_tuple", inputs=["camera_mono"], batch=1, seed=42, image_size=(20, 30), scale=True,
         sizes={"camera_mono": (45, 80, 3)}),
    dict(name="radar_unscaled", inputs=["radar_bev"], batch=3, seed=43, image_size=None, scale=False,
         sizes={"radar_bev": (4, 10, 6)}),
]


def reference_batch(ds, case):
    raw = feeder.synthetic_raw_batch(case["inputs"], case["batch"], seed=case["seed"], sizes=case["sizes"])
    samples = []
    for i in range(case["batch"]):
        sample = {k: (v[i].float() if v.dtype == torch.uint8 else v[i]) for k, v in raw.items()}   # read_image(...).type(float32)
        samples.append(reference_shim.reference_sample_pipeline(ds, sample, case["image_size"], case["scale"]))
    return default_collate(samples)


def main():
    ds = reference_shim.import_reference_dataset()
    out = []
    for case in CASES:
        want = reference_batch(ds, case)
        out.append(dict(case=case, batch={k: v.clone() for k, v in want.items()}))
        print(case["name"], {k: tuple(v.shape) for k, v in want.items()})
    path = os.path.join(HERE, "..", "tests", "golden", "feeder_small.pt")
    torch.save({"cases": out, "torch_version": torch.__version__}, path)
    print(os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
