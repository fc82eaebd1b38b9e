"""Generates tests/golden/metrics_small.pt: seeded predictions / labels through the UNMODIFIED reference metrics
(dprt.evaluation.metric.Metric with mAP3D + mGIoU3D, config/kradar.json 'evaluate' section).  pytorch3d is absent:
dpft_b200.criterion.box3d_overlap stands in for pytorch3d.ops.box3d_overlap (tools/reference_shim.py), so the fixture pins
the metric logic around that function.

  python tools/make_golden_metrics.py
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, ".."))

import reference_shim  # noqa: E402
from dpft_b200 import criterion  # noqa: E402

EVALUATE = {"metrics": {"mAP": "mAP3D", "mGIoU": "mGIoU3D"}}

CASES = [dict(name="mixed", seed=21, B=4, N=50, counts=[3, 6, 1, 4], C=2, near=True),
         dict(name="three_classes", seed=22, B=3, N=40, counts=[5, 2, 8], C=3, near=True),
         dict(name="no_overlap", seed=23, B=2, N=30, counts=[4, 3], C=2, near=False)]


def make_case(case):
    """Labels as the dataset makes them; predictions = noise plus, when ``near``, perturbed copies of the ground-truth boxes
    with high scores on their class (so that true positives, duplicates and misses all occur)."""
    g = torch.Generator().manual_seed(case["seed"])
    B, N, C = case["B"], case["N"], case["C"]
    rnd = lambda *s: torch.randn(*s, generator=g)
    out = {"class": rnd(B, N, C), "center": torch.stack((torch.rand(B, N, generator=g) * 72, rnd(B, N) * 6, rnd(B, N)), -1),
           "size": torch.relu(rnd(B, N, 3) + 2), "angle": torch.tanh(rnd(B, N, 2))}
    labels = []
    for b, m in enumerate(case["counts"]):
        a = torch.rand(m, generator=g) * 6.28
        cls = torch.randint(0, C, (m,), generator=g)
        lab = {"gt_class": torch.nn.functional.one_hot(cls, C).float(),
               "gt_center": torch.stack((torch.rand(m, generator=g) * 72, rnd(m) * 6, rnd(m)), -1),
               "gt_size": torch.rand(m, 3, generator=g) * 3 + 1.5, "gt_angle": torch.stack((torch.sin(a), torch.cos(a)), -1)}
        labels.append(lab)
        if case["near"]:
            for k in range(m):
                for rep in range(2):                   # two perturbed copies per box: a true positive and a duplicate
                    n = (3 * k + rep) % N
                    out["center"][b, n] = lab["gt_center"][k] + rnd(3) * (0.15 + 0.6 * rep)
                    out["size"][b, n] = lab["gt_size"][k] * (1 + rnd(3) * 0.05)
                    out["angle"][b, n] = lab["gt_angle"][k] + rnd(2) * 0.05
                    out["class"][b, n] = rnd(C) * 0.1
                    out["class"][b, n, int(cls[k])] += 3.0 - rep
    return out, labels


def main():
    ref = reference_shim.import_reference_metric(criterion.box3d_overlap)
    metric = ref.build_metric(EVALUATE)
    recs = []
    for case in CASES:
        out, labels = make_case(case)
        per_sample = ref.Metric.from_config({**EVALUATE, "reduction": "none"})(out, labels)
        mean = metric(out, labels)
        recs.append(dict(case=case, per_sample={k: v.clone() for k, v in per_sample.items()}, mean={k: v.clone() for k, v in mean.items()}))
        print(case["name"], {k: [round(float(x), 4) for x in v] for k, v in per_sample.items()})
    path = os.path.join(HERE, "..", "tests", "golden", "metrics_small.pt")
    torch.save({"cases": recs, "evaluate": EVALUATE, "torch_version": torch.__version__,
                "overlap": "dpft_b200.criterion.box3d_overlap (pytorch3d absent)"}, path)
    print(os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
