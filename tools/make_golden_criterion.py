"""Generates tests/golden/criterion_small.pt: seeded predictions / labels pushed through the UNMODIFIED reference loss
(dprt.training.loss.Loss with HungarianAnassigner + SetCriterion, built from the shipped config's 'train' section).

pytorch3d is not installed in the build container: dpft_b200.criterion.box3d_overlap stands in for
pytorch3d.ops.box3d_overlap (tools/reference_shim.import_reference_loss) — the fixture pins everything AROUND that function
(cost matrix, GIoU bookkeeping, validity masks, assignment, focal / L1 terms, weights, reductions), not the overlap itself.

  python tools/make_golden_criterion.py
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, ".."))

import reference_shim  # noqa: E402
from dpft_b200 import criterion  # noqa: E402

CASES = [
    dict(name="shipped_weights", seed=1, B=4, N=60, counts=[3, 0, 5, 1], weights=None, reduction="mean", degenerate=False),
    dict(name="all_terms", seed=2, B=3, N=40, counts=[2, 7, 4], reduction="mean", degenerate=True,
         weights={"total_class": 1.0, "object_class": 0.5, "center": 2.0, "size": 0.25, "angle": 1.5}),
    dict(name="sum_reduction", seed=3, B=2, N=25, counts=[6, 6], reduction="sum", degenerate=False,
         weights={"total_class": 0.5, "object_class": 1.0, "center": 1.0, "size": 1.0, "angle": 1.0}),
]


def make_case(case):
    """Predictions in the value ranges of the detection head (size >= 0 through a ReLU, angle through tanh) and labels as
    KRadarDataset.get_detection_label makes them (one-hot class incl. the background class, (sin, cos) angle)."""
    g = torch.Generator().manual_seed(case["seed"])
    B, N = case["B"], case["N"]
    rnd = lambda *s: torch.randn(*s, generator=g)
    out = {"class": rnd(B, N, 2), "center": torch.stack((torch.rand(B, N, generator=g) * 72, rnd(B, N) * 6, rnd(B, N)), -1),
           "size": torch.relu(rnd(B, N, 3) + 2), "angle": torch.tanh(rnd(B, N, 2))}
    if case["degenerate"]:
        out["size"][:, ::7] = 0.0                      # zero-volume predictions: masked out of the GIoU cost (giou = -1)
    labels = []
    for m in case["counts"]:
        a = torch.rand(m, generator=g) * 6.28
        labels.append({"gt_class": torch.nn.functional.one_hot(torch.randint(0, 2, (m,), generator=g), 2).float(),
                       "gt_center": torch.stack((torch.rand(m, generator=g) * 72, rnd(m) * 6, rnd(m)), -1),
                       "gt_size": torch.rand(m, 3, generator=g) * 3 + 1, "gt_angle": torch.stack((torch.sin(a), torch.cos(a)), -1)})
    return out, labels


def train_config(case):
    with open("/root/reference/config/kradar.json") as f:
        cfg = json.load(f)["train"]
    if case["weights"] is not None:
        cfg["loss_weights"] = dict(case["weights"])
    cfg["reduction"] = case["reduction"]
    return cfg


def main():
    ref = reference_shim.import_reference_loss(criterion.box3d_overlap)
    recs = []
    for case in CASES:
        out, labels = make_case(case)
        cfg = train_config(case)
        loss_fn = ref.build_loss(cfg)
        leaf = {k: v.clone().requires_grad_(True) for k, v in out.items()}
        total, losses = loss_fn(leaf, labels)
        total.backward()
        matches = []
        for b, lab in enumerate(labels):                 # the reference's assignment per sample (assigner.py:58-143)
            if lab["gt_class"].shape[0] == 0:
                matches.append(None)
                continue
            i, j = loss_fn.anassigner({k: v[b:b + 1] for k, v in out.items()}, {k: v[None] for k, v in lab.items()})
            matches.append((i[0].clone(), j[0].clone()))
        recs.append(dict(case=case, train_config=cfg, total=total.detach().clone(), losses={k: v.detach().clone() for k, v in losses.items()},
                         grads={k: v.grad.clone() for k, v in leaf.items()}, matches=matches))
        print(case["name"], float(total), {k: round(float(v), 5) for k, v in losses.items()})
    path = os.path.join(HERE, "..", "tests", "golden", "criterion_small.pt")
    torch.save({"cases": recs, "torch_version": torch.__version__, "overlap": "dpft_b200.criterion.box3d_overlap (pytorch3d absent)"}, path)
    print(os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
