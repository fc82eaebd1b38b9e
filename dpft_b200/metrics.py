"""Detection metrics of the reference without host synchronisation.

``dprt.evaluation.metric.Metric`` (src/dprt/evaluation/metric.py:233-330) is called on EVERY training step
(``training/trainer.py:136``) and evaluates, per sample, ``mAP3D`` (:16-134) and ``mGIoU3D`` (:137-230) with Python branches on
device values (``if npos == 0``, ``selection.any()``, boolean-mask assignments): a chain of device->host round trips on the
critical path of the step.  The restatement below computes the same numbers with tensor selects only — nothing reads a
device value on the host, so the metric kernels queue behind the step and the result is fetched when it is logged.

Conventions kept as the reference has them: predictions are ranked by their RAW class score; a ground-truth box is matched to
the highest-ranked prediction of its class with IoU > threshold; the precision/recall curve is "interpolated" by the straight
line through its first and last point (``utils/misc.py:43-84`` — not a piecewise interpolation); the smallest class id that
occurs in a sample is treated as background and left out of the mean; a sample without any foreground class scores 1.

Box overlaps come from ``dpft_b200.criterion`` (``box3d_overlap`` restates the absent ``pytorch3d`` op: not pinned to the package
itself, equal to an exact independent polyhedron intersection, see there); everything else is pinned against the unmodified reference metric code (tests/test_metrics.py).
"""
from __future__ import annotations

from typing import Any, Dict, List

import torch
from torch import nn

from .criterion import box3d_overlap, get_box_corners, giou3d, valid_boxes


def _corners(pred: Dict[str, torch.Tensor], tgt: Dict[str, torch.Tensor]):
    angle = torch.atan2(pred["angle"][..., 0], pred["angle"][..., 1])
    gt_angle = torch.atan2(tgt["gt_angle"][..., 0], tgt["gt_angle"][..., 1])
    return get_box_corners(pred["center"], pred["size"], angle), get_box_corners(tgt["gt_center"], tgt["gt_size"], gt_angle)


def _foreground_mean(per_class: torch.Tensor, label: torch.Tensor, gt_label: torch.Tensor) -> torch.Tensor:
    """Mean over the classes present in the sample except the smallest one (metric.py:126-133); 1 if none is left."""
    C = per_class.shape[0]
    ids = torch.arange(C, device=per_class.device)
    present = (label[None, :] == ids[:, None]).any(1) | (gt_label[None, :] == ids[:, None]).any(1)
    first = torch.where(present, ids, torch.full_like(ids, C)).min()
    sel = (present & (ids != first)).to(per_class.dtype)
    n = sel.sum()
    return torch.where(n > 0, (per_class * sel).sum() / n.clamp_min(1), torch.ones_like(n))


def map3d(pred: Dict[str, torch.Tensor], tgt: Dict[str, torch.Tensor], threshold: float = 0.5, nelem: int = 101) -> torch.Tensor:
    """One sample: pred class (N, C), center (N, 3), size (N, 3), angle (N, 2); tgt gt_* (M, ...) -> scalar mAP (metric.py:31-134)."""
    scores = pred["class"]
    N, C = scores.shape
    M = tgt["gt_class"].shape[0]
    dev, dt = scores.device, torch.float32
    label, gt_label = scores.argmax(-1), tgt["gt_class"].argmax(-1)
    corners, gt_corners = _corners(pred, tgt)
    if M:
        _, iou = box3d_overlap(corners, gt_corners)
        iou = torch.where(valid_boxes(corners)[:, None] & valid_boxes(gt_corners)[None, :], iou, torch.zeros_like(iou))
    else:
        iou = scores.new_zeros((N, 0))
    rec_x = torch.linspace(0, 1, nelem, dtype=dt, device=dev)
    aps = []
    for l in range(C):
        mask, gt_mask = label == l, gt_label == l
        order = torch.argsort(scores[:, l], descending=True)
        mask_s = mask[order]
        hit = (iou[order] > threshold) & mask_s[:, None] & gt_mask[None, :]                  # (N, M)
        rank = torch.arange(N, device=dev)
        first = torch.where(hit, rank[:, None], torch.full_like(rank, N)[:, None]).min(0)[0] if M else rank.new_zeros((0,))
        tp = torch.zeros(N + 1, dtype=dt, device=dev)
        tp[first] = 1.0                                                                      # columns without a hit land in slot N
        tp = tp[:N]
        fp = (mask_s & (tp == 0)).to(dt)
        tp, fp = torch.cumsum(tp, 0), torch.cumsum(fp, 0)
        den = fp + tp
        prec = torch.where(den != 0, tp / torch.where(den != 0, den, torch.ones_like(den)), torch.zeros_like(tp))
        npos = gt_mask.sum().to(dt)
        rec = torch.where(npos == 0, torch.ones_like(tp), tp / npos.clamp_min(1))
        # the reference's "interp" (misc.py:43-84): the straight line through the first and the last point of the curve
        x0, x1, y0, y1 = rec[0], rec[-1], prec[0], prec[-1]
        flat = torch.isclose(x1 - x0, torch.zeros_like(x0))
        y = torch.where(flat, torch.zeros_like(rec_x), y0 + (rec_x - x0) * (y1 - y0) / torch.where(flat, torch.ones_like(x0), x1 - x0))
        y = torch.where(rec_x < x0, y0, y)
        y = torch.where(rec_x > x1, torch.zeros_like(y), y)
        aps.append(torch.sum(y * 1 / (nelem - 1)))
    return _foreground_mean(torch.stack(aps), label, gt_label)


def mgiou3d(pred: Dict[str, torch.Tensor], tgt: Dict[str, torch.Tensor]) -> torch.Tensor:
    """One sample -> scalar mean generalised IoU of the best prediction per ground-truth box (metric.py:144-230)."""
    scores = pred["class"]
    N, C = scores.shape
    M = tgt["gt_class"].shape[0]
    dt = torch.float32
    label, gt_label = scores.argmax(-1), tgt["gt_class"].argmax(-1)
    corners, gt_corners = _corners(pred, tgt)
    g_all = giou3d(corners[None], gt_corners[None])[0] if M else scores.new_zeros((N, 0))    # invalid boxes: -1
    out = []
    for l in range(C):
        mask, gt_mask = label == l, gt_label == l
        pair = mask[:, None] & gt_mask[None, :]
        g = torch.where(pair, g_all, -torch.ones_like(g_all))
        none = -torch.ones((), dtype=dt, device=scores.device)
        val = torch.where(gt_mask.sum() == 0, torch.ones_like(none), none)
        if M:
            val = torch.where(pair.any(), g.max(0)[0].mean().to(dt), val)
        out.append(val)
    return _foreground_mean(torch.stack(out), label, gt_label)


class Metric(nn.Module):
    """``metrics = Metric.from_config(config['evaluate'])(outputs, labels)`` (metric.py:233-330): per-sample metrics, batch
    reduction ('mean' | 'sum' | 'none'); ``labels`` is the list of per-sample label dictionaries.  No host synchronisation."""

    FUNCTIONS = {"mAP3D": map3d, "mGIoU3D": mgiou3d}

    def __init__(self, metrics: Dict[str, str] = None, reduction: str = "mean", **kwargs):
        super().__init__()
        if reduction not in {"none", "mean", "sum"}:
            raise ValueError(f"Invalid Value for arg 'reduction': '{reduction}'\n Supported reduction modes: 'none', 'mean', 'sum'")
        self.metrics = dict(metrics) if metrics is not None else {}
        unknown = [v for v in self.metrics.values() if v not in self.FUNCTIONS]
        if unknown:
            raise NotImplementedError(f"metrics {unknown} are outside this path (mAP3D and mGIoU3D are built)")
        self.reduction = reduction

    @classmethod
    def from_config(cls, config: Dict[str, Any]) -> "Metric":
        return cls(metrics=config.get("metrics"), reduction=config.get("reduction", "mean"))

    @torch.no_grad()
    def forward(self, inputs: Dict[str, torch.Tensor], targets: List[Dict[str, torch.Tensor]]):
        if not self.metrics:
            return torch.ones(1)
        keys = ("class", "center", "size", "angle")
        per_sample = {name: [] for name in self.metrics}
        for b, target in enumerate(targets):
            pred = {k: inputs[k][b] for k in keys}
            for name, fn in self.metrics.items():
                per_sample[name].append(self.FUNCTIONS[fn](pred, target))
        out = {k: torch.stack(v) for k, v in per_sample.items()}
        if self.reduction != "none":
            out = {k: getattr(torch, self.reduction)(v) for k, v in out.items()}
        return out


def build_metric(config: Dict[str, Any]) -> Metric:
    return Metric.from_config(config)
