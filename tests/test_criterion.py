"""dpft_b200.criterion (assigner + set criterion, batched) against the unmodified reference loss — a
committed fixture (tools/make_golden_criterion.py) on any machine, the live reference in the build container — and the
box-overlap restatement (pytorch3d is absent: no pin to the package itself) against closed-form cases, a Monte-Carlo estimate
and an exact, independent polyhedron intersection (half-space intersection + convex hull, scipy)."""
import math

import pytest
import torch

from conftest import load_golden
from dpft_b200 import criterion as C


def _box(cx, cy, cz, l, w, h, yaw, dtype=torch.float64):
    t = lambda *v: torch.tensor([list(v)], dtype=dtype)
    return C.get_box_corners(t(cx, cy, cz), t(l, w, h), torch.tensor([yaw], dtype=dtype))


@pytest.mark.parametrize("a,b,vol", [
    ((0, 0, 0, 2, 2, 2, 0.0), (0, 0, 0, 2, 2, 2, 0.0), 8.0),                       # identical
    ((0, 0, 0, 2, 2, 2, 0.0), (5, 0, 0, 2, 2, 2, 0.3), 0.0),                       # disjoint
    ((0, 0, 0, 2, 2, 2, 0.0), (1, 1, 1, 2, 2, 2, 0.0), 1.0),                       # axis-aligned corner overlap
    ((0, 0, 0, 4, 4, 1, 0.0), (0.5, -0.5, 0, 1, 2, 3, 0.7), 2.0),                  # b's footprint inside a's, a's z range inside b's
    ((0, 0, 0, 1, 1, 1, 0.0), (0, 0, 0, 1, 1, 1, math.pi / 4), 2 * (math.sqrt(2) - 1)),   # square vs 45 deg square: octagon
    ((0, 0, 0, 2, 2, 2, 0.0), (2, 0, 0, 2, 2, 2, 0.0), 0.0),                       # touching faces
    ((0, 0, 0, 2, 2, 2, 0.0), (0, 0, 1.5, 2, 2, 2, math.pi / 2), 2.0),             # quarter turn, partial z overlap
    ((1, 2, 3, 4, 2, 1, 0.4), (1, 2, 3, 4, 2, 1, 0.4 + math.pi), 8.0),             # half turn of the same box
])
def test_box3d_overlap_closed_forms(a, b, vol):
    for dtype, tol in ((torch.float64, 1e-9), (torch.float32, 1e-5)):
        ba, bb = _box(*a, dtype=dtype), _box(*b, dtype=dtype)
        v, iou = C.box3d_overlap(ba, bb)
        va, vb = a[3] * a[4] * a[5], b[3] * b[4] * b[5]
        assert abs(float(v) - vol) < tol * max(1.0, vol), (float(v), vol)
        assert abs(float(iou) - vol / (va + vb - vol)) < tol
        v2, iou2 = C.box3d_overlap(bb, ba)                                          # symmetric
        assert abs(float(v2) - float(v)) < tol and abs(float(iou2) - float(iou)) < tol


def test_box3d_overlap_matches_monte_carlo_volumes():
    g = torch.Generator().manual_seed(0)
    n = 6
    c1, c2 = torch.rand(n, 3, generator=g, dtype=torch.float64) * 2, torch.rand(n, 3, generator=g, dtype=torch.float64) * 2
    s1, s2 = torch.rand(n, 3, generator=g, dtype=torch.float64) * 3 + 1, torch.rand(n, 3, generator=g, dtype=torch.float64) * 3 + 1
    a1, a2 = torch.rand(n, generator=g, dtype=torch.float64) * 6.28, torch.rand(n, generator=g, dtype=torch.float64) * 6.28
    vol, iou = C.box3d_overlap(C.get_box_corners(c1, s1, a1), C.get_box_corners(c2, s2, a2))
    pts = (torch.rand(400000, 3, generator=g, dtype=torch.float64) - 0.5) * 10 + 1.0

    def inside(c, s, a):                                       # (P, n): point in the yawed box
        d = pts[:, None, :] - c[None]
        x = torch.cos(a) * d[..., 0] + torch.sin(a) * d[..., 1]
        y = -torch.sin(a) * d[..., 0] + torch.cos(a) * d[..., 1]
        return (x.abs() <= s[:, 0] / 2) & (y.abs() <= s[:, 1] / 2) & (d[..., 2].abs() <= s[:, 2] / 2)

    in1, in2 = inside(c1, s1, a1), inside(c2, s2, a2)
    mc = (in1[:, :, None] & in2[:, None, :]).double().mean(0) * 1000.0
    assert float((vol - mc).abs().max()) < 0.25, float((vol - mc).abs().max())     # ~1 % of the box volumes (Monte-Carlo noise)
    assert float(iou.min()) >= 0.0 and float(iou.max()) <= 1.0 + 1e-9


_FACES = [[0, 1, 2, 3], [3, 2, 6, 7], [0, 1, 5, 4], [0, 3, 7, 4], [1, 2, 6, 5], [4, 5, 6, 7]]


def _exact_intersection_volume(c1, c2):
    """Volume of the intersection of two convex hexahedra given by their corners, by an algorithm that shares nothing with
    ``box3d_overlap``: the twelve face half-spaces, a Chebyshev centre from a linear programme, scipy's half-space
    intersection (qhull) and the volume of the hull of its vertices.  pytorch3d's op (exact clipping of the faces of one box
    by the planes of the other) computes this same quantity for any pair of boxes."""
    import numpy as np
    from scipy.optimize import linprog
    from scipy.spatial import ConvexHull, HalfspaceIntersection

    def halfspaces(corners):
        centre, rows = corners.mean(0), []
        for face in _FACES:
            p = corners[face]
            n = np.cross(p[1] - p[0], p[2] - p[0])
            n /= np.linalg.norm(n)
            if n @ (centre - p[0]) > 0:
                n = -n                                              # outward normal: n.x + d <= 0 inside
            rows.append(np.r_[n, -n @ p[0]])
        return np.array(rows)

    hs = np.vstack([halfspaces(c1), halfspaces(c2)])
    res = linprog([0, 0, 0, -1], A_ub=np.c_[hs[:, :3], np.ones(len(hs))], b_ub=-hs[:, 3],
                  bounds=[(None, None)] * 3 + [(0, None)])
    if res.status != 0 or res.x[3] < 1e-9:
        return 0.0
    return ConvexHull(HalfspaceIntersection(hs, res.x[:3]).intersections).volume


def test_box3d_overlap_matches_exact_polyhedron_intersection():
    """144 random pairs of yawed boxes in float64: volume and IoU equal the exact 3-D polyhedron intersection to 1e-12."""
    import numpy as np
    g = torch.Generator().manual_seed(1)
    n = 12
    rnd = lambda *shape: torch.rand(*shape, generator=g, dtype=torch.float64)
    c1, c2 = rnd(n, 3) * 2, rnd(n, 3) * 2
    s1, s2 = rnd(n, 3) * 3 + 0.5, rnd(n, 3) * 3 + 0.5
    b1, b2 = C.get_box_corners(c1, s1, rnd(n) * 6.28), C.get_box_corners(c2, s2, rnd(n) * 6.28)
    vol, iou = C.box3d_overlap(b1, b2)
    exact = np.array([[_exact_intersection_volume(b1[i].numpy(), b2[j].numpy()) for j in range(n)] for i in range(n)])
    assert (exact > 0).sum() > 100                                  # most pairs do overlap
    assert float(np.abs(vol.numpy() - exact).max()) < 1e-12 * max(1.0, float(exact.max()))
    union = s1.prod(-1)[:, None].numpy() + s2.prod(-1)[None, :].numpy() - exact
    assert float(np.abs(iou.numpy() - exact / union).max()) < 1e-12


def _run(case, train_config):
    from make_golden_criterion import make_case
    out, labels = make_case(case)
    loss_fn = C.build_loss(train_config)
    leaf = {k: v.clone().requires_grad_(True) for k, v in out.items()}
    total, losses = loss_fn(leaf, labels)
    total.backward()
    tgt, mask = C.pad_targets(labels, out["class"].device, out["class"].dtype)
    i, j, _ = loss_fn.anassigner(out, tgt, mask)
    return total, losses, {k: v.grad for k, v in leaf.items()}, (i, j, mask)


def _check(got, want_total, want_losses, want_grads, want_matches):
    total, losses, grads, (i, j, mask) = got
    assert abs(float(total) - float(want_total)) < 1e-5 * max(1.0, abs(float(want_total)))
    assert set(losses) == set(want_losses)
    for k, w in want_losses.items():
        assert abs(float(losses[k]) - float(w)) < 1e-5 * max(1.0, abs(float(w))), k
    for k, w in want_grads.items():
        assert float((grads[k] - w).abs().max()) < 1e-6 * max(1.0, float(w.abs().max())), k
    for b, m in enumerate(want_matches):
        n = int(mask[b].sum())
        if m is None:
            assert n == 0
        else:
            assert torch.equal(i[b, :n], m[0]) and torch.equal(j[b, :n], m[1])


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_criterion_matches_reference_fixture(idx):
    rec = load_golden("criterion_small")["cases"][idx]
    _check(_run(rec["case"], rec["train_config"]), rec["total"], rec["losses"], rec["grads"], rec["matches"])


def test_criterion_matches_live_reference(reference_models):
    import reference_shim
    from make_golden_criterion import make_case, train_config
    ref = reference_shim.import_reference_loss(C.box3d_overlap)
    case = dict(seed=11, B=5, N=33, counts=[0, 4, 1, 9, 2], reduction="mean", degenerate=True,
                weights={"total_class": 1.0, "object_class": 0.3, "center": 1.0, "size": 2.0, "angle": 0.7})
    cfg = train_config(case)
    out, labels = make_case(case)
    leaf = {k: v.clone().requires_grad_(True) for k, v in out.items()}
    loss_fn = ref.build_loss(cfg)
    total, losses = loss_fn(leaf, labels)
    total.backward()
    matches = []
    for b, lab in enumerate(labels):
        if lab["gt_class"].shape[0] == 0:
            matches.append(None)
            continue
        i, j = loss_fn.anassigner({k: v[b:b + 1] for k, v in out.items()}, {k: v[None] for k, v in lab.items()})
        matches.append((i[0], j[0]))
    _check(_run(case, cfg), total.detach(), {k: v.detach() for k, v in losses.items()}, {k: v.grad for k, v in leaf.items()}, matches)


def test_batch_without_any_target_gives_zero_loss_and_zero_gradients():
    out = {"class": torch.randn(2, 9, 2, requires_grad=True), "center": torch.randn(2, 9, 3, requires_grad=True),
           "size": torch.rand(2, 9, 3, requires_grad=True), "angle": torch.rand(2, 9, 2, requires_grad=True)}
    empty = {"gt_class": torch.zeros(0, 2), "gt_center": torch.zeros(0, 3), "gt_size": torch.zeros(0, 3), "gt_angle": torch.zeros(0, 2)}
    loss_fn = C.build_loss({"anassigner": "HungarianAnassigner", "criterion": "SetCriterion",
                            "loss_weights": {"total_class": 1.0, "object_class": 0.0, "center": 1.0, "size": 1.0, "angle": 1.0}})
    total, losses = loss_fn(out, [empty, empty])
    assert float(total) == 0.0 and all(float(v) == 0.0 for v in losses.values())
    total.backward()
    assert all(float(v.grad.abs().max()) == 0.0 for v in out.values())


def test_unsupported_criterion_configurations_are_named():
    with pytest.raises(NotImplementedError):
        C.build_loss({"criterion": "SetCriterion"})
    with pytest.raises(NotImplementedError):
        C.build_loss({"anassigner": "HungarianAnassigner", "criterion": "Other"})


def test_criterion_drives_a_model_training_step():
    """DPRT outputs -> criterion -> backward (host logic on the CPU, CUDA op swapped for the oracle op): finite loss, gradients
    on the decoder / head / backbone parameters."""
    from helpers import oracle_op_injected
    from dpft_b200 import configs, models, synthetic
    cfg = synthetic.offline_config(configs.make_config("kradar_radar_bev"), dropout=0.0)
    model = models.build("dprt", cfg).train()
    model.load_state_dict(synthetic.seeded_state_dict(model.state_dict(), seed=3))
    batch = synthetic.synthetic_batch(cfg, 2, seed=4, sizes={"radar_bev": (32, 107, 6)})
    g = torch.Generator().manual_seed(5)
    labels = []
    for m in (2, 0):
        a = torch.rand(m, generator=g) * 6.28
        labels.append({"gt_class": torch.nn.functional.one_hot(torch.randint(0, 2, (m,), generator=g), 2).float(),
                       "gt_center": torch.rand(m, 3, generator=g) * torch.tensor([72.0, 12.8, 8.0]) - torch.tensor([0.0, 6.4, 2.0]),
                       "gt_size": torch.rand(m, 3, generator=g) * 2 + 1, "gt_angle": torch.stack((torch.sin(a), torch.cos(a)), -1)})
    loss_fn = C.build_loss({"anassigner": "HungarianAnassigner", "criterion": "SetCriterion",
                            "loss_weights": {"total_class": 1.0, "object_class": 0.0, "center": 1.0, "size": 1.0, "angle": 1.0}})
    with oracle_op_injected():
        total, losses = loss_fn(model(batch), labels)
        total.backward()
    assert torch.isfinite(total) and float(total) > 0
    grads = {k: p.grad for k, p in model.named_parameters()}
    for key in ("fuser.heads.3.layers.class_head.0.weight", "fuser.heads.3.layers.center_head.6.weight", "fuser.query",
                "backbones.radar_bev.body.conv1.weight"):
        assert grads[key] is not None and torch.isfinite(grads[key]).all() and float(grads[key].abs().max()) > 0, key
