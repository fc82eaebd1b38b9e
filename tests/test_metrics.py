"""dpft_b200.metrics (mAP3D / mGIoU3D without host synchronisation) against the unmodified reference
metrics — committed fixture (tools/make_golden_metrics.py) and, in the build container, the live reference on more cases."""
import pytest
import torch

from conftest import load_golden
from dpft_b200 import criterion, metrics


def _compare(case, want_per_sample, want_mean, evaluate):
    from make_golden_metrics import make_case
    out, labels = make_case(case)
    per_sample = metrics.Metric.from_config({**evaluate, "reduction": "none"})(out, labels)
    mean = metrics.build_metric(evaluate)(out, labels)
    assert set(per_sample) == set(want_per_sample) == set(mean)
    for k, w in want_per_sample.items():
        assert per_sample[k].shape == w.shape
        assert float((per_sample[k] - w).abs().max()) < 1e-5, (k, per_sample[k], w)
        assert abs(float(mean[k]) - float(want_mean[k])) < 1e-5, k


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_metrics_match_reference_fixture(idx):
    rec = load_golden("metrics_small")
    c = rec["cases"][idx]
    _compare(c["case"], c["per_sample"], c["mean"], rec["evaluate"])


def test_metrics_match_live_reference(reference_models):
    import reference_shim
    from make_golden_metrics import EVALUATE, make_case
    ref = reference_shim.import_reference_metric(criterion.box3d_overlap)
    for seed in range(30, 36):
        case = dict(seed=seed, B=3, N=24 + seed % 5, counts=[(seed * 3) % 7, 1 + seed % 4, 5], C=2 + seed % 2, near=seed % 3 != 0)
        out, labels = make_case(case)
        want = ref.Metric.from_config({**EVALUATE, "reduction": "none"})(out, labels)
        _compare(case, want, ref.build_metric(EVALUATE)(out, labels), EVALUATE)


def test_empty_metric_set_and_unknown_names():
    assert torch.equal(metrics.Metric()({}, []), torch.ones(1))
    with pytest.raises(NotImplementedError):
        metrics.Metric({"x": "mAP2D"})
