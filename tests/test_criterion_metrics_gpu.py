"""dpft_b200.criterion / dpft_b200.metrics on cuda tensors against the fixtures of the unmodified reference (the CPU tests
hold the same modules to them on the host).  Green on B200 since round 2 (gpurun call r02_call01)."""
import pytest
import torch

from conftest import load_golden
from dpft_b200 import criterion, metrics

pytestmark = [pytest.mark.gpu]
DEV = "cuda:0"


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_criterion_on_cuda_matches_reference_fixture(idx):
    from make_golden_criterion import make_case
    rec = load_golden("criterion_small")["cases"][idx]
    out, labels = make_case(rec["case"])
    leaf = {k: v.to(DEV).requires_grad_(True) for k, v in out.items()}
    total, losses = criterion.build_loss(rec["train_config"])(leaf, [{k: v.to(DEV) for k, v in l.items()} for l in labels])
    total.backward()
    assert abs(float(total) - float(rec["total"])) < 1e-4 * max(1.0, abs(float(rec["total"])))
    for k, w in rec["losses"].items():
        assert abs(float(losses[k]) - float(w)) < 1e-4 * max(1.0, abs(float(w))), k
    for k, w in rec["grads"].items():
        assert float((leaf[k].grad.cpu() - w).abs().max()) < 1e-5 * max(1.0, float(w.abs().max())), k


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_metrics_on_cuda_match_reference_fixture(idx):
    from make_golden_metrics import make_case
    rec = load_golden("metrics_small")
    c = rec["cases"][idx]
    out, labels = make_case(c["case"])
    got = metrics.Metric.from_config({**rec["evaluate"], "reduction": "none"})(
        {k: v.to(DEV) for k, v in out.items()}, [{k: v.to(DEV) for k, v in l.items()} for l in labels])
    for k, w in c["per_sample"].items():
        assert float((got[k].cpu() - w).abs().max()) < 1e-4, (k, got[k], w)


def test_lsap_kernel_matches_scipy():
    """dpft_lsap_forward (one warp per sample, lsap_core.h with LANES = 32) against scipy on random cost tensors."""
    import numpy as np
    from scipy.optimize import linear_sum_assignment
    rng = np.random.default_rng(3)
    for B, N, Mmax in ((6, 300, 12), (3, 37, 37), (8, 400, 64), (1, 5, 1)):
        counts = rng.integers(0, Mmax + 1, size=B).astype(np.int32)
        counts[0] = Mmax
        cost = rng.standard_normal((B, N, Mmax)).astype(np.float32)
        out = criterion.lsap_device(torch.from_numpy(cost).to(DEV), torch.from_numpy(counts).to(DEV)).cpu().numpy()
        for b, m in enumerate(counts):
            assert (out[b, m:] == -1).all()
            if m:
                rows, cols = linear_sum_assignment(cost[b, :, :m].astype(np.float64))
                match = dict(zip(cols.tolist(), rows.tolist()))
                assert all(match[i] == int(out[b, i]) for i in range(m)), (B, N, Mmax, b)


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_criterion_with_the_device_solver_matches_reference_fixture(idx):
    from make_golden_criterion import make_case
    rec = load_golden("criterion_small")["cases"][idx]
    out, labels = make_case(rec["case"])
    leaf = {k: v.to(DEV).requires_grad_(True) for k, v in out.items()}
    loss_fn = criterion.build_loss({**rec["train_config"], "lsap_solver": "device"})
    total, losses = loss_fn(leaf, [{k: v.to(DEV) for k, v in l.items()} for l in labels])
    total.backward()
    assert abs(float(total) - float(rec["total"])) < 1e-4 * max(1.0, abs(float(rec["total"])))
    for k, w in rec["grads"].items():
        assert float((leaf[k].grad.cpu() - w).abs().max()) < 1e-5 * max(1.0, float(w.abs().max())), k
