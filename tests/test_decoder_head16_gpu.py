"""decoder_head16_kernel (sixteen lanes per query; dpft_decoder_head_forward with DPFT_HEAD_LANES16) against the validated
one-thread-per-query decoder_head_kernel: every output is accumulated in the same order, so the results must be
BIT-IDENTICAL (green on B200 since round 2, gpurun call r02_call01; the sixteen-lane kernel is the default now)."""
import pytest
import torch

from dpft_b200 import decoder as dec

pytestmark = [pytest.mark.gpu]
DEV = "cuda:0"


@pytest.mark.parametrize("reduction", [0, 1, 2])
@pytest.mark.parametrize("B,V,N,n_cls,last", [(2, 3, 300, 2, True), (1, 1, 37, 2, False), (3, 2, 400, 5, True), (8, 3, 300, 2, False)])
def test_head16_is_bit_identical_to_the_per_query_kernel(reduction, B, V, N, n_cls, last):
    g = torch.Generator(device=DEV).manual_seed(B * 1000 + N + reduction)
    views = torch.randn(B, V, N, 16, generator=g, device=DEV)
    n_w = (16 * V * 16 if reduction == 0 else 0) + 4 * 512 + (3 + 3 + 2 + n_cls) * 16
    weights = torch.randn(n_w, generator=g, device=DEV) * 0.3
    center_in = torch.randn(B, N, 3, generator=g, device=DEV) * 10
    outs = []
    for lanes16 in (False, True):
        q = torch.full((B, N, 16), float("nan"), device=DEV)
        c = torch.full((B, N, 3), float("nan"), device=DEV)
        s = torch.full((B, N, 3), float("nan"), device=DEV) if last else None
        a = torch.full((B, N, 2), float("nan"), device=DEV) if last else None
        k = torch.full((B, N, n_cls), float("nan"), device=DEV) if last else None
        dec.head_forward(views, weights, center_in, q, c, s, a, k, B, V, N, n_cls, reduction, lanes16=lanes16)
        outs.append([t for t in (q, c, s, a, k) if t is not None])
    torch.cuda.synchronize()
    for old, new in zip(*outs):
        assert not torch.isnan(new).any()
        assert torch.equal(old, new)


def test_head16_shared_centres():
    """center_in (N, 3) shared by the batch (first iteration: the static query grid)."""
    g = torch.Generator(device=DEV).manual_seed(5)
    B, V, N = 4, 3, 300
    views = torch.randn(B, V, N, 16, generator=g, device=DEV)
    weights = torch.randn(16 * V * 16 + 4 * 512 + 10 * 16, generator=g, device=DEV) * 0.3
    center_in = torch.randn(N, 3, generator=g, device=DEV)
    res = []
    for lanes16 in (False, True):
        q = torch.empty(B, N, 16, device=DEV)
        c = torch.empty(B, N, 3, device=DEV)
        dec.head_forward(views, weights, center_in, q, c, None, None, None, B, V, N, 2, 0, lanes16=lanes16)
        res.append((q, c))
    torch.cuda.synchronize()
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
