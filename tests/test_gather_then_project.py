"""Host logic of the composed decoder: sampling the unprojected features and projecting afterwards
(MSDeformAttn._gather_then_project) equals the reference order value_proj -> sample (ms_deform_attn.py:172, 206-214) in the
output and in every gradient, including samples whose bilinear corners fall outside the map (the bias term).  fp64 on the
CPU with the oracle op standing in for the CUDA op; tolerance 1e-12."""
import pytest
import torch

from helpers import oracle_op_injected
from dpft_b200.models.fuser import MSDeformAttn


@pytest.mark.parametrize("shapes", [[(7, 9), (4, 5), (2, 3)], [(1, 1)], [(37, 11), (2, 4)]])
def test_gather_then_project_equals_project_then_gather(shapes):
    torch.manual_seed(len(shapes))
    L = len(shapes)
    m = MSDeformAttn(16, L, 8, 4).double()
    m.value_proj.bias.data.normal_()
    m.sampling_offsets.weight.data.normal_(0, 0.5)
    m.attention_weights.weight.data.normal_()
    S = sum(h * w for h, w in shapes)
    x = torch.randn(2, S, 16, dtype=torch.double, requires_grad=True)
    q = torch.randn(2, 11, 16, dtype=torch.double, requires_grad=True)
    ref = (torch.rand(2, 11, L, 2, dtype=torch.double) * 1.4 - 0.2).requires_grad_(True)      # some points out of bounds
    sh = torch.tensor(shapes)
    sizes = [h * w for h, w in shapes]
    lsi = torch.tensor([sum(sizes[:i]) for i in range(L)])
    res = {}
    with oracle_op_injected():
        for flag in (True, False):
            m.gather_then_project = flag
            for t in (x, q, ref):
                t.grad = None
            m.zero_grad()
            out = m(q, ref, x, sh, lsi)
            (out ** 2).sum().backward()
            res[flag] = [out.detach(), x.grad.clone(), q.grad.clone(), ref.grad.clone()] + [p.grad.clone() for p in m.parameters()]
    for a, b in zip(res[True], res[False]):
        assert (a - b).abs().max().item() <= 1e-12 * max(1.0, b.abs().max().item())


def test_padding_mask_uses_the_dense_path():
    m = MSDeformAttn(16, 1, 8, 4)
    x = torch.randn(1, 12, 16)
    mask = torch.zeros(1, 12, dtype=torch.bool)
    mask[0, 3] = True
    with oracle_op_injected():
        out = m(torch.randn(1, 5, 16), torch.rand(1, 5, 1, 2), x, torch.tensor([(3, 4)]), torch.tensor([0]), mask)
    assert out.shape == (1, 5, 16)
