"""Native training path of the ResNet stages (dpft_b200/train_backbone.py) vs the same module through torch autograd in
fp32 (TF32 off): stage outputs, every parameter gradient, BatchNorm running statistics.

Floating point, bf16 activations and activation gradients through ~50 conv+BN layers each way: outputs within 3e-2 of the
fp32 maximum, parameter gradients within 1.5e-1 relative L2 with cosine similarity > 0.98 per tensor (measured values are
printed; the fp32 accumulation itself is held to 2e-5 by tests/test_train_ops_gpu.py)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _loss(feats):
    return sum((v.float() ** 2).mean() for v in feats.values())


@pytest.mark.parametrize("arch,cin,size", [("resnet50", 3, (2, 128, 160)), ("resnet50", 6, (3, 96, 72)), ("resnet101", 3, (2, 64, 96))])
def test_native_stages_match_torch_autograd(arch, cin, size):
    from dpft_b200.models.backbone import Backbone
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(3)
    ref = Backbone(arch, in_channels=cin, multi_scale=4).to(DEV).train()
    for m in ref.modules():                                   # non-trivial affine parameters
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data.uniform_(0.5, 1.5)
            m.bias.data.normal_(0, 0.1)
    nat = copy.deepcopy(ref)
    nat.native_train = True
    B, H, W = size
    x = torch.rand(B, H, W, cin, device=DEV) * 255

    out_ref = ref(x)
    _loss(out_ref).backward()
    out_nat = nat(x)
    assert nat._stages is not None, "the native training plan was not built"
    _loss(out_nat).backward()
    torch.cuda.synchronize()

    for k in out_ref:
        assert out_nat[k].shape == out_ref[k].shape and out_nat[k].dtype == torch.float32
        err = (out_nat[k] - out_ref[k]).abs().max().item() / out_ref[k].abs().max().item()
        assert err < 3e-2, (k, err)
    worst_l2, worst_cos = 0.0, 1.0
    ref_p = dict(ref.named_parameters())
    for name, p in nat.named_parameters():
        g, w = p.grad, ref_p[name].grad
        assert (g is None) == (w is None), name
        if g is None:
            continue
        assert g.shape == w.shape
        l2 = float((g.double() - w.double()).norm() / w.double().norm().clamp_min(1e-30))
        cos = float(torch.nn.functional.cosine_similarity(g.double().flatten(), w.double().flatten(), dim=0))
        worst_l2, worst_cos = max(worst_l2, l2), min(worst_cos, cos)
        assert l2 < 1.5e-1 and cos > 0.98, (name, l2, cos)
    print(f"{arch} cin={cin}: worst relative L2 {worst_l2:.4f}, worst cosine {worst_cos:.5f}")
    ref_b = dict(ref.named_buffers())
    for name, b in nat.named_buffers():
        if name.endswith("num_batches_tracked"):
            assert int(b) == int(ref_b[name]) == 1, name
        else:
            assert torch.allclose(b, ref_b[name], rtol=5e-2, atol=5e-3), name


def test_native_stages_second_step_sees_updated_weights():
    """The 16-bit operand copies are refreshed after an optimiser step (one launch), and a second backward accumulates."""
    from dpft_b200.models.backbone import Backbone
    torch.manual_seed(4)
    m = Backbone("resnet50", in_channels=3, multi_scale=2).to(DEV).train()
    m.native_train = True
    opt = torch.optim.SGD(m.parameters(), lr=1e-2)
    x = torch.rand(2, 64, 64, 3, device=DEV) * 255
    l0 = _loss(m(x))
    l0.backward()
    g1 = m.body.layer1[0].conv1.weight.grad.clone()
    opt.step()
    w_after = m.body.layer1[0].conv1.weight.detach().clone()
    l1 = _loss(m(x))
    packed = m._stages.packer.fwd[m._stages.blocks[0][0].idx]
    assert torch.equal(packed, w_after.permute(0, 2, 3, 1).to(torch.bfloat16))
    l1.backward()                                            # .grad was not cleared: accumulates
    g2 = m.body.layer1[0].conv1.weight.grad
    assert torch.isfinite(g2).all() and not torch.equal(g1, g2)
    assert float(l1) != float(l0)
