"""Adds tests/golden/mlfusion_self_attn.pt: the self-attention sub-layer of the UNMODIFIED reference decoder layer
(MLFusion.forward_self_attn, src/dprt/models/fusers/mpfusion.py:122-148) on seeded weights and inputs, for the shipped head layout
(d_model 16, 8 heads of 2 channels, 400 queries) and for a wide one (d_model 64, 4 heads of 16 channels, 257 queries).

Run in the build container only:  python tools/make_golden_selfattn.py   (does not touch the other fixtures)
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, ".."))

import reference_shim  # noqa: E402
from dpft_b200 import synthetic  # noqa: E402

GOLDEN = os.path.join(HERE, "..", "tests", "golden")


def main():
    reference_shim.import_reference_models()
    from dprt.models.fusers.mpfusion import MLFusion
    cases = []
    for i, (d_model, heads, n, b) in enumerate([(16, 8, 400, 2), (64, 4, 257, 1)]):
        torch.manual_seed(11 + i)
        layer = MLFusion(d_model=d_model, d_ffn=2 * d_model, n_levels=1, n_heads=heads, n_points=4, norm=True, dropout=0.1,
                         activation="Mish").eval()
        sd = synthetic.seeded_state_dict(layer.state_dict(), seed=300 + i)
        layer.load_state_dict(sd)
        g = torch.Generator().manual_seed(400 + i)
        x = torch.randn(b, n, d_model, generator=g)
        pos = torch.rand(b, n, d_model, generator=g)
        with torch.no_grad():
            out = layer.forward_self_attn(x, pos)
        keep = {k: v for k, v in sd.items() if k.startswith(("self_attn.", "norm1."))}
        cases.append(dict(d_model=d_model, n_heads=heads, state_dict=keep, x=x, pos=pos, out=out))
        print(d_model, heads, tuple(out.shape), float(out.abs().max()))
    torch.save(dict(cases=cases, torch_version=torch.__version__), os.path.join(GOLDEN, "mlfusion_self_attn.pt"))


if __name__ == "__main__":
    main()
