"""Shared test helpers: rebuild golden cases from their recipe; inject the CPU oracle op for CPU-only checks."""
import contextlib

import torch

from dpft_b200 import configs, synthetic


def case_setup(rec):
    """(cfg, batch) of a golden record, rebuilt from seeds."""
    case = rec["case"]
    cfg = synthetic.offline_config(configs.make_config(case["config"]), n_queries=case["n_queries"],
                                   multi_scale=case.get("multi_scale"), d_model=case.get("d_model"))
    batch = synthetic.synthetic_batch(cfg, case["batch"], seed=rec["input_seed"], sizes=case["sizes"])
    return cfg, batch


def rel_err(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-12))


@contextlib.contextmanager
def oracle_op_injected():
    """TEST ONLY: lets the product's host logic run on the CPU by swapping the CUDA op for the oracle."""
    import dpft_b200.msda as msda_ops
    from oracle import msda as O

    class OracleFn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, value, shapes, lsi, loc, attn, step=64):
            ctx.save_for_backward(value, shapes, loc, attn)
            return O.msda_forward_torch(value, shapes, loc, attn)

        @staticmethod
        def backward(ctx, g):
            value, shapes, loc, attn = ctx.saved_tensors
            gv, gl, ga = O.msda_backward_torch(value, shapes, loc, attn, g)
            return gv, None, None, gl, ga, None

    saved = msda_ops.MSDeformAttnFunction
    msda_ops.MSDeformAttnFunction = OracleFn
    try:
        yield
    finally:
        msda_ops.MSDeformAttnFunction = saved
