"""Intermediate quantities of a DPRT forward ("taps", SURVEY.md §8c's golden-vector list), collected the same way from the
unmodified reference model and from this repo's model: both expose the same module tree (embeddings.<input>,
fuser.mpfusion.fusion<i>.ml_fusion_layers.ms_deform_attn<v>.ms_deform_attn, fuser.heads[i]) and a
``fuser.get_reference_points`` callable, so forward hooks at those points see the same tensors.

  features_<input>      per-level FPN + positional-embedding maps (B, H, W, C), finest first      (dprt.py:228-231)
  ref_points_<it>       per view: normalised reference points (B, N, 2)                           (mpfusion.py:617-696)
  query_<it>            fused query state after iteration it (B, N, C)                            (mpfusion.py:740)
  center_<it>           refined box centres after iteration it (B, N, 3)                          (detection.py:273)
  msda_out_<it>_<v>     output of the MSDeformAttn module of view v in iteration it (B, N, C)     (ms_deform_attn.py:215)
"""
from collections import OrderedDict

import torch


def collect(model, batch, n_iter: int, n_views: int):
    """Runs ``model(batch)`` with hooks attached and returns (outputs, taps).  Works on CPU or GPU models."""
    taps = OrderedDict()
    handles = []
    fuser = model.fuser

    for name in model.inputs:
        def feat_hook(_m, _inp, out, name=name):
            taps[f"features_{name}"] = [t.detach().clone() for t in out.values()]
        handles.append(model.embeddings[name].register_forward_hook(feat_hook))

    refs = []
    orig = fuser.get_reference_points

    def ref_wrapper(*a, **kw):
        r = orig(*a, **kw)
        refs.append(r.detach().clone())
        return r

    fuser.get_reference_points = ref_wrapper          # instance attribute shadows the (static)method for this call
    for it in range(n_iter):
        def q_hook(_m, _inp, out, it=it):
            taps[f"query_{it}"] = out.detach().clone()
        handles.append(fuser.mpfusion[f"fusion{it}"].register_forward_hook(q_hook))

        def c_hook(_m, _inp, out, it=it):
            taps[f"center_{it}"] = out["center"].detach().clone()
        handles.append(fuser.heads[it].register_forward_hook(c_hook))
        for v in range(n_views):
            mod = fuser.mpfusion[f"fusion{it}"].ml_fusion_layers[f"ms_deform_attn{v}"].ms_deform_attn

            def m_hook(_m, _inp, out, it=it, v=v):
                taps[f"msda_out_{it}_{v}"] = out.detach().clone()
            handles.append(mod.register_forward_hook(m_hook))
    try:
        with torch.no_grad():
            out = model(batch)
    finally:
        for h in handles:
            h.remove()
        del fuser.get_reference_points
    assert len(refs) == n_iter * n_views
    for it in range(n_iter):
        taps[f"ref_points_{it}"] = refs[it * n_views:(it + 1) * n_views]
    return out, taps


LEVEL0_STRIDE = (5, 7)        # the finest level is stored subsampled (rows ::5, columns ::7) plus its mean and L2 norm


def compress_features(levels):
    """Fixture form of one view's feature list: full tensors for the coarse levels, a strided subsample + two moments for
    the finest one (keeps tests/golden small)."""
    f0 = levels[0]
    return {"level0_sub": f0[:, ::LEVEL0_STRIDE[0], ::LEVEL0_STRIDE[1]].clone(), "level0_mean": float(f0.double().mean()),
            "level0_norm": float(f0.double().norm()), "shapes": [tuple(t.shape) for t in levels],
            "coarse": [t.clone() for t in levels[1:]]}


def feature_errors(levels, rec):
    """Largest relative deviation of a feature list from its fixture form."""
    from_sub = levels[0][:, ::LEVEL0_STRIDE[0], ::LEVEL0_STRIDE[1]].cpu()
    errs = [float((from_sub.double() - rec["level0_sub"].double()).abs().max() / rec["level0_sub"].double().abs().max()),
            abs(float(levels[0].double().norm()) - rec["level0_norm"]) / rec["level0_norm"]]
    assert [tuple(t.shape) for t in levels] == rec["shapes"]
    for t, w in zip(levels[1:], rec["coarse"]):
        errs.append(float((t.cpu().double() - w.double()).abs().max() / w.double().abs().max()))
    return max(errs)


def gradient_digest(named_grads, n_samples: int = 16):
    """Fingerprint of a gradient set (90 M gradients do not fit a fixture): per parameter its L2 norm, its sum and
    n_samples entries at indices drawn from a generator seeded by the parameter name; packed into four tensors.
    Parameters without a gradient are listed under 'none'."""
    import zlib
    names, none, norms, sums, idxs, vals = [], [], [], [], [], []
    for k, g in named_grads.items():
        if g is None:
            none.append(k)
            continue
        flat = g.detach().reshape(-1).cpu()
        gen = torch.Generator().manual_seed(zlib.crc32(k.encode()))
        idx = torch.randint(0, flat.numel(), (n_samples,), generator=gen)
        names.append(k)
        norms.append(float(flat.double().norm()))
        sums.append(float(flat.double().sum()))
        idxs.append(idx)
        vals.append(flat[idx])
    return {"names": names, "none": none, "norm": torch.tensor(norms, dtype=torch.float64),
            "sum": torch.tensor(sums, dtype=torch.float64), "idx": torch.stack(idxs), "values": torch.stack(vals)}


def digest_errors(named_grads, rec, details=None):
    """(worst relative norm error, worst sampled-entry error) against a digest.  The entry error of a parameter is the largest
    |got - want| over its sampled entries relative to max(gradient RMS, largest sampled |want|): a max-norm figure like
    tests/helpers.py::rel_err (weight gradients are heavy-tailed: an entry of 30 x RMS carrying a 3 % error is 0.9 RMS).
    ``details``: a list that receives one (name, norm error, entry error, rms, largest sampled |want|) row per parameter."""
    assert sorted(k for k, g in named_grads.items() if g is None) == sorted(rec["none"])
    worst_norm, worst_val = 0.0, 0.0
    for i, k in enumerate(rec["names"]):
        flat = named_grads[k].detach().reshape(-1).cpu()
        n = float(flat.double().norm())
        e_norm = abs(n - float(rec["norm"][i])) / max(float(rec["norm"][i]), 1e-30)
        rms = float(rec["norm"][i]) / flat.numel() ** 0.5
        want = rec["values"][i].double()
        d = float((flat[rec["idx"][i]].double() - want).abs().max())
        scale = max(rms, float(want.abs().max()), 1e-30)
        e_val = d / scale
        worst_norm, worst_val = max(worst_norm, e_norm), max(worst_val, e_val)
        if details is not None:
            details.append((k, e_norm, e_val, rms, float(want.abs().max())))
    return worst_norm, worst_val
