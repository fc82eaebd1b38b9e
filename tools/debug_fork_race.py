"""Diagnostic: gradients of one train-mode step (torch fp32 dense layers, two radar views) with the training views forked,
against the single-stream step, repeated; which variant (feature extraction forked / decoder layers forked / device-wide
synchronise after backward) removes a mismatch?"""
import copy
import sys

import torch

sys.path.insert(0, ".")
from dpft_b200 import configs, models, synthetic  # noqa: E402

dev = "cuda:0"
cfg = synthetic.offline_config(configs.make_config("kradar_radar"), dropout=0.0)
sizes = {"radar_bev": (64, 40, 6), "radar_front": (37, 40, 6)}
batch = {k: v.to(dev) for k, v in synthetic.synthetic_batch(cfg, 2, seed=5, sizes=sizes).items()}
base = models.build("dprt", cfg).train()
base.load_state_dict(synthetic.seeded_state_dict(base.state_dict(), seed=6))
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def run(fork_feats, fork_fuser, sync):
    m = copy.deepcopy(base).to(dev).train()
    m.native_train = False
    m.train_parallel_views = fork_feats
    out = None
    if fork_fuser != fork_feats:                       # decouple the two switches for the experiment
        real = m.forward_composed

        def patched(b):
            for mp in m.fuser.mpfusion.values():
                mp.train_parallel_views = fork_fuser
            feats = m.extract_features(b)
            o = m.querent(b)
            return m.fuser(batch=[feats[i] for i in m.inputs], shape=[b[f"{i}_shape"][:, :2] for i in m.inputs],
                           projection=[(b[f"label_to_{i}_t"], b[f"label_to_{i}_p"]) for i in m.inputs], out=o)
        out = patched(batch)
    else:
        out = m(batch)
    sum((v ** 2).mean() for v in out.values()).backward()
    if sync:
        torch.cuda.synchronize()
    return {n: p.grad.detach().cpu().double() for n, p in m.named_parameters() if p.grad is not None}


ref = run(False, False, True)
for name, args in (("serial", (False, False, False)), ("feats forked", (True, False, False)), ("fuser forked", (False, True, False)),
                   ("both forked", (True, True, False)), ("both forked + synchronize", (True, True, True))):
    bad = {}
    for rep in range(12):
        g = run(*args)
        for n in ref:
            e = float((g[n] - ref[n]).norm() / ref[n].norm().clamp_min(1e-12))
            if e > 1e-3:
                bad[n] = bad.get(n, 0) + 1
    print(name, "mismatching parameters over 12 runs:", dict(sorted(bad.items(), key=lambda kv: -kv[1])[:6]), flush=True)
