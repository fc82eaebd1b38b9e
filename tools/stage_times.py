"""Where the native forward spends its time on the bench workload (CUDA events, mean of 5 after warm-up)."""
import json
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from dpft_b200 import configs, features, models, synthetic  # noqa: E402

dev = "cuda:0"
small = "--small" in sys.argv
cfg = synthetic.offline_config(configs.make_config("kradar"), n_queries=(20, 15, 1))
sizes = dict(synthetic.BASELINE_SIZES)
if small:
    sizes = {"camera_mono": (96, 160, 3), "radar_bev": (64, 64, 6), "radar_front": (64, 64, 6)}
model = models.build("dprt", cfg).eval()
model.load_state_dict(synthetic.seeded_state_dict(model.state_dict(), seed=1))
model = model.to(dev)
batch = synthetic.synthetic_batch(cfg, 8, seed=1, sizes=sizes, device=dev)


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3     # us


with torch.no_grad():
    model.use_cuda_graph = False
    model(batch)
    eng = model._engine
    res = {}
    for name, nv in zip(model.inputs, eng.views):
        x = batch[name].contiguous()
        stem = features.stem_forward(x, nv.stem_w, nv.stem_b, nv.dtype, w_packed=nv.stem_w_packed)
        res[f"{name}.stem"] = timeit(lambda: features.stem_forward(x, nv.stem_w, nv.stem_b, nv.dtype, w_packed=nv.stem_w_packed))
        pooled = features.maxpool_forward(stem)
        res[f"{name}.maxpool"] = timeit(lambda: features.maxpool_forward(stem))
        y = pooled
        for si, blocks in enumerate(nv.stages):
            def run_stage(y0=y, blocks=blocks):
                t = y0
                for c1, c2, c3, ds in blocks:
                    idn = ds(t, relu=False) if ds is not None else t
                    t = c3(c2(c1(t, relu=True), relu=True), relu=True, residual=idn)
                return t
            res[f"{name}.stage{si + 1}"] = timeit(run_stage)
            y = run_stage()
        res[f"{name}.backbone"] = timeit(lambda: nv.backbone(x))
        res[f"{name}.pyramid_total"] = timeit(lambda: nv.pyramid(x))
    pyrs = eng.pyramids(batch)
    res["decoder"] = timeit(lambda: eng.decode(batch, pyrs))
    res["forward_eager"] = timeit(lambda: eng.forward_eager(batch))
    model.use_cuda_graph = True
    model(batch); model(batch)
    res["forward_graph"] = timeit(lambda: model(batch))
print(json.dumps({k: round(v, 1) for k, v in res.items()}))
