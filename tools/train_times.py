"""Where the native training step of one backbone spends its time (bench workload: camera ResNet-101 at bs 8, 1280x720):
CUDA-event time of forward and backward, and the per-kernel totals from torch.profiler (CUPTI)."""
import json
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from dpft_b200.models.backbone import Backbone  # noqa: E402

dev = "cuda:0"
arch = "resnet101"
B, H, W = 8, 720, 1280
if "--small" in sys.argv:
    B, H, W = 2, 192, 256
if "--radar" in sys.argv:
    arch, H, W = "resnet50", 256, 256
native = "--torch" not in sys.argv
cin = 6 if "--radar" in sys.argv else 3
torch.manual_seed(0)
if "--full" in sys.argv:
    # the whole DPRT training step of the bench workload (three views + FPN + decoder), loss = sum_k mean(out_k^2)
    from dpft_b200 import configs, models, synthetic
    cfg = synthetic.offline_config(configs.make_config("kradar"), n_queries=(20, 15, 1))
    m = models.build("dprt", cfg)
    m.load_state_dict(synthetic.seeded_state_dict(m.state_dict(), seed=1))
    m = m.to(dev).train()
    m.native_train = native
    x = synthetic.synthetic_batch(cfg, B, seed=1, sizes=dict(synthetic.BASELINE_SIZES), device=dev)
    arch = "dprt-kradar"
else:
    m = Backbone(arch, in_channels=cin, multi_scale=4).to(dev).train()
    m.native_train = native
    x = torch.rand(B, H, W, cin, device=dev) * 255


def loss_of(out):
    return sum((v.float() ** 2).mean() for v in out.values())


def step():
    for p in m.parameters():
        p.grad = None
    out = m(x)
    loss = loss_of(out)
    loss.backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
res = {"arch": arch, "native": native, "shape": [B, H, W, cin]}
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
fw, bw = [], []
for _ in range(5):
    for p in m.parameters():
        p.grad = None
    e[0].record()
    out = m(x)
    loss = loss_of(out)
    e[1].record()
    loss.backward()
    e[2].record()
    torch.cuda.synchronize()
    fw.append(e[0].elapsed_time(e[1]))
    bw.append(e[1].elapsed_time(e[2]))
res["forward_ms"] = sum(fw) / len(fw)
res["backward_ms"] = sum(bw) / len(bw)
res["peak_mem_GB"] = torch.cuda.max_memory_allocated() / 1e9
print(json.dumps(res))

from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
rows = []
for ev in prof.key_averages():
    t = getattr(ev, "device_time_total", None)
    if t is None:
        t = getattr(ev, "cuda_time_total", 0)
    if t and ev.device_type is not None and "DeviceType.CUDA" in str(ev.device_type):
        rows.append((t, ev.count, ev.key))
rows.sort(reverse=True)
n_rows = 45 if '--full' in sys.argv else 25
tot = sum(r[0] for r in rows)
print(f"kernel time total {tot / 1e3:.2f} ms")
for t, n, k in rows[:n_rows]:
    print(f"{t / 1e3:9.3f} ms  {100 * t / tot:5.1f}%  x{n:<5d} {k[:110]}")
