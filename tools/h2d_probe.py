"""Host->device copy bandwidth of this box from pinned memory (one 113.6 MB batch = the e2e leg's bytes per step), alone and
while the GPU is busy with the bench forward — tells whether `e2e` is bound by the PCIe link or by the pipeline's structure."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
dev = "cuda:0"
n = 113_642_816
host = torch.empty(n, dtype=torch.uint8).pin_memory()
devbuf = torch.empty(n, dtype=torch.uint8, device=dev)
s = torch.cuda.Stream()
res = {}
for label, chunks in (("one_copy", 1), ("12_chunks", 12)):
    ts = []
    for _ in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(s):
            e0.record()
            step = n // chunks
            for c in range(chunks):
                devbuf[c * step:(c + 1) * step].copy_(host[c * step:(c + 1) * step], non_blocking=True)
            e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    res[label + "_GBps"] = round(n / min(ts) / 1e6, 1)
# while a large device-side copy loop keeps HBM busy
a = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
b = torch.empty_like(a)
for _ in range(50):
    b.copy_(a)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(s):
    e0.record()
    devbuf.copy_(host, non_blocking=True)
    e1.record()
e1.synchronize()
torch.cuda.synchronize()
res["under_hbm_load_GBps"] = round(n / e0.elapsed_time(e1) / 1e6, 1)
print(json.dumps(res))
