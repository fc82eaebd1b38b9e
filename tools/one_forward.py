"""N eager (no CUDA graph, serial streams) eval forwards of the bench workload — the target process of the ncu passes
that list every launch of one step (profiles/*_launches*.csv) and sum the DRAM traffic of the convolution kernel."""
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from dpft_b200 import configs, models, synthetic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = "cuda:0"
cfg = synthetic.offline_config(configs.make_config("kradar"), n_queries=(20, 15, 1))
m = models.build("dprt", cfg).eval()
m.load_state_dict(synthetic.seeded_state_dict(m.state_dict(), seed=1))
m = m.to(dev)
m.use_cuda_graph, m.parallel_views = False, False
batch = synthetic.synthetic_batch(cfg, 8, seed=1000, sizes=dict(synthetic.BASELINE_SIZES), device=dev)
with torch.no_grad():
    for i in range(n):
        if i == n - 1:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()          # ncu --profile-from-start off: only the last forward is profiled
        out = m(batch)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print({k: tuple(v.shape) for k, v in out.items()})
