#!/bin/bash
# Round 2, call 18: forked training views against the single-stream step (loss + every gradient), eager and graphed.
O=gpurun_out/r02c18; mkdir -p $O
timeout 600 python -m pytest tests/test_train_step.py -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -15
python - <<'PY'
# the same comparison on the bench workload (full sizes, one eager step each): loss and gradient norms
import copy, torch, sys
sys.path.insert(0, '.')
from dpft_b200 import configs, models, synthetic
cfg = synthetic.offline_config(configs.make_config("kradar"), n_queries=(20, 15, 1))
base = models.build("dprt", cfg)
base.load_state_dict(synthetic.seeded_state_dict(base.state_dict(), seed=1))
batch = synthetic.synthetic_batch(cfg, 8, seed=42, sizes=dict(synthetic.BASELINE_SIZES), device="cuda:0")
out = {}
for forked in (False, True, False, True):
    m = copy.deepcopy(base).to("cuda:0").train()
    m.train_parallel_views = forked
    torch.manual_seed(5); torch.cuda.manual_seed(5)
    loss = sum((v ** 2).mean() for v in m(batch).values())
    loss.backward()
    torch.cuda.synchronize()
    gn = torch.stack([p.grad.float().norm() for p in m.parameters() if p.grad is not None]).norm()
    print('forked', forked, 'loss', float(loss), 'grad norm', float(gn))
    del m
    torch.cuda.empty_cache()
PY
