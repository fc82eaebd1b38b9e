"""World-size-2 gloo test (CPU) of the flat-bucket gradient all-reduce: the averaged per-rank gradients of two batch
shards equal the single-process gradients of the whole batch."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _build():
    from dpft_b200 import configs, models, synthetic
    cfg = synthetic.offline_config(configs.make_config("kradar_radar_front"), dropout=0.0)
    model = models.build("dprt", cfg)
    model.load_state_dict(synthetic.seeded_state_dict(model.state_dict(), seed=5))
    model.eval()        # BatchNorm on running statistics: samples are independent, so shards average exactly
    for p in model.parameters():
        p.requires_grad_(True)
    batch = synthetic.synthetic_batch(cfg, 4, seed=6, sizes={"radar_front": (37, 40, 6)})
    return model, batch


def _loss(out):
    return sum((v ** 2).mean() for v in out.values())


def _worker(rank, world, port, result_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from helpers import oracle_op_injected
    from dpft_b200 import ddp, streams
    assert not streams.single_process()            # data parallel: the training step keeps the single-stream schedule
    model, batch = _build()
    if rank == 1:                                   # rank 1 starts from different weights: broadcast must fix that
        with torch.no_grad():
            for p in model.parameters():
                p.add_(1.0)
    ddp.broadcast_parameters(model, src=0)
    bucket = ddp.GradientBucket(model, n_chunks=3)
    shard = {k: v[rank * 2:(rank + 1) * 2] for k, v in batch.items()}
    with oracle_op_injected():
        _loss(model(shard)).backward()
    bucket.finish()
    if rank == 0:
        torch.save({n: p.grad.clone() for n, p in zip(bucket.names, bucket.params)}, result_path)
    # a second step must keep working on the same views
    bucket.zero()
    assert float(bucket.flat.abs().max()) == 0.0
    with oracle_op_injected():
        _loss(model(shard)).backward()
    bucket.finish()
    assert float(bucket.flat.abs().max()) > 0.0
    dist.barrier()
    dist.destroy_process_group()


def test_flat_bucket_allreduce_matches_full_batch(tmp_path):
    from helpers import oracle_op_injected
    from dpft_b200 import ddp
    result = str(tmp_path / "grads.pt")
    mp.spawn(_worker, args=(2, _free_port(), result), nprocs=2, join=True)
    got = torch.load(result)
    model, batch = _build()
    with oracle_op_injected():
        _loss(model(batch)).backward()
    want = {n: p.grad for n, p in model.named_parameters() if p.grad is not None}
    assert set(got) == set(want)                    # exactly the parameters that receive gradients
    assert len([n for n, _ in model.named_parameters()]) - len(want) == 39 == len(ddp.unused_parameter_names(model))
    for n, g in want.items():
        scale = float(g.abs().max()) + 1e-12
        assert float((got[n] - g).abs().max()) <= 2e-4 * scale + 1e-7, n


def test_bucket_layout_single_process():
    from dpft_b200 import ddp, streams
    assert streams.single_process()                # no process group: forked training streams are allowed
    model, _ = _build()
    bucket = ddp.GradientBucket(model, n_chunks=4)
    assert bucket.n_chunks == 4 and bucket.chunk_bounds[0][0] == 0 and bucket.chunk_bounds[-1][1] == bucket.flat.numel()
    assert all(a[1] == b[0] for a, b in zip(bucket.chunk_bounds, bucket.chunk_bounds[1:]))
    assert all(p.grad.data_ptr() == bucket.flat[a:b].data_ptr() for p, (a, b) in zip(bucket.params, bucket.offsets))
    assert bucket.names[0].startswith("fuser.") and bucket.names[-1].startswith("backbones.")
