"""GraphedTrainStep (dpft_b200/train_step.py): the eager sequence on the CPU against a hand-written loop (driver loop of
reference src/dprt/training/trainer.py:116-135), and on the GPU the captured CUDA graph against the eager sequence."""
import copy

import pytest
import torch

from dpft_b200 import configs, ddp, models, synthetic
from dpft_b200.train_step import GraphedTrainStep
from helpers import oracle_op_injected


def _loss(out, _batch):
    return sum((v ** 2).mean() for v in out.values())


def _small(dropout=0.0):
    cfg = synthetic.offline_config(configs.make_config("kradar_radar_bev"), n_queries=(6, 5, 1))
    cfg["model"]["fuser"]["dropout"] = dropout
    sizes = {"radar_bev": (64, 64, 6)}
    return cfg, sizes


def test_eager_sequence_equals_manual_loop_cpu():
    cfg, sizes = _small()
    torch.manual_seed(0)
    a = models.build("dprt", cfg).train()
    b = copy.deepcopy(a)
    batch = synthetic.synthetic_batch(cfg, 2, seed=5, sizes=sizes)
    with oracle_op_injected():
        bucket = ddp.GradientBucket(a, n_chunks=2)
        opt = torch.optim.AdamW(bucket.params, lr=1e-3)
        step = GraphedTrainStep(a, bucket, opt, _loss, graph=False)
        losses = [float(step(batch)) for _ in range(2)]
        skip = set(ddp.unused_parameter_names(b))
        opt_b = torch.optim.AdamW([p for n, p in reversed(list(b.named_parameters())) if n not in skip], lr=1e-3)
        ref = []
        for _ in range(2):
            opt_b.zero_grad()
            l = _loss(b(batch), batch)
            l.backward()
            opt_b.step()
            ref.append(float(l))
    assert losses == pytest.approx(ref, rel=1e-6)
    for (n, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
        assert torch.allclose(p, q, rtol=1e-5, atol=1e-7), n


def test_graph_needs_capturable_optimizer():
    cfg, _ = _small()
    m = models.build("dprt", cfg).train()
    with oracle_op_injected():
        bucket = ddp.GradientBucket(m, n_chunks=2)
    with pytest.raises(ValueError, match="capturable"):
        GraphedTrainStep(m, bucket, torch.optim.AdamW(bucket.params, lr=1e-3), _loss, graph=True)


@pytest.mark.gpu
@pytest.mark.parametrize("native_train", [True, False])
def test_graph_replay_matches_eager_gpu(native_train):
    """Same model, same batches: five steps replayed from one CUDA graph vs five eager steps.  The steps are identical launch
    sequences; the only difference allowed is the order of fp32 atomic accumulation in the split-K weight gradients and
    the deformable-attention scatter (tolerance 2e-3 on the loss after five AdamW updates, lr 1e-3)."""
    dev = "cuda:0"
    cfg, sizes = _small()
    torch.manual_seed(0)
    base = models.build("dprt", cfg)
    base.load_state_dict(synthetic.seeded_state_dict(base.state_dict(), seed=1))
    batches = [synthetic.synthetic_batch(cfg, 2, seed=10 + i, sizes=sizes, device=dev) for i in range(2)]
    results = []
    for graph in (False, True):
        m = copy.deepcopy(base).to(dev).train()
        m.native_train = native_train
        bucket = ddp.GradientBucket(m, n_chunks=3)
        opt = torch.optim.AdamW(bucket.params, lr=1e-3, capturable=True)
        step = GraphedTrainStep(m, bucket, opt, _loss, graph=graph, warmup=3)
        losses = []
        # (the capture warm-up runs at learning rate 0 and undoes its side effects: the first replay IS step 1, as in eager mode)
        for i in range(5):
            losses.append(float(step(batches[i % 2]).clone()))
        torch.cuda.synchronize()
        if graph:
            assert step.warmup_steps_taken == 0
            assert (step.native_launches_per_step > 0) and step._graph is not None
        results.append((losses, {n: p.detach().clone() for n, p in m.named_parameters()},
                        {n: b.detach().clone() for n, b in m.named_buffers()}))
    (l_e, p_e, b_e), (l_g, p_g, b_g) = results
    assert l_g == pytest.approx(l_e, rel=2e-3), (l_e, l_g)
    assert l_g[0] != l_g[-1]                              # the replayed optimiser really updates the weights
    for n in b_e:
        if n.endswith("num_batches_tracked"):
            assert int(b_e[n]) == int(b_g[n]) == 5, n     # five steps each: the warm-up passes leave no trace
    with pytest.raises(RuntimeError, match="shapes"):
        step({k: v[:1] for k, v in batches[0].items()})


@pytest.mark.gpu
@pytest.mark.parametrize("native_train", [True, False])
def test_forked_training_views_equal_the_single_stream_step(native_train):
    """train(): the extra views' backbone + neck (forward and, because autograd runs a node's backward on its forward's stream,
    backward) on forked streams must give the loss and the gradients of the single-stream step — same kernels, same order per
    view; only fp32 atomics in the weight gradients / deformable-attention scatter may reorder (bound: 1e-4 of each tensor's
    largest entry).  Dropout is left at the config's value: the masks depend on host program order, which does not change."""
    dev = "cuda:0"
    cfg = synthetic.offline_config(configs.make_config("kradar"), n_queries=(6, 5, 1))          # three views, dropout as shipped
    sizes = {"camera_mono": (96, 160, 3), "radar_bev": (64, 64, 6), "radar_front": (64, 64, 6)}
    torch.manual_seed(0)
    base = models.build("dprt", cfg)
    base.load_state_dict(synthetic.seeded_state_dict(base.state_dict(), seed=1))
    batch = synthetic.synthetic_batch(cfg, 2, seed=11, sizes=sizes, device=dev)
    old_tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    res = []
    from dpft_b200 import streams
    calls = {"n": 0}
    real_fork_map = streams.fork_map

    def counting_fork_map(fns, device, **kw):
        calls["n"] += 1
        return real_fork_map(fns, device, **kw)

    streams.fork_map = counting_fork_map
    try:
        for forked in (False, False, True, True):       # two runs each: the serial pair gives the noise floor of the atomics
            m = copy.deepcopy(base).to(dev).train()
            m.native_train, m.train_parallel_views = native_train, forked
            torch.manual_seed(123)
            torch.cuda.manual_seed(123)
            loss = _loss(m(batch), batch)
            loss.backward()
            torch.cuda.synchronize()
            res.append((float(loss.detach()), {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}))
            # the feature extraction and each of the four decoder iterations fork when the switch is on, nothing forks when off
            assert calls["n"] == (0 if not forked else 5 * (len(res) - 2)), (forked, calls["n"])
    finally:
        streams.fork_map = real_fork_map
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old_tf32

    def worst(a, b):
        return max(float((a[n] - b[n]).abs().max()) / (float(a[n].abs().max()) + 1e-12) for n in a)

    (l0, g0), (l0b, g0b), (l1, g1), (l1b, g1b) = res
    floor = worst(g0, g0b)
    assert abs(l0 - l1) <= 1e-5 * abs(l0) and abs(l0 - l1b) <= 1e-5 * abs(l0), (l0, l1, l1b)
    assert g0.keys() == g1.keys()
    for g in (g1, g1b):
        w = worst(g0, g)
        assert w <= max(3.0 * floor, 1e-4), (w, floor)
