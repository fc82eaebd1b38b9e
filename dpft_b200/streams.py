"""Forked CUDA streams for the training step.

Autograd runs every backward node on the stream its forward ran on, so work that is forked in the forward — the extra views'
backbones and necks, the per-view decoder layers of an iteration — is forked in the backward as well: the latency-bound
radar kernels and the ~100 small torch kernels of a decoder layer overlap the camera's instead of queueing behind them.
Everything here is plumbing around torch.cuda streams / events; it is capturable (fork and join are events of the capturing
stream) and the gradient bucket waits for the registered streams before it all-reduces a chunk (dpft_b200/ddp.py).
"""
from __future__ import annotations

import os
from typing import Callable, Dict, List, Sequence

import torch

DEFAULT = os.environ.get("DPFT_TRAIN_PARALLEL_VIEWS", "1") == "1"


def single_process() -> bool:
    """Forked training streams are used in a single process only.  Under data parallelism the captured step also holds the NCCL
    all-reduces of the gradient bucket, and with forked streams on top it did not finish at N = 8 (bench.py's watchdog printed the
    line without `train`) and hung once in three runs at N = 2 with the weight-gradient streams — not understood yet, so the
    multi-process step stays on the single-stream schedule that is validated at N = 2 / 4 / 8."""
    import torch.distributed as dist
    return not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1)


_pool: Dict[torch.device, List[torch.cuda.Stream]] = {}
_extra: Dict[torch.device, List[torch.cuda.Stream]] = {}


def side_streams(device, n: int) -> List[torch.cuda.Stream]:
    dev = torch.device(device)
    pool = _pool.setdefault(dev, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=dev))
    return pool[:n]


def new_stream(device) -> torch.cuda.Stream:
    """A registered stream of its own (not one of the fork_map slots): e.g. the weight-gradient stream of a backbone."""
    dev = torch.device(device)
    s = torch.cuda.Stream(device=dev)
    _extra.setdefault(dev, []).append(s)
    return s


def registered(device) -> List[torch.cuda.Stream]:
    """Every side stream handed out on this device so far."""
    return list(_pool.get(torch.device(device), [])) + list(_extra.get(torch.device(device), []))


def _record(obj, stream) -> None:
    if isinstance(obj, torch.Tensor):
        if obj.is_cuda:
            obj.record_stream(stream)
    elif isinstance(obj, dict):
        for v in obj.values():
            _record(v, stream)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            _record(v, stream)


def fork_map(fns: Sequence[Callable[[], object]], device, reads=()) -> List[object]:
    """Results of ``fn()`` for every fn: the first on the current stream, the others on side streams forked from it and joined
    back before returning.  Host issue order is the list order (random-number offsets are assigned in that order, so dropout
    masks do not depend on whether the work is forked).

    ``reads``: every tensor (nested lists / dicts allowed) that was allocated OUTSIDE the branches and is read inside them.
    The branches — and later their backward nodes, on the same side streams — read these after the host has moved on;
    autograd frees a saved tensor the moment the last node that saved it has been ISSUED, and without ``record_stream`` the
    caching allocator would hand its block straight back to the allocating stream while a side stream still reads it (seen:
    the weight gradient of a forked view's first layer, which reads the input batch, wrong in 2 of 3 runs of a long-lived
    process and never in a fresh one)."""
    dev = torch.device(device)
    main = torch.cuda.current_stream(dev)
    sides = side_streams(dev, len(fns) - 1)
    fork = torch.cuda.Event()
    fork.record(main)
    results: List[object] = [None] * len(fns)
    joins = []
    for k, fn in enumerate(fns):
        if k == 0:
            results[0] = fn()
            continue
        side = sides[k - 1]
        side.wait_event(fork)
        _record(reads, side)
        with torch.cuda.stream(side):
            results[k] = fn()
            done = torch.cuda.Event()
            done.record(side)
        _record(results[k], main)
        joins.append(done)
    for done in joins:
        main.wait_event(done)
    return results
