"""Single-node data parallelism for the DPRT training step: one process per GPU, batch sharded across ranks, ONE flat
fp32 gradient bucket all-reduced over NCCL (NVLink 5 / NVSwitch) per step.

The reference is single-GPU (``CentralizedTrainer``, src/dprt/training/trainer.py:20; no DDP/NCCL anywhere), so this is
an extension with nothing to match except the fp32 numerics of the per-rank step.  Forward work is independent per
sample — the only exchange on the path is the gradient sum:

  * parameters that never receive a gradient are left out of the bucket: ``head.*`` (registered at
    src/dprt/models/dprt.py:112 but never called) and the size/angle/class branches of the intermediate heads
    ``fuser.heads.{0..I-2}`` (only ``center`` feeds the next iteration, mpfusion.py:732-743);
  * every other parameter's ``.grad`` is a VIEW into one contiguous buffer, so backward accumulates straight into the
    bucket (no flatten copy) and the optimizer reads the reduced values in place;
  * the bucket is cut into a few contiguous chunks in backward order (fuser, necks, backbone stages); each chunk's
    all-reduce is launched asynchronously from a post-accumulate hook as soon as its last gradient lands, so the
    communication overlaps the remaining backbone backward.
"""
from __future__ import annotations

import re
from typing import Dict, Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist
from torch import nn


def unused_parameter_names(model: nn.Module) -> List[str]:
    """Names of parameters that receive no gradient in a DPRT training step (39 for the shipped configs)."""
    names = []
    n_iter = getattr(getattr(model, "fuser", None), "i_iter", 0)
    for name, _ in model.named_parameters():
        if name.startswith("head."):
            names.append(name)
            continue
        m = re.match(r"fuser\.heads\.(\d+)\.layers\.(\w+)_head\.", name)
        if m and int(m.group(1)) < n_iter - 1 and m.group(2) != "center":
            names.append(name)
    return names


def broadcast_parameters(model: nn.Module, src: int = 0, group=None) -> None:
    """Rank ``src``'s parameters and buffers to every rank (one flat broadcast per dtype)."""
    tensors = [p.data for p in model.parameters()] + [b.data for b in model.buffers()]
    by_dtype: Dict[torch.dtype, List[torch.Tensor]] = {}
    for t in tensors:
        by_dtype.setdefault(t.dtype, []).append(t)
    for ts in by_dtype.values():
        flat = torch.cat([t.reshape(-1) for t in ts])
        dist.broadcast(flat, src=src, group=group)
        off = 0
        for t in ts:
            t.copy_(flat[off:off + t.numel()].view_as(t))
            off += t.numel()


class GradientBucket:
    """Flat gradient buffer with chunked, overlapped all-reduce.

    usage:   bucket = GradientBucket(model);  loss.backward();  bucket.finish();  optimizer.step();  bucket.zero()
    """

    def __init__(self, model: nn.Module, n_chunks: int = 4, group=None, average: bool = True):
        self.group = group
        self.average = average
        self.communicate = True          # False: skip the all-reduces (bench.py times the step with and without them)
        # streams other than the current one that produce gradients (DPRT.training_streams(): the extra views' backward runs
        # on their forward streams); an all-reduce of a chunk waits for them so that it never reads a gradient still being written
        self._model = model
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        skip = set(unused_parameter_names(model))
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad and n not in skip]
        # backward produces gradients roughly in reverse registration order: put late modules first
        named.reverse()
        self.names = [n for n, _ in named]
        self.params = [p for _, p in named]
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.offsets: List[Tuple[int, int]] = []
        off = 0
        for p in self.params:
            if p.dtype != torch.float32:
                raise TypeError("GradientBucket expects fp32 master parameters")
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            self.offsets.append((off, off + p.numel()))
            off += p.numel()
        # contiguous chunks of roughly equal byte size
        n_chunks = max(1, min(n_chunks, len(self.params)))
        target = total / n_chunks
        self.chunk_of: List[int] = []
        self.chunk_bounds: List[List[int]] = []
        cur, start = 0, 0
        for i, (a, b) in enumerate(self.offsets):
            self.chunk_of.append(cur)
            last_param = i == len(self.offsets) - 1
            if (b - start >= target and cur < n_chunks - 1) or last_param:
                self.chunk_bounds.append([start, b])
                start = b
                cur += 1
        self.n_chunks = len(self.chunk_bounds)
        self.chunk_size = [0] * self.n_chunks
        for c in self.chunk_of:
            self.chunk_size[c] += 1
        self._ready = [0] * self.n_chunks
        self._handles: List[Optional[object]] = [None] * self.n_chunks
        self._hooks = [p.register_post_accumulate_grad_hook(self._make_hook(i)) for i, p in enumerate(self.params)]

    def _make_hook(self, index: int):
        chunk = self.chunk_of[index]

        def hook(param):
            # autograd may have replaced .grad with a fresh tensor on first accumulation; keep the view
            a, b = self.offsets[index]
            view = self.flat[a:b].view_as(param)
            if param.grad is not None and param.grad.data_ptr() != view.data_ptr():
                view.copy_(param.grad)
                param.grad = view
            self._ready[chunk] += 1
            if self._ready[chunk] == self.chunk_size[chunk]:
                self._launch(chunk)
        return hook

    def _launch(self, chunk: int) -> None:
        if self.world == 1 or not self.communicate or self._handles[chunk] is not None:
            return
        a, b = self.chunk_bounds[chunk]
        if hasattr(self._model, "training_streams") and self.flat.is_cuda:
            cur = torch.cuda.current_stream(self.flat.device)
            capturing = torch.cuda.is_current_stream_capturing()
            for s in self._model.training_streams():
                if s == cur:
                    continue
                if capturing:
                    # a registered stream that has not been forked into this capture yet (a backbone's weight-gradient stream
                    # before that backbone's backward starts) holds no work of this step: waiting on it would pull an
                    # un-captured dependency into the graph and invalidate the capture
                    with torch.cuda.stream(s):
                        if not torch.cuda.is_current_stream_capturing():
                            continue
                cur.wait_stream(s)
        self._handles[chunk] = dist.all_reduce(self.flat[a:b], op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish(self) -> None:
        """Call after backward: launches any chunk whose hooks did not all fire, waits, averages."""
        if self.world > 1:
            for c in range(self.n_chunks):
                self._launch(c)
            for h in self._handles:
                if h is not None:
                    h.wait()
            if self.average:
                self.flat.div_(self.world)
        self._ready = [0] * self.n_chunks
        self._handles = [None] * self.n_chunks

    def zero(self) -> None:
        """Replacement for optimizer.zero_grad(): keeps the gradient views in place."""
        self.flat.zero_()
        for (a, b), p in zip(self.offsets, self.params):
            if p.grad is None or p.grad.data_ptr() != self.flat[a:b].data_ptr():
                p.grad = self.flat[a:b].view_as(p)

    def bytes(self) -> int:
        return self.flat.numel() * 4
