"""The DPRT training step as ONE CUDA graph.

A native training step of the bench workload is ~1900 launches of our own kernels plus ~3000 small torch kernels (FPN,
decoder, optimiser); issued eagerly from Python the GPU idles for ~45 % of the step (65 ms of kernel time in a 116 ms
step on B200).  Every launch on the path is shape-static and free of host synchronisation, so the whole step

    zero gradient bucket -> forward -> loss -> backward (+ chunked gradient all-reduce) -> optimiser

is captured once (after eager warm-up steps on a side stream, the torch whole-network-capture recipe) and replayed with
new inputs copied into the captured input buffers.  Replaces, for the driver loop at reference
src/dprt/training/trainer.py:116-135 (``zero_grad -> model(data) -> loss -> backward -> optimizer.step``), the per-op
launch stream the reference issues.

Requirements (checked): CUDA model in ``train()``; an optimiser constructed with ``capturable=True``; batches of constant
shape/dtype.  Dropout keeps working under replay (torch's Philox offsets are graph-aware).
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch

from .ddp import GradientBucket


class GraphedTrainStep:
    """``step = GraphedTrainStep(model, bucket, optimizer, loss_fn); loss = step(batch)``.

    ``loss_fn(outputs, batch) -> scalar tensor``.  ``graph=False`` runs the same sequence eagerly (debugging, or shapes
    that change from step to step).  The returned loss is a device tensor that the next call overwrites."""

    def __init__(self, model: torch.nn.Module, bucket: GradientBucket, optimizer: torch.optim.Optimizer,
                 loss_fn: Callable[[Dict[str, torch.Tensor], Dict[str, torch.Tensor]], torch.Tensor],
                 graph: bool = True, warmup: int = 3):
        self.model, self.bucket, self.optimizer, self.loss_fn = model, bucket, optimizer, loss_fn
        self.use_graph, self.warmup = graph, max(int(warmup), 1)
        if graph:
            for group in optimizer.param_groups:
                if not group.get("capturable", False):
                    raise ValueError("GraphedTrainStep(graph=True) needs an optimizer built with capturable=True")
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._static: Optional[Dict[str, torch.Tensor]] = None
        self._loss: Optional[torch.Tensor] = None
        self._signature = None
        self.native_launches_per_step = 0

    def _eager(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        self.bucket.zero()
        out = self.model(batch)
        loss = self.loss_fn(out, batch)
        loss.backward()
        self.bucket.finish()
        self.optimizer.step()
        return loss.detach()

    @staticmethod
    def _sig(batch: Dict[str, torch.Tensor]):
        return tuple((k, tuple(v.shape), v.dtype) for k, v in batch.items())

    def _capture(self, batch: Dict[str, torch.Tensor]) -> None:
        dev = next(self.model.parameters()).device
        if dev.type != "cuda" or not self.model.training:
            raise RuntimeError("GraphedTrainStep: the model must be on a CUDA device and in train() mode")
        self._static = {k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in batch.items()}
        self._load(batch)
        # The warm-up passes exist for the allocator, lazy initialisations and the weight-packer tables — they must not train:
        # learning rate 0 (AdamW then leaves every parameter bit-identical), BatchNorm buffers snapshotted and restored in
        # place, and the optimiser moments / step counters the passes created zeroed in place afterwards, so that the first
        # replayed step is step 1 on untouched statistics, exactly as the reference loop's first iteration (trainer.py:116-135).
        groups = self.optimizer.param_groups
        saved_lr = [g["lr"] for g in groups]
        buffers = [b for b in self.model.buffers()]
        saved_buffers = [b.detach().clone() for b in buffers]
        for g in groups:
            g["lr"] = g["lr"] * 0.0                      # keeps a tensor learning rate a tensor
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                self._eager(self._static)
            with torch.no_grad():
                for b, s0 in zip(buffers, saved_buffers):
                    b.copy_(s0)
                for state in self.optimizer.state.values():
                    for v in state.values():
                        if isinstance(v, torch.Tensor):
                            v.zero_()
        for g, lr in zip(groups, saved_lr):
            g["lr"] = lr
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self._graph = torch.cuda.CUDAGraph()
        from . import native
        l0 = native.launches()
        # thread_local: other threads (the NCCL watchdog polling its events) keep calling the CUDA API during a capture that
        # now spans several streams; in the default global mode such a call invalidates the capture
        with torch.cuda.graph(self._graph, capture_error_mode="thread_local"):
            self._loss = self._eager(self._static)
        self.native_launches_per_step = native.launches() - l0      # our own kernels inside one replay
        self._signature = self._sig(batch)

    def _load(self, batch: Dict[str, torch.Tensor]) -> None:
        for k, buf in self._static.items():
            buf.copy_(batch[k], non_blocking=True)       # device->device, or H2D from pinned host memory

    def __call__(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        if not self.use_graph:
            dev = next(self.model.parameters()).device
            return self._eager({k: v.to(dev, non_blocking=True) for k, v in batch.items()})
        if self._graph is None:
            self._capture(batch)                          # warm-up steps + capture; the captured step is NOT run here
        elif self._sig(batch) != self._signature:
            raise RuntimeError("GraphedTrainStep: batch shapes/dtypes changed after capture; build a new step object "
                               "(or pass graph=False)")
        self._load(batch)
        self._graph.replay()
        return self._loss

    def release(self) -> None:
        """Drops the captured graph and its memory pool (call before tearing down the process group it captured)."""
        self._graph, self._loss, self._static, self._signature = None, None, None, None

    @property
    def warmup_steps_taken(self) -> int:
        """Optimiser updates applied by the capture warm-up: none (the warm-up passes run at learning rate 0 and their side
        effects on BatchNorm buffers and optimiser state are undone before the capture)."""
        return 0
