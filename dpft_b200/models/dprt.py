"""DPRT model assembly (mirror of reference src/dprt/models/dprt.py): per-input backbone -> skip link (raw
input as level '0') -> FPN neck -> sinusoidal embedding, then querent -> fuser(+heads).

``DPRT.from_config(config)`` reads the same JSON keys as the reference (``config['computing']`` merged into
every sub-module config, dprt.py:35) and ``forward(batch)`` takes / returns the same dictionaries
(dprt.py:200-244).  In ``eval()`` on a CUDA device the forward runs through the fused sm_100a pipeline
(dpft_b200/engine.py) when the configuration is eligible; otherwise module by module on the GPU with the
deformable-attention op from libdpft_b200.so.
"""
from __future__ import annotations

import os

from collections import OrderedDict
from typing import Any, Callable, Dict, List, Optional, Tuple

import torch
from torch import nn

from .backbone import build_backbone
from .embedding import build_embedding
from .fuser import build_fuser
from .head import build_head
from .neck import build_neck
from .querent import build_querent


def _build(build_fn: Callable, section: Optional[Dict[str, Any]], computing: Dict[str, Any], **kwargs):
    if section is None:
        return None
    return build_fn(section["name"], dict(computing | section), **kwargs)


def _build_each(build_fn: Callable, sections: Optional[Dict[str, Any]], computing: Dict[str, Any]):
    if sections is None:
        return None
    return {k: _build(build_fn, v, computing) for k, v in sections.items()}


class DPRT(nn.Module):
    def __init__(self, inputs: List[str], skiplinks: Dict[str, bool] = None, backbones: Dict[str, nn.Module] = None,
                 necks: Dict[str, nn.Module] = None, embeddings: Dict[str, nn.Module] = None,
                 querent: nn.Module = None, fuser: nn.Module = None, head: nn.Module = None, **kwargs):
        super().__init__()
        self.inputs = list(inputs)
        skiplinks = skiplinks or {}
        self.skiplinks = {i: bool(skiplinks.get(i, False)) for i in self.inputs}

        def per_input(mods):
            mods = mods or {}
            return nn.ModuleDict({i: (mods.get(i) if mods.get(i) is not None else nn.Identity()) for i in self.inputs})

        self.backbones = per_input(backbones)
        self.necks = per_input(necks)
        self.embeddings = per_input(embeddings)
        self.querent = querent if querent is not None else nn.Identity()
        self.fuser = fuser if fuser is not None else nn.Identity()
        self.head = head if head is not None else nn.Identity()   # registered but unused, like the reference (dprt.py:112)
        self.use_fused = True          # eval-mode fused pipeline switch (tests flip it to compare the two paths)
        self.native_features = True    # fused pipeline: 16-bit tcgen05 backbone/FPN (True) or torch fp32 features (False)
        self.feature_dtype = torch.float16    # activation type of the native backbone: torch.float16 (10-bit mantissa,
                                              # outputs saturate at +-65504) or torch.bfloat16
        self.pyramid_dtype = torch.float16    # storage of the native (B, S, 16) feature pyramid: float16 or float32
        self.use_cuda_graph = True     # fused pipeline: replay a captured graph once an input shape repeats
        self.parallel_views = True     # fused pipeline: run the per-view feature extractors on forked streams
        self.side_view_priority = False     # ... the other views on high-priority streams (measured: no gain)
        # ... and with their persistent convolution grids capped at this many CTAs (0 = one per SM).  A conv CTA owns its SM
        # (~200 KB of shared memory); the radar layers are latency-bound, so 148 CTAs buy them nothing and starve the camera's
        # kernels of the neighbouring forwards: bench step 4.64 -> 4.49 ms at 48, one forward at a time unchanged (4.74 -> 4.72);
        # 16 is best pipelined (4.45) but makes the radar the critical path of a single forward (5.12)
        # (profiles/r02_side_view_ctas_ab.txt); 64 since the two radar views share launches (below)
        self.side_view_ctas = int(os.environ.get("DPFT_SIDE_VIEW_CTAS", "64"))
        # two side views of the same architecture on equally shaped inputs share their Bottleneck launches (conv2d_nhwc_pair;
        # the cap above then covers both): 52 launches fewer per forward, the serial conv time of a step 5.33 -> 4.70 ms, the
        # pipelined step within the run-to-run spread (4.52 -> 4.47 ms), one forward at a time +0.5 % (profiles/r02_pair_side_views_ab.txt)
        self.pair_side_views = os.environ.get("DPFT_PAIR_SIDE_VIEWS", "1") == "1"
        self.native_train = True       # train() on CUDA: ResNet stages through the sm_100a training kernels (16-bit
                                       # activations, dpft_b200/train_backbone.py); False = torch/cuDNN autograd in fp32
        self.train_dtype = torch.float16
        # train(): the extra views' backbones + necks on forked streams (their backward follows them there); False = one stream
        self.train_parallel_views = os.environ.get("DPFT_TRAIN_PARALLEL_VIEWS", "1") == "1"
        self._engine = None

    @classmethod
    def from_config(cls, config: Dict[str, Any]) -> "DPRT":
        computing, model = config["computing"], config["model"]
        head = _build(build_head, model.get("head"), computing)
        fuser = _build(build_fuser, model.get("fuser"), computing, head=head)
        return cls(inputs=model.get("inputs"), skiplinks=model.get("skiplinks"),
                   backbones=_build_each(build_backbone, model.get("backbones"), computing),
                   necks=_build_each(build_neck, model.get("necks"), computing),
                   embeddings=_build_each(build_embedding, model.get("embeddings"), computing),
                   querent=_build(build_querent, model.get("querent"), computing),
                   fuser=fuser, head=head)

    def __getstate__(self):            # torch.save(model) (reference trainer.py:258) must not pickle device caches
        state = self.__dict__.copy()
        state["_engine"] = None
        return state

    # -- composed (module by module) path -------------------------------------------------------------------
    def _view_features(self, name: str, batch: Dict[str, torch.Tensor]) -> "OrderedDict[str, torch.Tensor]":
        bb = self.backbones[name]
        if hasattr(bb, "native_train"):
            bb.native_train, bb.train_dtype = self.native_train, self.train_dtype
        f = bb(batch[name])
        if self.skiplinks[name]:
            f = OrderedDict([("0", batch[name])] + list(f.items()))
        f = self.necks[name](f)
        return self.embeddings[name](f)

    def extract_features(self, batch: Dict[str, torch.Tensor], only=None) -> Dict[str, "OrderedDict[str, torch.Tensor]"]:
        names = list(self.inputs if only is None else only)
        x0 = batch[names[0]] if names else None
        from ..streams import single_process
        if (self.training and self.train_parallel_views and len(names) > 1 and x0 is not None and x0.is_cuda
                and torch.is_grad_enabled() and single_process()):
            return self._extract_features_forked(batch, names)
        return {name: self._view_features(name, batch) for name in names}

    def _extract_features_forked(self, batch, names) -> Dict[str, "OrderedDict[str, torch.Tensor]"]:
        """Training: the views are independent until the fuser, and autograd runs every backward node on the stream its forward
        ran on — so forward AND backward of the small radar backbones (2 x ~50 layers of latency-bound training kernels) overlap
        the camera's instead of queueing behind them (dpft_b200/streams.py)."""
        from ..streams import fork_map
        feats = fork_map([(lambda n=name: self._view_features(n, batch)) for name in names], batch[names[0]].device,
                         reads=[batch[name] for name in names])
        return dict(zip(names, feats))

    def training_streams(self):
        """Side streams the forked parts of a training step run on (GradientBucket waits for them before an all-reduce)."""
        from ..streams import registered
        p = next(self.parameters(), None)
        return registered(p.device) if p is not None and p.is_cuda else []

    def forward_composed(self, batch: Dict[str, torch.Tensor]) -> "OrderedDict[str, torch.Tensor]":
        from ..streams import single_process
        for mp in getattr(self.fuser, "mpfusion", {}).values():       # the per-view decoder layers follow the same switch
            mp.train_parallel_views = self.train_parallel_views and single_process()
        feats = self.extract_features(batch)
        out = self.querent(batch)
        return self.fuser(batch=[feats[i] for i in self.inputs],
                          shape=[batch[f"{i}_shape"][:, :2] for i in self.inputs],
                          projection=[(batch[f"label_to_{i}_t"], batch[f"label_to_{i}_p"]) for i in self.inputs],
                          out=out)

    # -- dispatch -------------------------------------------------------------------------------------------
    def forward(self, batch: Dict[str, torch.Tensor]) -> "OrderedDict[str, torch.Tensor]":
        if self.use_fused and not self.training and not torch.is_grad_enabled():
            from ..engine import FusedEngine
            if self._engine is None:
                self._engine = FusedEngine.try_create(self)
            if self._engine is not None and self._engine.accepts(batch):
                return self._engine.forward(batch)
        if any(batch[name].dtype == torch.uint8 for name in self.inputs):      # uint8 frames outside the native feature path:
            batch = {k: (v.float() if k in self.inputs and v.dtype == torch.uint8 else v) for k, v in batch.items()}   # plumbing
            if self.use_fused and not self.training and not torch.is_grad_enabled() and self._engine is not None \
                    and self._engine.accepts(batch):
                return self._engine.forward(batch)
        return self.forward_composed(batch)

    def infer_stream(self, batches, depth: int = 2):
        """Inference over a sequence of equally shaped batches, yielding each batch's output dictionary in order.  On the
        fused sm_100a pipeline up to ``depth`` forwards are in flight at once (``FusedEngine.stream``: one captured graph,
        memory pool and stream per slot), which is how a serving / evaluation loop (reference
        src/dprt/evaluation/evaluator.py:120-135) should drive the model; otherwise the batches run one after the other."""
        from ..engine import FusedEngine
        it = iter(batches)
        try:
            first = next(it)
        except StopIteration:
            return
        import itertools
        chain = itertools.chain([first], it)
        if self.use_fused and not self.training and self.use_cuda_graph:
            if self._engine is None:
                self._engine = FusedEngine.try_create(self)
            if self._engine is not None and self._engine.accepts(first):
                yield from self._engine.stream(chain, depth)
                return
        for batch in chain:
            with torch.no_grad():                    # (not held across the yield: grad mode is the caller's between items)
                out = self(batch)
            yield out


    def stream_input_slots(self, example: Dict[str, torch.Tensor], depth: int = 2):
        """Device input buffers of the ``depth`` pipeline slots of ``infer_stream`` (``FusedEngine.stream_slots``): fill
        ``slots[i % depth]`` in place with batch i and pass that dictionary as the i-th item of ``infer_stream(..., depth)`` to
        skip the input staging copy.  ``None`` when the fused pipeline does not serve this model / batch."""
        from ..engine import FusedEngine
        if not (self.use_fused and not self.training and self.use_cuda_graph):
            return None
        if self._engine is None:
            self._engine = FusedEngine.try_create(self)
        if self._engine is None or not self._engine.accepts(example):
            return None
        return self._engine.stream_slots(example, depth)


def build_dprt(config: Dict[str, Any], *args, **kwargs) -> DPRT:
    return DPRT.from_config(config)
