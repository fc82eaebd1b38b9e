"""Fused inference pipeline behind ``DPRT.forward`` in ``eval()`` (reference src/dprt/models/dprt.py:200-244).

Stages, all on the GPU through libdpft_b200.so:
  1. per view: backbone + FPN neck + positional embedding -> one feature pyramid buffer (B, S, 16) fp32;
  2. per iteration: ``dpft_decoder_layer_forward`` (all views in one launch) and ``dpft_decoder_head_forward``.
There is no host synchronisation anywhere in the forward (the reference has >= 24).

The engine is built lazily from the module's parameters (``FusedEngine.try_create``); configurations the fused
kernels do not cover fall back to the composed GPU path in ``DPRT.forward_composed`` (never to the CPU).
"""
from __future__ import annotations

import os
from collections import OrderedDict
from typing import Dict, List, Optional

import torch

from . import decoder as dec
from . import native
from .features import NativeView
from .models.fuser import FeaturePyramid, IMPFusion
from .models.querent import DataAgnosticStaticQueries


class FusedEngine:
    def __init__(self, model):
        self.model = model
        fuser: IMPFusion = model.fuser
        self.device = fuser.query.device
        self.V, self.N, self.I = fuser.m_views, fuser.n_queries, fuser.i_iter
        layer0 = fuser.mpfusion["fusion0"].ml_fusion_layers["ms_deform_attn0"]
        self.P = layer0.ms_deform_attn.n_points
        self.d_ffn = layer0.d_ffn
        self.act = dec.ACTIVATIONS[layer0.activation]
        self.reduction = dec.REDUCTIONS[fuser.reduction]
        self.n_cls = fuser.heads[0].num_classes
        self.levels = [fuser.mpfusion["fusion0"].ml_fusion_layers[f"ms_deform_attn{v}"].ms_deform_attn.n_levels
                       for v in range(self.V)]
        # packed weights: one image per (iteration, view) and one per head
        self.layer_w: List[List[torch.Tensor]] = []
        self.head_w: List[torch.Tensor] = []
        for it in range(self.I):
            mp = fuser.mpfusion[f"fusion{it}"]
            self.layer_w.append([dec.pack_layer(mp.ml_fusion_layers[f"ms_deform_attn{v}"]).to(self.device)
                                 for v in range(self.V)])
            self.head_w.append(dec.pack_head(mp.reduction_layer, fuser.heads[it], fuser.reduction).to(self.device))
        self.feature_dtype = getattr(model, "feature_dtype", torch.float16)
        self.pyramid_dtype = getattr(model, "pyramid_dtype", torch.float16)
        # native 16-bit feature path per view where the configuration allows it (model.native_features switches it)
        self.views: List[Optional[NativeView]] = []
        for name in model.inputs:
            why = NativeView.ineligible_reason(model.backbones[name], model.necks[name], model.embeddings[name],
                                               model.skiplinks[name])
            self.views.append(NativeView(model.backbones[name], model.necks[name], model.embeddings[name],
                                         model.skiplinks[name], self.device, self.feature_dtype, self.pyramid_dtype)
                              if why is None else None)
        self.query = fuser.query.detach().float().contiguous()
        self.pos = fuser.query_embedding.weight.detach().float().contiguous()
        self._row_0001 = torch.tensor([0.0, 0.0, 0.0, 1.0], device=self.device)
        self._param_version = self._version(model)
        self._mode = self._mode_of(model)
        self._graphs: Dict[tuple, List["_CapturedForward"]] = {}
        self._pipelines: Dict[tuple, List["_PipelineSlot"]] = {}
        self._seen: Dict[tuple, int] = {}
        self._slot = 0
        self._side_streams = [torch.cuda.Stream(device=self.device) for _ in range(max(0, len(model.inputs) - 1))]
        self._side_streams_hp = [torch.cuda.Stream(device=self.device, priority=-1) for _ in range(max(0, len(model.inputs) - 1))]
        self._copy_stream = torch.cuda.Stream(device=self.device)
        # spare pipeline slots for host batches (stream()): e2e 4.73 -> 4.61 ms per step with one, no further gain with two or
        # three (profiles/r02_upload_slots_ab.txt; the differences are close to the run-to-run spread of a power-capped box)
        self.upload_slots = int(os.environ.get("DPFT_UPLOAD_SLOTS", "1"))

    # -- construction ---------------------------------------------------------------------------------------------
    @staticmethod
    def _version(model) -> int:
        return sum(p._version for p in model.parameters()) + sum(b._version for b in model.buffers())

    @staticmethod
    def _mode_of(model) -> tuple:
        """The model switches a captured graph bakes in: changing one of them must not replay a graph captured under another."""
        return (getattr(model, "native_features", True), getattr(model, "parallel_views", True),
                getattr(model, "side_view_priority", False), getattr(model, "side_view_ctas", 0),
                getattr(model, "pair_side_views", True))

    @staticmethod
    def ineligible_reason(model) -> Optional[str]:
        fuser = model.fuser
        if not isinstance(fuser, IMPFusion):
            return "fuser is not IMPFusion"
        if not isinstance(model.querent, DataAgnosticStaticQueries):
            return "querent is not the static data-agnostic grid"
        if fuser.query.device.type != "cuda":
            return "model is not on a CUDA device"
        if fuser.reduction not in dec.REDUCTIONS:
            return f"reduction {fuser.reduction}"
        if not 1 <= fuser.m_views <= 4:
            return f"m_views={fuser.m_views}"
        if len(set(fuser.n_levels)) != 1:
            return "views with different level counts"
        for mp in fuser.mpfusion.values():
            for layer in mp.ml_fusion_layers.values():
                why = dec.layer_eligible(layer, fuser.d_model)
                if why:
                    return why
        for h in fuser.heads:
            why = dec.head_eligible(h)
            if why:
                return why
        native.load_library()          # raises loudly when the library is missing
        return None

    @classmethod
    def try_create(cls, model) -> Optional["FusedEngine"]:
        if cls.ineligible_reason(model) is not None:
            return None
        return cls(model)

    def accepts(self, batch: Dict[str, torch.Tensor]) -> bool:
        if (self._version(self.model) != self._param_version
                or getattr(self.model, "feature_dtype", torch.float16) != self.feature_dtype
                or getattr(self.model, "pyramid_dtype", torch.float16) != self.pyramid_dtype):   # repack
            self.__init__(self.model)
        if self._mode_of(self.model) != self._mode:       # same weights, another path: drop the graphs captured under the old one
            self._mode = self._mode_of(self.model)
            self._graphs.clear()
            self._pipelines.clear()
            self._seen.clear()
        # float32 tensors (the reference's dataset contract) or, for a view on the native feature path, the uint8 frames an image
        # decoder produces (a quarter of the bytes over PCIe and out of HBM; converted on load inside the stem / FPN kernels)
        native_ok = getattr(self.model, "native_features", True)
        for name, nv in zip(self.model.inputs, self.views):
            x = batch[name]
            if not (x.is_cuda or x.device.type == "cpu"):
                return False
            if x.dtype == torch.uint8:
                if nv is None or not native_ok:
                    return False
            elif x.dtype != torch.float32:
                return False
        return True

    # -- stage 1: feature pyramids --------------------------------------------------------------------------------
    def pyramids(self, batch: Dict[str, torch.Tensor]) -> List[FeaturePyramid]:
        model = self.model
        use_native = getattr(model, "native_features", True)
        out: List[Optional[FeaturePyramid]] = []
        torch_views = [n for n, nv in zip(model.inputs, self.views) if nv is None or not use_native]
        feats = model.extract_features(batch, only=torch_views) if torch_views else {}
        native_idx = [i for i, nv in enumerate(self.views) if nv is not None and use_native]
        results: Dict[int, FeaturePyramid] = {}
        # Two side views with the same architecture on equally shaped inputs (the range-azimuth and elevation-azimuth radar
        # projections of the shipped fusion config) share the launches of their Bottleneck convolutions
        pair = self._paired_side_views(batch, native_idx) if getattr(model, "pair_side_views", True) else None
        if pair is not None and not getattr(model, "parallel_views", True):
            ia, ib = pair
            xa, xb = batch[model.inputs[ia]].contiguous(), batch[model.inputs[ib]].contiguous()
            fa, fb = NativeView.backbone_pair(self.views[ia], self.views[ib], xa, xb)
            for i, x, f in ((ia, xa, fa), (ib, xb, fb)):
                flat, shapes = self.views[i].pyramid(x, feats=f)
                results[i] = FeaturePyramid(flat, shapes)
            native_idx = [i for i in native_idx if i not in pair]
        if len(native_idx) > 1 and getattr(model, "parallel_views", True):
            # the views are independent until the decoder: fork one stream per extra view (the small radar backbones
            # fill the tails of the camera's kernels), join before decoding
            main = torch.cuda.current_stream(self.device)
            fork = torch.cuda.Event()
            fork.record(main)
            joins = []
            if pair is not None:
                # both paired views on ONE forked stream: shared backbone launches, then each view's neck
                ia, ib = pair
                from . import conv as _conv
                side = self._side_streams[-1]
                side.wait_event(fork)
                with torch.cuda.stream(side), _conv.cta_budget(getattr(model, "side_view_ctas", 0)):
                    xa, xb = batch[model.inputs[ia]].contiguous(), batch[model.inputs[ib]].contiguous()
                    fa, fb = NativeView.backbone_pair(self.views[ia], self.views[ib], xa, xb)
                    for i, x, f in ((ia, xa, fa), (ib, xb, fb)):
                        flat, shapes = self.views[i].pyramid(x, feats=f)
                        flat.record_stream(main)
                        results[i] = FeaturePyramid(flat, shapes)
                    done = torch.cuda.Event()
                    done.record(side)
                joins.append(done)
            for k, i in enumerate(i for i in native_idx if pair is None or i not in pair):
                name = model.inputs[i]
                if k == 0:
                    flat, shapes = self.views[i].pyramid(batch[name])
                else:
                    # model.side_view_priority: the small (radar) views on HIGH-priority streams get SMs as soon as any
                    # frees up and are out of the way early; on normal streams they interleave with the first view's kernels
                    hp = getattr(model, "side_view_priority", False)
                    side = (self._side_streams_hp if hp else self._side_streams)[k - 1]
                    side.wait_event(fork)
                    from . import conv as _conv
                    with torch.cuda.stream(side), _conv.cta_budget(getattr(model, "side_view_ctas", 0)):
                        flat, shapes = self.views[i].pyramid(batch[name])
                        done = torch.cuda.Event()
                        done.record(side)
                    flat.record_stream(main)
                    joins.append(done)
                results[i] = FeaturePyramid(flat, shapes)
            for done in joins:
                main.wait_event(done)
        else:
            for i in native_idx:
                flat, shapes = self.views[i].pyramid(batch[model.inputs[i]])
                results[i] = FeaturePyramid(flat, shapes)
        for i, name in enumerate(model.inputs):
            out.append(results[i] if i in results else FeaturePyramid.from_levels(feats[name]))
        return out

    def _paired_side_views(self, batch, native_idx):
        """(i, j): two native views other than the first one with the same architecture and equally shaped float32 inputs."""
        side = native_idx[1:]
        for a in range(len(side)):
            for b in range(a + 1, len(side)):
                i, j = side[a], side[b]
                xi, xj = batch[self.model.inputs[i]], batch[self.model.inputs[j]]
                if xi.shape == xj.shape and xi.dtype == xj.dtype == torch.float32 and self.views[i].same_architecture(self.views[j]):
                    return i, j
        return None

    # -- stage 2: decoder -----------------------------------------------------------------------------------------
    def decode(self, batch: Dict[str, torch.Tensor], pyramids: List[FeaturePyramid]) -> "OrderedDict[str, torch.Tensor]":
        model, dev = self.model, self.device
        B = pyramids[0].flat.shape[0]
        N, V = self.N, self.V
        keep = []                                  # keeps the small per-call device tensors alive
        base_views = []
        flats = [pyr.flat if pyr.flat.is_contiguous() else pyr.flat.contiguous() for pyr in pyramids]
        if any(f.dtype not in (torch.float32, torch.float16) for f in flats) or len({f.dtype for f in flats}) != 1:
            flats = [f.float() for f in flats]     # mixed native / torch views: one storage type per launch
        pyr_dtype = flats[0].dtype
        for name, pyr, flat in zip(model.inputs, pyramids, flats):
            t = batch[f"label_to_{name}_t"].float().contiguous()
            p = batch[f"label_to_{name}_p"].float()
            if p.shape[1] == 3:                    # radar projections are 3x4 (dataset.py:271-293)
                p = torch.cat((p, self._row_0001.expand(B, 1, 4)), dim=1)
            p = p.contiguous()
            shape_hw = batch[f"{name}_shape"][:, :2].float().contiguous()
            flag = t.any().to(torch.int32).reshape(1)
            keep += [t, p, shape_hw, flag, flat]
            view = dec.DecoderView()
            view.pyramid, view.transform, view.projection = flat.data_ptr(), t.data_ptr(), p.data_ptr()
            view.shape_hw, view.use_transform, view.S = shape_hw.data_ptr(), flag.data_ptr(), flat.shape[1]
            start = 0
            for l, (h, w) in enumerate(pyr.shapes):
                view.level_h[l], view.level_w[l], view.level_start[l] = h, w, start
                start += h * w
            base_views.append(view)

        center = model.querent.grid(torch.float32, dev).contiguous()      # (N, 3): same for every sample
        query = self.query                                                # (N, 16)
        views_out = torch.empty((B, V, N, 16), dtype=torch.float32, device=dev)
        out = None
        with torch.cuda.device(dev):
            for it in range(self.I):
                for v, view in enumerate(base_views):
                    view.weights = self.layer_w[it][v].data_ptr()
                dec.layer_forward(base_views, query, self.pos, center, views_out, B, N, self.levels[0], self.P,
                                  self.d_ffn, self.act, self.layer_w[it][0].numel(), pyr_dtype)
                last = it == self.I - 1
                query_out = torch.empty((B, N, 16), dtype=torch.float32, device=dev)
                center_out = torch.empty((B, N, 3), dtype=torch.float32, device=dev)
                size = torch.empty((B, N, 3), dtype=torch.float32, device=dev) if last else None
                angle = torch.empty((B, N, 2), dtype=torch.float32, device=dev) if last else None
                klass = torch.empty((B, N, self.n_cls), dtype=torch.float32, device=dev) if last else None
                dec.head_forward(views_out, self.head_w[it], center, query_out, center_out, size, angle, klass,
                                 B, V, N, self.n_cls, self.reduction)
                query, center = query_out, center_out
                if last:
                    out = OrderedDict(center=center_out, size=size, angle=angle)
                    out["class"] = klass
        return out

    def forward_eager(self, batch: Dict[str, torch.Tensor]) -> "OrderedDict[str, torch.Tensor]":
        return self.decode(batch, self.pyramids(batch))

    # -- CUDA graph: the forward is ~250 short launches with no host decisions in between ---------------------------
    def _keys(self) -> List[str]:
        keys = []
        for name in self.model.inputs:
            keys += [name, f"{name}_shape", f"label_to_{name}_t", f"label_to_{name}_p"]
        return keys

    def forward(self, batch: Dict[str, torch.Tensor]) -> "OrderedDict[str, torch.Tensor]":
        """``batch`` may live on the device or on the host (pinned host tensors are uploaded on a copy stream into
        one of two captured input slots, so the upload of step k+1 overlaps the compute of step k)."""
        on_host = not batch[self.model.inputs[0]].is_cuda
        if not getattr(self.model, "use_cuda_graph", True) or torch.cuda.is_current_stream_capturing():
            return self.forward_eager(self._to_device(batch) if on_host else batch)
        sig = tuple((k, tuple(batch[k].shape), batch[k].dtype) for k in self._keys())
        caps = self._graphs.get(sig)
        if caps is None:
            self._seen[sig] = self._seen.get(sig, 0) + 1
            if self._seen[sig] < 2:                      # first sighting of these shapes: run eagerly (fills the caches)
                return self.forward_eager(self._to_device(batch) if on_host else batch)
            dev_batch = self._to_device(batch) if on_host else batch
            first = _CapturedForward(self, dev_batch, None)
            caps = self._graphs[sig] = [first, _CapturedForward(self, dev_batch, first.graph.pool())]
        self._slot ^= 1
        return caps[self._slot].replay(batch, self._copy_stream if on_host else None)

    def _to_device(self, batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        return {k: batch[k].to(self.device, non_blocking=True) for k in self._keys()}

    # -- pipelined replay: consecutive forwards overlap on the GPU -------------------------------------------------------
    def _pipeline(self, batch: Dict[str, torch.Tensor], depth: int) -> List["_PipelineSlot"]:
        on_host = not batch[self.model.inputs[0]].is_cuda
        sig = ("stream", depth) + tuple((k, tuple(batch[k].shape), batch[k].dtype) for k in self._keys())
        slots = self._pipelines.get(sig)
        if slots is None:
            with torch.no_grad():
                dev_batch = self._to_device(batch) if on_host else batch
                self.forward_eager(dev_batch)                # fills the per-shape caches outside any capture
                slots = self._pipelines[sig] = [_PipelineSlot(self, dev_batch) for _ in range(depth)]
        return slots

    def stream_slots(self, example: Dict[str, torch.Tensor], depth: int = 2) -> List[Dict[str, torch.Tensor]]:
        """The captured INPUT buffers of the ``depth`` pipeline slots for batches shaped like ``example``.  A producer that
        already works on the device (a decode / pre-processing kernel, a DMA engine) can write batch i straight into
        ``slots[i % depth]`` and hand that very dictionary to ``stream``: the replay then reads it in place and the
        device-to-device staging copy of the input (113.6 MB per step on the bench workload) is skipped.  The buffer of a slot
        may be refilled once the output of the forward that read it has been handed out."""
        return [s.captured.static_in for s in self._pipeline(example, max(1, int(depth)))]

    def stream(self, batches, depth: int = 2):
        """Generator over ``batches`` (same shapes) yielding their outputs in order, with up to ``depth`` forwards in flight:
        forward k+1 is replayed on its own stream, from its own captured graph and memory pool, before the outputs of
        forward k are handed out.  A single forward leaves SMs idle in the second wave of every 2-wave convolution (225
        m-tiles on 148 SMs in stage 3), in the latency-bound decoder and while the stem ramps up; the neighbouring
        forward's kernels fill those holes.  Batches may be device tensors or pinned host tensors (uploaded on the copy
        stream).  Outputs are private copies, valid on the caller's current stream."""
        from collections import deque
        depth = max(1, int(depth))
        pending = deque()
        done_events = deque(maxlen=depth)          # completion events of the last `depth` forwards launched
        slots = None
        index = 0
        for batch in batches:
            on_host = not batch[self.model.inputs[0]].is_cuda
            if slots is None:
                # Host batches get `upload_slots` slots more than forwards in flight: a slot's captured input buffer is read until the very
                # end of its forward (the FPN raw level), so with `depth` slots the upload of batch k+1 could only start when
                # forward k+1-depth had finished and its 2 ms sat exposed in front of every replay.  With spare slots the
                # upload lands in a buffer nobody reads while `depth` forwards run; the replay itself is still gated on the
                # completion of forward k+1-depth, so never more than `depth` forwards share the GPU.
                n_slots = depth + (self.upload_slots if on_host else 0)
                slots = self._pipeline(batch, n_slots)
                shapes = {k: tuple(batch[k].shape) for k in self._keys()}
            elif {k: tuple(batch[k].shape) for k in self._keys()} != shapes:
                raise RuntimeError("FusedEngine.stream: every batch of one stream must have the same shapes")
            gate = done_events[0] if (len(slots) > depth and len(done_events) == depth) else None
            with torch.no_grad():
                pending.append(slots[index % len(slots)].launch(batch, self._copy_stream if on_host else None, gate))
            done_events.append(pending[-1][1])
            index += 1
            if len(pending) >= depth:
                yield _PipelineSlot.collect(pending.popleft(), self.device)
        while pending:
            yield _PipelineSlot.collect(pending.popleft(), self.device)


class _CapturedForward:
    """One captured forward for one set of input shapes: static input buffers -> graph -> static outputs."""

    def __init__(self, engine: FusedEngine, batch: Dict[str, torch.Tensor], pool):
        self.device = engine.device
        self.keys = engine._keys()
        self.static_in = {k: torch.empty_like(batch[k], device=engine.device) for k in self.keys}
        for k in self.keys:
            self.static_in[k].copy_(batch[k])
        side = torch.cuda.Stream(device=engine.device)
        side.wait_stream(torch.cuda.current_stream(engine.device))
        with torch.cuda.stream(side):
            engine.forward_eager(self.static_in)
        torch.cuda.current_stream(engine.device).wait_stream(side)
        before = native.launches()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, pool=pool):
            self.static_out = engine.forward_eager(self.static_in)
        self.n_launches = native.launches() - before
        self.consumed = torch.cuda.Event()           # the last replay has finished reading static_in
        self.consumed.record(torch.cuda.current_stream(self.device))

    def replay(self, batch: Dict[str, torch.Tensor], copy_stream) -> "OrderedDict[str, torch.Tensor]":
        main = torch.cuda.current_stream(self.device)
        if copy_stream is None:
            for k in self.keys:
                src = batch[k]
                if src.data_ptr() == self.static_in[k].data_ptr():
                    continue                     # the caller filled this slot's captured input buffer in place (stream_slots)
                self.static_in[k].copy_(src, non_blocking=True)
                # `main` may be a pipeline slot's private stream while `src` was allocated on the caller's (or the feeder's copy)
                # stream: tell the caching allocator that this stream still reads it, or the block can be handed out again while
                # the copy is queued behind the slot's previous replay
                if src.is_cuda:
                    src.record_stream(main)
        else:
            copy_stream.wait_event(self.consumed)
            with torch.cuda.stream(copy_stream):
                for k in self.keys:
                    self.static_in[k].copy_(batch[k], non_blocking=True)
                uploaded = torch.cuda.Event()
                uploaded.record(copy_stream)
            main.wait_event(uploaded)
        self.graph.replay()
        self.consumed.record(main)
        native.count_launch(self.n_launches)
        return OrderedDict((k, v.clone()) for k, v in self.static_out.items())


class _PipelineSlot:
    """A captured forward with a PRIVATE memory pool and its own stream, so that several can replay concurrently."""

    def __init__(self, engine: FusedEngine, batch: Dict[str, torch.Tensor]):
        self.captured = _CapturedForward(engine, batch, None)
        self.stream = torch.cuda.Stream(device=engine.device)
        self.device = engine.device

    def launch(self, batch: Dict[str, torch.Tensor], copy_stream, gate=None):
        caller = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(caller)          # inputs made on the caller's stream; also orders after this slot's last hand-out
        if gate is not None:
            self.stream.wait_event(gate)         # the replay (not the upload) waits until an older forward has left the GPU
        with torch.cuda.stream(self.stream):
            outs = self.captured.replay(batch, copy_stream)
            done = torch.cuda.Event()
            done.record(self.stream)
        return outs, done

    @staticmethod
    def collect(pending, device) -> "OrderedDict[str, torch.Tensor]":
        outs, done = pending
        caller = torch.cuda.current_stream(device)
        caller.wait_event(done)
        for v in outs.values():
            v.record_stream(caller)
        return outs
