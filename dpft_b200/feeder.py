"""Batch feeder: the steps of the reference data path that sit between the decoded sample files and ``DPRT.forward``,
restated batch-wise on the device.

The reference runs them per sample on the CPU inside ``KRadarDataset.__getitem__`` (src/dprt/datasets/kradar/dataset.py:139-178)
and then collates (``listed_collating``, src/dprt/datasets/loader.py:10-33):

  scale_radar_data      (dataset.py:295-317)   power cube -> (v - min_power) / (max_power - min_power) * 255, clipped to 0..255
  _add_transformations  (dataset.py:192-212)   camera: zeros_like(calibration);  radar: the calibration matrix itself
  _add_projections      (dataset.py:214-233)   camera: the calibration matrix;   radar: the fixed RA / EA raster projections
                                               (dataset.py:257-293, raster lengths from kradar/utils/radar_info.py)
  _add_shape            (dataset.py:235-255)   the input shape BEFORE the image is resized
  resize_image          (dataset.py:319-341)   torchvision resize of the camera image (smaller edge -> image_size)

Here the decoded data is uploaded as it comes out of the decoders — camera frames as **uint8** (what ``read_image`` returns
before the reference's ``.type(float32)``; a quarter of the bytes), radar cubes as float32 or float16 — on a copy stream
from pinned host memory, and the arithmetic above runs on the GPU on whole batches.  torch / torchvision are used for the
plumbing (copies, the library resize); no custom kernel is involved, so the module also runs on CPU tensors, which is how
the parity tests check it against the reference's own dataset methods (tests/test_feeder.py).
"""
from __future__ import annotations

from typing import Dict, Iterable, Iterator, Optional, Sequence, Tuple, Union

import torch

# kradar/utils/radar_info.py: power range of the cubes, raster lengths / maximum range of the projections
MIN_POWER, MAX_POWER = 100.0, 200.0
AZIMUTH_BINS, ELEVATION_BINS, RANGE_BINS, RANGE_MAX = 107, 37, 256, 118.03710938

RADAR_VIEWS = ("radar_bev", "radar_front")


def radar_projection(view: str, dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """(3, 4) raster projection of a radar view: [u, v, 1] = P [r, phi, rho, 1] (dataset.py:257-293)."""
    if view == "radar_bev":
        return torch.tensor([[0.0, -1.0, 0.0, (AZIMUTH_BINS - 1) / 2], [RANGE_BINS / RANGE_MAX, 0.0, 0.0, 0.0],
                             [0.0, 0.0, 0.0, 1.0]], dtype=dtype)
    if view == "radar_front":
        return torch.tensor([[0.0, -1.0, 0.0, (AZIMUTH_BINS - 1) / 2], [0.0, 0.0, 1.0, (ELEVATION_BINS - 1) / 2],
                             [0.0, 0.0, 0.0, 1.0]], dtype=dtype)
    raise ValueError(f"no raster projection for view {view!r}")


class BatchFeeder:
    """``prepare(raw)`` turns a batch of decoded samples into the dictionary ``DPRT.forward`` takes; ``stream(raws)`` does
    it for a sequence of batches with the upload of batch k+1 overlapping whatever consumes batch k.

    raw: ``{view: (B, H, W, C) uint8 | float16 | float32, f"label_to_{view}": (B, 4, 4) float}`` for every view in ``inputs``.
    """

    def __init__(self, inputs: Sequence[str], image_size: Optional[Union[int, Tuple[int, int]]] = None, scale: bool = True,
                 device: Union[str, torch.device] = "cuda", dtype: torch.dtype = torch.float32):
        self.inputs = list(inputs)
        self.image_size = image_size
        self.scale = scale
        self.device = torch.device(device)
        self.dtype = dtype
        self._copy_stream = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None
        self._proj = {v: radar_projection(v, dtype).to(self.device) for v in self.inputs if v in RADAR_VIEWS}

    @classmethod
    def from_config(cls, config: dict, device: Union[str, torch.device] = "cuda") -> "BatchFeeder":
        """Reads the keys the reference dataset reads (``config['data']``: image_size, scale; ``config['model']['inputs']``)."""
        data = config.get("data", {})
        return cls(config["model"]["inputs"], image_size=data.get("image_size"), scale=data.get("scale", True), device=device,
                   dtype=getattr(torch, config.get("computing", {}).get("dtype", "float32")))

    # -- upload ---------------------------------------------------------------------------------------------------------
    def upload(self, raw: Dict[str, torch.Tensor]):
        """Host -> device copies of one raw batch on the copy stream.  Returns (device tensors, event)."""
        if self._copy_stream is None:
            return dict(raw), None
        main = torch.cuda.current_stream(self.device)
        self._copy_stream.wait_stream(main)                  # staging buffers of the caching allocator: order after their last use
        with torch.cuda.stream(self._copy_stream):
            dev = {k: v.to(self.device, non_blocking=True) for k, v in raw.items()}
            done = torch.cuda.Event()
            done.record(self._copy_stream)
        return dev, done

    # -- the dataset arithmetic, batch-wise -----------------------------------------------------------------------------
    def transform(self, dev: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        batch: Dict[str, torch.Tensor] = {}
        for view in self.inputs:
            x = dev[view]
            B, H, W, C = x.shape
            calib = dev[f"label_to_{view}"].to(self.dtype)
            x = x.to(self.dtype)                                                        # read_image(...).type(float32)
            if view in RADAR_VIEWS:
                if self.scale:                                                          # dataset.py:308-315
                    x = torch.clip((x - MIN_POWER) / (MAX_POWER - MIN_POWER) * (255 - 0) + 0, 0, 255)
                batch[f"label_to_{view}_t"] = calib                                     # dataset.py:207-210
                batch[f"label_to_{view}_p"] = self._proj[view].unsqueeze(0).expand(B, -1, -1).contiguous()   # :229-232
            else:
                batch[f"label_to_{view}_t"] = torch.zeros_like(calib)                   # dataset.py:203-206
                batch[f"label_to_{view}_p"] = calib                                     # dataset.py:225-228
            batch[f"{view}_shape"] = torch.tensor([[H, W, C]] * B, dtype=torch.int64, device=x.device)   # before the resize
            if view not in RADAR_VIEWS and self.image_size is not None:                 # dataset.py:334-339
                from torchvision.transforms.functional import resize
                x = resize(x.movedim(-1, 1), self.image_size).movedim(1, -1).contiguous()
            batch[view] = x
        return batch

    def prepare(self, raw: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        dev, done = self.upload(raw)
        if done is not None:
            torch.cuda.current_stream(self.device).wait_event(done)
            for t in dev.values():
                t.record_stream(torch.cuda.current_stream(self.device))
        return self.transform(dev)

    def stream(self, raws: Iterable[Dict[str, torch.Tensor]]) -> Iterator[Dict[str, torch.Tensor]]:
        """Model batches for a sequence of raw batches; the upload of the next raw batch is in flight while the caller
        works on the current one (feed it to ``DPRT.infer_stream``)."""
        it = iter(raws)
        try:
            nxt = self.upload(next(it))
        except StopIteration:
            return
        while nxt is not None:
            dev, done = nxt
            try:
                nxt = self.upload(next(it))
            except StopIteration:
                nxt = None
            if done is not None:
                main = torch.cuda.current_stream(self.device)
                main.wait_event(done)
                for t in dev.values():
                    t.record_stream(main)
            yield self.transform(dev)


def synthetic_raw_batch(inputs: Sequence[str], batch_size: int, seed: int = 0,
                        sizes: Optional[Dict[str, Tuple[int, int, int]]] = None, pin: bool = False) -> Dict[str, torch.Tensor]:
    """Decoder-side synthetic data: uint8 camera frames, radar power cubes around the 100..200 dB range (some values
    outside, so the clip is exercised), calibration matrices as dpft_b200.synthetic makes them."""
    from . import synthetic
    sizes = {**synthetic.DEFAULT_SIZES, **(sizes or {})}
    g = torch.Generator().manual_seed(seed)
    raw: Dict[str, torch.Tensor] = {}
    for view in inputs:
        H, W, C = sizes[view]
        if view in RADAR_VIEWS:
            raw[view] = torch.rand(batch_size, H, W, C, generator=g) * 120.0 + 90.0
            t = torch.eye(4).repeat(batch_size, 1, 1)
            t[:, :3, 3] = (torch.rand(batch_size, 3, generator=g) - 0.5) * 0.2
            raw[f"label_to_{view}"] = t
        else:
            raw[view] = torch.randint(0, 256, (batch_size, H, W, C), generator=g, dtype=torch.uint8)
            raw[f"label_to_{view}"] = synthetic._camera_projection(H, W).repeat(batch_size, 1, 1)
    if pin:
        raw = {k: v.pin_memory() for k, v in raw.items()}
    return raw
