"""Deformable-attention HBM roofline sweep (BASELINE.json config 5 and the shipped shape): times the forward and
backward kernels with CUDA events, L2 flushed between launches, and prints one JSON line per case.

  python tools/msda_sweep.py [--quick] [--only NAME]

Algorithmic bytes per SURVEY.md §8d:
  fwd = B*N*M*L*P*(4D+3)*s + B*N*M*D*s
  bwd = B*N*M*L*P*(4D + 2*4D + 3 + 3)*s + B*N*M*D*s   (value read, grad_value RMW, loc/attn read, grads write)
"""
import argparse
import json
import os
import statistics
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from dpft_b200 import msda  # noqa: E402

CAM4 = [(180, 320), (90, 160), (45, 80), (23, 40)]              # 4 levels below the raw 720x1280 image
CAM5 = [(720, 1280)] + CAM4
CASES = {
    # name: (B, N, M, D, shapes, P, dtype)
    "shipped_fp32_D2": (8, 400, 8, 2, CAM5, 4, torch.float32),
    "cfg3_300q_fp32_D2": (8, 300, 8, 2, CAM5, 4, torch.float32),
    "cfg5_bf16_D2": (16, 900, 8, 2, CAM5[:4], 4, torch.bfloat16),
    "cfg5_bf16_D8": (16, 900, 8, 8, CAM5[:4], 4, torch.bfloat16),
    "cfg5_bf16_D32": (16, 900, 8, 32, CAM5[:4], 4, torch.bfloat16),
    "cfg5_fp32_D32": (16, 900, 8, 32, CAM5[:4], 4, torch.float32),
}


def run(name, reps):
    B, N, M, D, shapes, P, dtype = CASES[name]
    dev = "cuda:0"
    L = len(shapes)
    S = sum(h * w for h, w in shapes)
    g = torch.Generator(device=dev).manual_seed(0)
    value = torch.randn(B, S, M, D, generator=g, device=dev, dtype=torch.float32).to(dtype)
    sh = torch.tensor(shapes, dtype=torch.int64, device=dev)
    lsi = torch.tensor([sum(h * w for h, w in shapes[:i]) for i in range(L)], dtype=torch.int64, device=dev)
    loc = torch.rand(B, N, M, L, P, 2, generator=g, device=dev).to(dtype)
    attn = torch.softmax(torch.randn(B, N, M, L * P, generator=g, device=dev), -1).view(B, N, M, L, P).to(dtype)
    go = torch.randn(B, N, M * D, generator=g, device=dev).to(dtype)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lib = msda.native.load_library()
    acc = torch.float32
    gv = torch.zeros(B, S, M, D, device=dev, dtype=acc)
    gl, ga = torch.empty_like(loc), torch.empty_like(attn)
    out = torch.empty(B, N, M * D, device=dev, dtype=dtype)
    code = msda.native.dtype_code(value)
    st = msda.native.stream_ptr(dev)
    P_ = msda.native.ptr

    def fwd():
        return lib.dpft_msda_forward(P_(value), P_(sh), P_(lsi), P_(loc), P_(attn), P_(out), B, S, M, D, N, L, P, code, st)

    def bwd():
        return lib.dpft_msda_backward(P_(value), P_(sh), P_(lsi), P_(loc), P_(attn), P_(go), P_(gv), P_(gl), P_(ga),
                                      B, S, M, D, N, L, P, code, st)

    def timeit(fn):
        for _ in range(3):
            assert fn() == 0
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3)
        return statistics.mean(ts), min(ts)

    s = value.element_size()
    n_s = B * N * M * L * P
    alg_f = n_s * (4 * D + 3) * s + B * N * M * D * s
    alg_b = n_s * (4 * D * s + 2 * 4 * D * 4 + 3 * s + 3 * s) + B * N * M * D * s
    peak = 6549.4
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["hbm_gbs"]
    tf, tfmin = timeit(fwd)
    tb, tbmin = timeit(bwd)
    print(json.dumps({"case": name, "dtype": str(dtype), "B": B, "N": N, "M": M, "D": D, "L": L, "P": P, "S": S,
                      "fwd_us": tf * 1e6, "fwd_us_min": tfmin * 1e6, "fwd_alg_MB": alg_f / 1e6,
                      "fwd_GBps": alg_f / tf / 1e9, "fwd_frac": alg_f / tf / 1e9 / peak,
                      "bwd_us": tb * 1e6, "bwd_us_min": tbmin * 1e6, "bwd_alg_MB": alg_b / 1e6,
                      "bwd_GBps": alg_b / tb / 1e9, "bwd_frac": alg_b / tb / 1e9 / peak, "peak_GBps": peak}), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    for name in CASES:
        if a.only is None or a.only == name:
            run(name, a.reps)
