"""Decoder self-attention core: the tcgen05 flash-attention kernel vs torch scaled_dot_product_attention on the shapes of
BASELINE configs 3/4 (8 heads of 2 channels, 300/400 queries) and 5 (900 queries, d_model 64/256)."""
import json
import math
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from dpft_b200 import attention  # noqa: E402

DEV = "cuda:0"
CASES = [("cfg3_D2", 8, 8, 300, 2), ("shipped_D2", 8, 8, 400, 2), ("cfg5_D8", 16, 8, 900, 8), ("cfg5_D32", 16, 8, 900, 32),
         ("D64_N1024", 16, 8, 1024, 64)]


def timed(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for name, B, H, N, D in CASES:
    C = H * D
    q, k, v = (torch.randn(B, N, C, device=DEV) for _ in range(3))
    row = {"case": name, "B": B, "H": H, "N": N, "D": D, "flops": 4.0 * B * H * N * N * D}

    def sdpa(q=q, k=k, v=v):
        qh, kh, vh = (t.view(B, N, H, D).transpose(1, 2) for t in (q, k, v))
        return F.scaled_dot_product_attention(qh, kh, vh).transpose(1, 2).reshape(B, N, C)

    row["torch_sdpa_f32_us"] = timed(sdpa)
    row["native_precise_f32_us"] = timed(lambda: attention.self_attention(q, k, v, H, precise=True))
    row["native_fast_f32_us"] = timed(lambda: attention.self_attention(q, k, v, H, precise=False))
    qh, kh, vh = q.half(), k.half(), v.half()
    row["native_f16_us"] = timed(lambda: attention.self_attention(qh, kh, vh, H))
    row["torch_sdpa_f16_us"] = timed(lambda: sdpa(qh, kh, vh))
    ref = sdpa(q.double(), k.double(), v.double())
    row["err_precise"] = float((attention.self_attention(q, k, v, H).double() - ref).abs().max() / ref.abs().max())
    row["err_sdpa_f32"] = float((sdpa().double() - ref).abs().max() / ref.abs().max())
    row["native_f16_tflops"] = row["flops"] / row["native_f16_us"] / 1e6
    print(json.dumps(row), flush=True)
