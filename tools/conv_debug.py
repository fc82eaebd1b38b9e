"""Diagnostic for the tcgen05 conv kernel: each case runs in its own subprocess with a timeout (a barrier bug must not
hang the whole session) and prints one JSON line with error statistics against torch fp32."""
import json
import subprocess
import sys

CASES = [
    # name, B, H, W, Cin, Cout, R, stride, pad, block_n
    ("pw_64_64", 2, 16, 24, 64, 64, 1, 1, 0, 0),
    ("pw_tail", 1, 9, 7, 64, 256, 1, 1, 0, 0),
    ("pw_k256", 2, 23, 40, 256, 64, 1, 1, 0, 0),
    ("pw_k1024_n256", 1, 12, 20, 1024, 256, 1, 1, 0, 256),
    ("pw_n128", 2, 16, 24, 64, 128, 1, 1, 0, 128),
    ("c3_64", 2, 16, 24, 64, 64, 3, 1, 1, 0),
    ("c3_small", 1, 9, 7, 128, 128, 3, 1, 1, 0),
    ("c3_s2", 3, 10, 27, 128, 128, 3, 2, 1, 0),
    ("pw_s2", 2, 23, 40, 256, 512, 1, 2, 0, 0),
    ("c3_tiny", 1, 2, 4, 512, 512, 3, 1, 1, 0),
    ("big_pw", 8, 45, 80, 1024, 256, 1, 1, 0, 0),
    ("big_c3", 8, 45, 80, 256, 256, 3, 1, 1, 0),
    ("pair_pw_small", 2, 16, 24, 256, 256, 1, 1, 0, 256, 2),
    ("pair_pw_tail", 1, 9, 21, 512, 128, 1, 1, 0, 128, 2),
    ("pair_c3", 2, 16, 24, 128, 256, 3, 1, 1, 256, 2),
    ("pair_big_pw", 8, 45, 80, 1024, 256, 1, 1, 0, 256, 2),
    ("pair_big_c3", 8, 45, 80, 256, 256, 3, 1, 1, 256, 2),
]

CHILD = r"""
import json, sys, torch, torch.nn.functional as F
sys.path.insert(0, '.')
from dpft_b200 import conv
cfg = json.loads(sys.argv[1]); name, B, H, W, Cin, Cout, R, stride, pad, bn = cfg[:10]; cm = cfg[10] if len(cfg) > 10 else 1
dev = 'cuda:0'
torch.backends.cudnn.allow_tf32 = False
g = torch.Generator(device=dev).manual_seed(1)
x = torch.randn(B, H, W, Cin, generator=g, device=dev).bfloat16()
w = (torch.randn(Cout, R, R, Cin, generator=g, device=dev) / (R * R * Cin) ** 0.5).bfloat16()
bias = torch.randn(Cout, generator=g, device=dev)
want = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), bias, stride=stride, padding=pad).permute(0, 2, 3, 1)
got = conv.conv2d_nhwc(x, w, bias, stride, pad, False, None, block_n=bn, cluster_mode=cm)
torch.cuda.synchronize()
d = (got.float() - want).abs()
res = {"case": name, "max_err": d.max().item(), "scale": want.abs().max().item(), "mean_err": d.mean().item(),
       "frac_bad": (d > 0.05 * want.abs().max()).float().mean().item(), "finite": bool(torch.isfinite(got.float()).all())}
if res["frac_bad"] > 0:
    bad = (d > 0.05 * want.abs().max())
    res["bad_rows_first"] = bad.flatten(0, 2).any(1).nonzero().flatten()[:12].tolist()
    res["bad_cols_first"] = bad.flatten(0, 2).any(0).nonzero().flatten()[:12].tolist()
    res["got0"] = got.flatten(0, 2)[0, :6].float().tolist(); res["want0"] = want.flatten(0, 2)[0, :6].tolist()
# timing
import time
for _ in range(3): conv.conv2d_nhwc(x, w, bias, stride, pad, True, None, block_n=bn, cluster_mode=cm)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): conv.conv2d_nhwc(x, w, bias, stride, pad, True, None, block_n=bn, cluster_mode=cm)
e1.record(); e1.synchronize()
ms = e0.elapsed_time(e1) / 10
P = (H + 2 * pad - R) // stride + 1; Q = (W + 2 * pad - R) // stride + 1
res["us"] = ms * 1e3; res["tflops"] = 2.0 * B * P * Q * Cout * R * R * Cin / (ms * 1e-3) / 1e12
print(json.dumps(res))
"""

if __name__ == "__main__":
    only = sys.argv[1:] or None
    for case in CASES:
        if only and case[0] not in only:
            continue
        try:
            r = subprocess.run([sys.executable, "-c", CHILD, json.dumps(case)], capture_output=True, text=True, timeout=120)
            out = r.stdout.strip().splitlines()
            print(out[-1] if out else json.dumps({"case": case[0], "rc": r.returncode, "stderr": r.stderr[-600:]}), flush=True)
        except subprocess.TimeoutExpired:
            print(json.dumps({"case": case[0], "error": "TIMEOUT (hang)"}), flush=True)
