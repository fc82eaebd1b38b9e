"""Times the tcgen05 conv kernel on the camera ResNet-101 layer shapes of the bench workload (bs 8, 720x1280) and prints
achieved TFLOP/s and GB/s per layer.  `python tools/conv_bench.py NAME` runs one case a few times (for ncu)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from dpft_b200 import conv  # noqa: E402

# name: (H, W, Cin, Cout, R, stride, pad, residual)
LAYERS = {
    "s1_conv1": (180, 320, 256, 64, 1, 1, 0, False), "s1_conv2": (180, 320, 64, 64, 3, 1, 1, False),
    "s1_conv3": (180, 320, 64, 256, 1, 1, 0, True),
    "s2_conv1": (90, 160, 512, 128, 1, 1, 0, False), "s2_conv2": (90, 160, 128, 128, 3, 1, 1, False),
    "s2_conv3": (90, 160, 128, 512, 1, 1, 0, True),
    "s3_conv1": (45, 80, 1024, 256, 1, 1, 0, False), "s3_conv2": (45, 80, 256, 256, 3, 1, 1, False),
    "s3_conv3": (45, 80, 256, 1024, 1, 1, 0, True),
    "s4_conv1": (23, 40, 2048, 512, 1, 1, 0, False), "s4_conv2": (23, 40, 512, 512, 3, 1, 1, False),
    "s4_conv3": (23, 40, 512, 2048, 1, 1, 0, True),
    "s3_conv3_nores": (45, 80, 256, 1024, 1, 1, 0, False),
    "s3_down": (90, 160, 512, 1024, 1, 2, 0, False), "s3_conv2_s2": (90, 160, 256, 256, 3, 2, 1, False),
    # radar ResNet-50 at 256x256 (bs 8): small-M layers
    "r1_conv2": (64, 64, 64, 64, 3, 1, 1, False), "r2_conv3": (32, 32, 128, 512, 1, 1, 0, True),
    "r3_conv1": (16, 16, 1024, 256, 1, 1, 0, False), "r3_conv2": (16, 16, 256, 256, 3, 1, 1, False),
    "r3_conv3": (16, 16, 256, 1024, 1, 1, 0, True), "r4_conv2": (8, 8, 512, 512, 3, 1, 1, False),
}
B = 8
dev = "cuda:0"
SWEEP = "--sweep" in sys.argv
NO_LIB = "--no-lib" in sys.argv          # skip the cuDNN timings (variant sweeps)
only = [a for a in sys.argv[1:] if not a.startswith("--")] or None
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, (H, W, Cin, Cout, R, stride, pad, res) in LAYERS.items():
    if only and name not in only:
        continue
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(B, H, W, Cin, generator=g, device=dev).half()
    w = (torch.randn(Cout, R, R, Cin, generator=g, device=dev) / (R * R * Cin) ** 0.5).half()
    bias = torch.randn(Cout, generator=g, device=dev)
    P, Q = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - R) // stride + 1
    r = torch.randn(B, P, Q, Cout, generator=g, device=dev).half() if res else None
    out = torch.empty(B, P, Q, Cout, device=dev, dtype=torch.float16)
    if SWEEP:   # every (output-channel tile, CTA-pair) choice, L2 flushed
        res_row = {"layer": name}
        for bn in (64, 128, 256):
            if Cout % bn:
                continue
            for cm in (1, 2):
                if cm == 2 and bn < 128:
                    continue
                ts = []
                for _ in range(2):
                    conv.conv2d_nhwc(x, w, bias, stride, pad, True, r, out, block_n=bn, cluster_mode=cm)
                for _ in range(5):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    conv.conv2d_nhwc(x, w, bias, stride, pad, True, r, out, block_n=bn, cluster_mode=cm)
                    e1.record()
                    e1.synchronize()
                    ts.append(e0.elapsed_time(e1) * 1e3)
                res_row[f"bn{bn}_cg{cm}"] = round(min(ts), 1)
        print(json.dumps(res_row), flush=True)
        continue
    for _ in range(3):
        conv.conv2d_nhwc(x, w, bias, stride, pad, True, r, out)
    ts_cold, ts_warm = [], []
    for cold in (True, False):
        for _ in range(5):
            if cold:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            conv.conv2d_nhwc(x, w, bias, stride, pad, True, r, out)
            e1.record()
            e1.synchronize()
            (ts_cold if cold else ts_warm).append(e0.elapsed_time(e1) * 1e3)
    # the library on the same box: cuDNN through torch, f16 channels_last, bias fused into the conv call;
    # `cudnn_chain` adds what eval-mode torch would run after it for this layer (residual add, ReLU) as separate kernels
    torch.backends.cudnn.benchmark = True
    xc = x.permute(0, 3, 1, 2)                                  # NHWC storage seen as NCHW = channels_last
    wc = w.permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
    bc = bias.half()
    rc = r.permute(0, 3, 1, 2) if res else None

    def lib_conv():
        return torch.nn.functional.conv2d(xc, wc, bc, stride, pad)

    def lib_chain():
        y = torch.nn.functional.conv2d(xc, wc, bc, stride, pad)
        if res:
            y += rc
        return torch.relu_(y)

    lib = {}
    for nm, fn in (() if NO_LIB else (("cudnn_conv", lib_conv), ("cudnn_chain", lib_chain))):
        for _ in range(3):
            fn()
        tl = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            tl.append(e0.elapsed_time(e1) * 1e3)
        lib[nm + "_us_cold"] = round(min(tl), 1)
    M = B * P * Q
    flops = 2.0 * M * Cout * R * R * Cin
    bytes_ = 2.0 * (B * H * W * Cin + Cout * R * R * Cin + M * Cout * (2 if res else 1))
    tc, tw = min(ts_cold), min(ts_warm)
    print(json.dumps({"layer": name, "M": M, "N": Cout, "K": R * R * Cin, "us_cold": round(tc, 1), "us_warm": round(tw, 1),
                      "tflops_cold": round(flops / tc / 1e6, 1), "gbps_cold": round(bytes_ / tc / 1e3, 1),
                      "tflops_warm": round(flops / tw / 1e6, 1), "MB": round(bytes_ / 1e6, 1), **lib,
                      **({} if NO_LIB else {"speedup_vs_cudnn_chain": round(lib["cudnn_chain_us_cold"] / tc, 2)})}), flush=True)
