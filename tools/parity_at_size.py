"""Parity of the GPU paths at BASELINE.json's full sizes against the reference's CPU eval forward, next to the yardstick that
matters for a 16-bit claim: the reference's OWN forward on the same GPU in PyTorch's default precision (fp32 parameters,
TF32 convolutions: 10-bit mantissa operands like float16).  One JSON line per (config, path):

    python tools/parity_at_size.py [config2|config3] [--batch 8]
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import bench  # noqa: E402
from dpft_b200 import configs, models, synthetic  # noqa: E402

CASES = {"config2": ("kradar_camera_mono", None, {"camera_mono": synthetic.BASELINE_SIZES["camera_mono"]}),
         "config3": ("kradar", (20, 15, 1), dict(synthetic.BASELINE_SIZES)),
         "config1": ("kradar_radar_bev", None, {"radar_bev": synthetic.BASELINE_SIZES["radar_bev"]})}
DEV = "cuda:0"


def main():
    names = [a for a in sys.argv[1:] if a in CASES] or ["config2", "config3"]
    B = int(sys.argv[sys.argv.index("--batch") + 1]) if "--batch" in sys.argv else 8
    for name in names:
        cfg_name, nq, sizes = CASES[name]
        cfg = synthetic.offline_config(configs.make_config(cfg_name), n_queries=nq)
        sd = synthetic.seeded_state_dict(models.build("dprt", cfg).state_dict(), seed=1)
        batch = synthetic.synthetic_batch(cfg, B, seed=1000, sizes=sizes)
        torch.set_num_threads(os.cpu_count())
        fwd, kind, note = bench.reference_forward_fn(cfg, sd)
        with torch.no_grad():
            want = fwd({k: v.clone() for k, v in batch.items()})
        gb = {k: v.to(DEV) for k, v in batch.items()}

        def ours(native, fdt=torch.float16, pdt=torch.float16, fused=True):
            m = models.build("dprt", cfg).eval()
            m.load_state_dict(sd)
            m = m.to(DEV)
            m.use_fused, m.native_features, m.feature_dtype, m.pyramid_dtype = fused, native, fdt, pdt
            with torch.no_grad():
                out = m(gb)
            grid = m.querent.grid(torch.float32, DEV).cpu()
            return {k: v.float().cpu() for k, v in out.items()}, grid

        def emit(path, got, grid):
            rep = bench.parity_report(got, want, grid)
            print(json.dumps({"config": name, "batch": B, "path": path, "against": kind,
                              "max_norm_rel": {k: round(v["max_norm_rel"], 6) for k, v in rep.items()},
                              "elem_rel_p99": {k: round(v["elem_rel_p99"], 6) for k, v in rep.items()}}), flush=True)

        for tf32 in (True, False):                     # the reference's own GPU forward: default TF32 convolutions, then strict fp32
            torch.backends.cudnn.allow_tf32 = tf32
            ref_models, _ = bench.reference_package()
            ref = ref_models.build("dprt", cfg).eval()
            ref.load_state_dict(sd)
            ref = ref.to(DEV)
            with torch.no_grad():
                out = {k: v.float().cpu() for k, v in ref(gb).items()}
            grid = models.build("dprt", cfg).querent.grid(torch.float32, "cpu")
            emit("reference_on_gpu_tf32_default" if tf32 else "reference_on_gpu_fp32_strict", out, grid)
            del ref
        torch.backends.cudnn.allow_tf32 = False
        emit("ours_composed_fp32", *ours(False, fused=False))
        emit("ours_fused_decoder_fp32_features", *ours(False))
        emit("ours_native_f16_pyramid_f16", *ours(True))
        emit("ours_native_f16_pyramid_f32", *ours(True, pdt=torch.float32))
        emit("ours_native_bf16_pyramid_f32", *ours(True, fdt=torch.bfloat16, pdt=torch.float32))
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
