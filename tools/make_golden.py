"""Generates tests/golden/*.pt by running the UNMODIFIED reference (imported from /root/reference, with the CPU
oracle standing in for its absent compiled op — tools/reference_shim.py) on seeded synthetic inputs.

Run in the build container only:  python tools/make_golden.py
Fixtures hold the reference OUTPUTS plus the recipe (config name, seeds, sizes); weights and inputs are rebuilt
from the seeds by dpft_b200.synthetic, so the files stay small.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, ".."))

import reference_shim  # noqa: E402
from dpft_b200 import configs, synthetic  # noqa: E402

GOLDEN = os.path.join(HERE, "..", "tests", "golden")

CASES = [
    # name, config, batch, sizes, n_queries, weight seed, input seed
    dict(name="radar_bev_native", config="kradar_radar_bev", batch=1, sizes=None, n_queries=None),
    dict(name="radar_bev_256", config="kradar_radar_bev", batch=1, sizes={"radar_bev": (256, 256, 6)}, n_queries=None),
    dict(name="radar_front_native", config="kradar_radar_front", batch=2, sizes=None, n_queries=None),
    dict(name="camera_mono_small", config="kradar_camera_mono", batch=2, sizes={"camera_mono": (90, 160, 3)},
         n_queries=None),
    dict(name="fusion_small_300q", config="kradar", batch=2,
         sizes={"camera_mono": (96, 160, 3), "radar_bev": (64, 48, 6), "radar_front": (37, 48, 6)},
         n_queries=(20, 15, 1)),
    dict(name="fusion_native_1", config="kradar", batch=1,
         sizes={"camera_mono": (128, 228, 3)}, n_queries=None),
    # BASELINE config 5's structure: 900 queries, 4 levels (raw + 3 stages), d_model 64 (8 heads of 8 channels)
    dict(name="stress_4level_d64_900q", config="kradar", batch=1,
         sizes={"camera_mono": (96, 160, 3), "radar_bev": (64, 48, 6), "radar_front": (37, 48, 6)},
         n_queries=(30, 30, 1), multi_scale=3, d_model=64),
    # the two-view radar configuration (config/kradar_radar.json): V = 2 in the view reduction and the fused decoder
    dict(name="radar_two_views", config="kradar_radar", batch=2,
         sizes={"radar_bev": (64, 107, 6), "radar_front": (37, 107, 6)}, n_queries=None),
]


def main(only=None):
    """``python tools/make_golden.py [case ...]``: all fixtures, or only the named model cases."""
    ref = reference_shim.import_reference_models()
    os.makedirs(GOLDEN, exist_ok=True)
    for i, case in enumerate(CASES):
        if only and case["name"] not in only:
            continue
        cfg = synthetic.offline_config(configs.make_config(case["config"]), n_queries=case["n_queries"],
                                       multi_scale=case.get("multi_scale"), d_model=case.get("d_model"))
        model = ref.build("dprt", cfg).eval()
        wseed, iseed = 100 + i, 200 + i
        sd = synthetic.seeded_state_dict(model.state_dict(), seed=wseed)
        model.load_state_dict(sd, strict=True)
        batch = synthetic.synthetic_batch(cfg, case["batch"], seed=iseed, sizes=case["sizes"])
        with torch.no_grad():
            out = model({k: v.clone() for k, v in batch.items()})
        rec = dict(case=case, weight_seed=wseed, input_seed=iseed, torch_version=torch.__version__,
                   outputs={k: v.clone() for k, v in out.items()},
                   state_dict_keys=list(sd.keys()) if case["config"] == "kradar" and i == 4 else None,
                   state_dict_shapes={k: tuple(v.shape) for k, v in sd.items()} if i == 4 else None)
        torch.save(rec, os.path.join(GOLDEN, case["name"] + ".pt"))
        print(case["name"], {k: (tuple(v.shape), float(v.abs().max())) for k, v in out.items()})

    if only:
        return
    # module-level fixture: the reference's MSDeformAttn (its Python arithmetic around the op)
    from dprt.models.layers import MSDeformAttn
    torch.manual_seed(7)
    mod = MSDeformAttn(d_model=16, n_levels=3, n_heads=8, n_points=4).eval()
    msd = synthetic.seeded_state_dict(mod.state_dict(), seed=77)
    mod.load_state_dict(msd)
    g = torch.Generator().manual_seed(78)
    shapes = [(12, 9), (6, 5), (2, 4)]
    S = sum(h * w for h, w in shapes)
    query = torch.randn(2, 11, 16, generator=g)
    refp = torch.rand(2, 11, 3, 2, generator=g)
    flat = torch.randn(2, S, 16, generator=g)
    sh = torch.tensor(shapes)
    lsi = torch.tensor([0, 108, 138])
    with torch.no_grad():
        out = mod(query, refp, flat, sh, lsi)
    torch.save(dict(state_dict=msd, query=query, ref=refp, flat=flat, shapes=shapes, out=out),
               os.path.join(GOLDEN, "msdeformattn_module.pt"))
    print("msdeformattn_module", tuple(out.shape))


if __name__ == "__main__":
    main(sys.argv[1:] or None)
