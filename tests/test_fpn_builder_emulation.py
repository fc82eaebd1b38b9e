"""Index maps of the column-owning raw-level FPN tile builder (fpn_output_tc2_kernel), emulated on the CPU
(tools/emulate_fpn_builder.py): every halo entry the MMAs read is written exactly once with the value the definition
gives, and the staged coarse patch is large enough wherever the host-side eligibility test lets the kernel run."""
import numpy as np
import pytest

import emulate_fpn_builder as E


@pytest.mark.parametrize("H,W,Hc,Wc,cin", [(37, 107, 10, 27, 6), (45, 300, 12, 75, 3), (64, 256, 16, 64, 6), (17, 129, 5, 33, 6)])
def test_column_builder_tiles_match_the_definition(H, W, Hc, Wc, cin):
    assert E.eligible(H, W, Hc, Wc)
    rng = np.random.default_rng(H * W)
    raw = (rng.random((H, W, cin)) * 255).astype(np.float32)
    coarse = rng.standard_normal((Hc, Wc, 16)).astype(np.float32)
    lat_w = (rng.standard_normal((16, cin)) * 0.01).astype(np.float32)
    lat_b = rng.standard_normal(16).astype(np.float32)
    for p0 in range(0, H, E.TC_TH):
        for q0 in range(0, W, E.TC_TW):
            got, written = E.build_tile(raw, coarse, lat_w, lat_b, p0, q0)
            assert (written == 1).all()
            want = E.direct_tile(raw, coarse, lat_w, lat_b, p0, q0)
            np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("H,W", [(720, 1280), (512, 910), (256, 256), (256, 107), (37, 107), (96, 160), (50, 37 + 96)])
def test_coarse_patch_covers_every_tile_of_the_shipped_level_sizes(H, W):
    """Rows / columns of the coarse patch a tile touches, for the raw level over the stage-1 level ((H-1)//4+1: stem stride 2,
    max-pool stride 2), computed with the kernel's own float arithmetic."""
    Hc, Wc = ((H - 1) // 2 + 1 - 1) // 2 + 1, ((W - 1) // 2 + 1 - 1) // 2 + 1
    assert E.eligible(H, W, Hc, Wc)
    sh, sw = np.float32(Hc) / np.float32(H), np.float32(Wc) / np.float32(W)
    for p0 in range(0, H, E.TC_TH):
        lo = E.nearest_src_scaled(max(p0 - 1, 0), sh, Hc)
        hi = E.nearest_src_scaled(min(p0 + E.TC_TH, H - 1), sh, Hc)
        assert hi - lo < E.FB_CR
    for q0 in range(0, W, E.TC_TW):
        lo = E.nearest_src_scaled(max(q0 - 1, 0), sw, Wc)
        hi = E.nearest_src_scaled(min(q0 + E.TC_TW, W - 1), sw, Wc)
        assert hi - lo < E.FB_CC


def test_ineligible_shapes_are_refused():
    assert not E.eligible(64, 256, 32, 128)        # a 2x coarser map would need a larger patch
    assert not E.eligible(64, 256, 0, 0)
