"""CPU emulation of the index maps of ``fpn_output_tc2_kernel`` (dpft_b200/csrc/feature.cu: the column-owning tile builder
of the raw FPN level).  There is no GPU in the build container, so the thread -> (row, column) mapping, the staged coarse
patch (origin hc0 / wc0, FB_CR x FB_CC cells, clamping) and the host-side eligibility bound are restated here loop for
loop and checked against the direct definition  inner = lateral(raw) + nearest_upsample(coarse)  with zero padding
(tests/test_fpn_builder_emulation.py).  It mirrors the CUDA source by hand: change both together.
"""
import numpy as np

TC_TH, TC_TW, TC_THREADS = 8, 128, 256
FB_CR, FB_CC = 5, 36
FB_ROWS = (TC_TH + 2) // 2


def nearest_src_scaled(dst, scale, in_size):
    s = int(np.floor(np.float32(np.float32(dst) * scale)))
    return s if s < in_size - 1 else in_size - 1


def eligible(H, W, Hc, Wc):
    if Hc <= 0 or Wc <= 0 or Hc > H or Wc > W:
        return False
    return ((TC_TH + 1) * Hc) // H + 3 <= FB_CR and ((TC_TW + 1) * Wc) // W + 3 <= FB_CC


def build_tile(raw, coarse, lat_w, lat_b, p0, q0):
    """raw (H, W, CIN), coarse (Hc, Wc, 16), lat_w (16, CIN), lat_b (16,) -> ((TC_TH+2, TC_TW+2, 16) tile, written mask),
    following the kernel thread by thread."""
    H, W, _ = raw.shape
    Hc, Wc, _ = coarse.shape
    scale_h = np.float32(Hc) / np.float32(H)
    scale_w = np.float32(Wc) / np.float32(W)
    hc0 = nearest_src_scaled(p0 - 1 if p0 > 0 else 0, scale_h, Hc)
    wc0 = nearest_src_scaled(q0 - 1 if q0 > 0 else 0, scale_w, Wc)
    s_cb = np.zeros((FB_CR, FB_CC, 16), np.float32)
    for cell in range(FB_CR * FB_CC):
        cr, cc = divmod(cell, FB_CC)
        hc = min(hc0 + cr, Hc - 1)
        wc = min(wc0 + cc, Wc - 1)
        s_cb[cr, cc] = coarse[hc, wc] + lat_b
    tile = np.zeros((TC_TH + 2, TC_TW + 2, 16), np.float32)
    written = np.zeros((TC_TH + 2, TC_TW + 2), np.int32)

    def column(px, rr0, rows):
        ww = q0 - 1 + px
        col_ok = 0 <= ww < W
        ws = ww if col_ok else 0
        wcl = nearest_src_scaled(ws, scale_w, Wc) - wc0
        for u in range(rows):
            rr = rr0 + u
            hh = p0 - 1 + rr
            ok = col_ok and 0 <= hh < H
            v = np.zeros(16, np.float32)
            if ok:
                hcl = nearest_src_scaled(hh, scale_h, Hc) - hc0
                assert 0 <= hcl < FB_CR and 0 <= wcl < FB_CC, (hcl, wcl, "coarse patch too small")
                v = s_cb[hcl, wcl] + lat_w @ raw[hh, ww]
            tile[rr, px] = v
            written[rr, px] += 1

    for tid in range(TC_THREADS):
        column(tid & (TC_TW - 1), (tid >> 7) * FB_ROWS, FB_ROWS)
    for tid in range(2 * (TC_TH + 2)):
        column(TC_TW + (tid & 1), tid >> 1, 1)
    return tile, written


def direct_tile(raw, coarse, lat_w, lat_b, p0, q0):
    """Definition: inner[h, w] = lat_w @ raw[h, w] + lat_b + coarse[nearest(h), nearest(w)] (torch 'nearest'), zero outside."""
    H, W, _ = raw.shape
    Hc, Wc, _ = coarse.shape
    tile = np.zeros((TC_TH + 2, TC_TW + 2, 16), np.float32)
    sh, sw = np.float32(Hc) / np.float32(H), np.float32(Wc) / np.float32(W)
    for rr in range(TC_TH + 2):
        for px in range(TC_TW + 2):
            hh, ww = p0 - 1 + rr, q0 - 1 + px
            if 0 <= hh < H and 0 <= ww < W:
                hc = min(int(np.floor(np.float32(hh) * sh)), Hc - 1)
                wc = min(int(np.floor(np.float32(ww) * sw)), Wc - 1)
                tile[rr, px] = coarse[hc, wc] + lat_b + lat_w @ raw[hh, ww]
    return tile
