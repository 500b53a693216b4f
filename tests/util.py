"""Seeded synthetic inputs shared by the tests and bench.py (S1-S5)."""
from __future__ import annotations

import io

import numpy as np
from PIL import Image

SUBS = {"444": 0, "422": 1, "420": 2}


def photo(w: int, h: int, seed: int) -> np.ndarray:
    """Photo-like RGB field: low-frequency sinusoids + N(0, 12) noise."""
    r = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    ph = r.uniform(0, 6.28, (3, 4))
    fr = r.uniform(0.004, 0.05, (3, 4))
    base = np.stack([128 + 50 * np.sin(xx * fr[c, 0] + ph[c, 0]) * np.cos(yy * fr[c, 1] + ph[c, 1])
                     + 40 * np.sin((xx + yy) * fr[c, 2] + ph[c, 2]) for c in range(3)], -1)
    base += r.normal(0, 12, (h, w, 3)).astype(np.float32)
    return np.clip(base, 0, 255).astype(np.uint8)


def jpeg_bytes(w: int, h: int, subsampling: str = "420", quality: int = 85, seed: int = 0, gray: bool = False,
               progressive: bool = False) -> bytes:
    img = Image.fromarray(photo(w, h, seed))
    if gray:
        img = img.convert("L")
    b = io.BytesIO()
    kw = dict(quality=quality, progressive=progressive)
    if not gray:
        kw["subsampling"] = SUBS[subsampling]
    img.save(b, "JPEG", **kw)
    return b.getvalue()


def logo_rgba(w: int, h: int, tile: int = 256, radius: int = 110) -> np.ndarray:
    """Tiled alpha logo: alpha = clip((radius - r) * 8, 0, 255) around each tile centre (S3)."""
    yy, xx = np.mgrid[0:h, 0:w]
    rgb = np.stack([xx * 255 // max(1, w - 1), yy * 255 // max(1, h - 1), (xx + yy) * 255 // max(1, w + h - 2)], -1)
    r = np.hypot((xx % tile) - tile / 2, (yy % tile) - tile / 2)
    a = np.clip((radius - r) * 8, 0, 255)
    return np.dstack([rgb, a]).astype(np.uint8)


def watermark_rgba(size: int = 1024, r_opaque: int = 384, r_clear: int = 480) -> np.ndarray:
    """S2: rgb gradient, alpha 255 inside r_opaque, linear ramp to 0 at r_clear."""
    yy, xx = np.mgrid[0:size, 0:size]
    rgb = np.stack([xx * 255 // (size - 1), yy * 255 // (size - 1), 255 - xx * 255 // (size - 1)], -1)
    r = np.hypot(xx - size / 2, yy - size / 2)
    a = np.clip((r_clear - r) * 255.0 / (r_clear - r_opaque), 0, 255)
    return np.dstack([rgb, a]).astype(np.uint8)


def noisy_rgba(w: int, h: int, seed: int) -> np.ndarray:
    r = np.random.default_rng(seed)
    return r.integers(0, 256, (h, w, 4), dtype=np.uint8)


def wavy_alpha_rgba(w: int, h: int) -> np.ndarray:
    """S4(ii): alpha = 96 + 64 sin(x/5) cos(y/3): every block generic."""
    yy, xx = np.mgrid[0:h, 0:w]
    rgb = np.stack([xx * 255 // max(1, w - 1), yy * 255 // max(1, h - 1), 128 + 0 * xx], -1)
    a = 96 + 64 * np.sin(xx / 5.0) * np.cos(yy / 3.0)
    return np.dstack([rgb, np.clip(a, 0, 255)]).astype(np.uint8)


def diff_stats(a: np.ndarray, b: np.ndarray) -> dict:
    d = a.astype(np.int32) - b.astype(np.int32)
    return dict(n=int(a.size), differing=int((d != 0).sum()), max_abs=int(np.abs(d).max()) if d.size else 0)


def ingest_raw(raw: np.ndarray, colorspace: int, blend: int):
    """numpy restatement of mj_read_dropon_from_raw (reference: src/dropon.c:203-323):
    -> (image3, alpha3, stored_colorspace, stored_blend).  Checked against the reference in
    tests/test_oracle_golden.py."""
    blend = min(max(int(blend), 0), 255)
    raw = np.asarray(raw, np.uint8)
    if raw.ndim == 2:
        raw = raw[:, :, None]
    h, w = raw.shape[:2]
    ncolor = 3 if colorspace in (1, 2, 5, 6) else 1
    has_alpha = colorspace in (2, 4, 6)
    stored = {1: 1, 2: 1, 5: 5, 6: 5, 3: 3, 4: 3}[colorspace]
    image3 = np.repeat(raw[:, :, :1], 3, 2) if ncolor == 1 else raw[:, :, :3].copy()
    if has_alpha:
        alpha3 = np.repeat(raw[:, :, ncolor:ncolor + 1], 3, 2)
        blend = -1
    else:
        alpha3 = np.full((h, w, 3), blend, np.uint8)
    return np.ascontiguousarray(image3), np.ascontiguousarray(alpha3), stored, blend


def oracle_compose(port, planes, qtables, width, height, colorspace, samp, image3, alpha3, dropon_cs, blend, align, ox, oy):
    """mj_compose restated with the oracle port (geometry + compile + blend), in place on `planes`.
    Returns (rv, geometry, D, W) -- rv 6 when libjpeg would reject the conversion."""
    from oracle import oracle_py as O

    dh, dw = image3.shape[:2]
    max_h = max(h for h, _ in samp)
    max_v = max(v for _, v in samp)
    g = port.geometry(width, height, max_h * 8, max_v * 8, dw, dh, align, ox, oy)
    if blend == 0 or not g["visible"]:
        return 0, g, None, None
    L = O.make_layout(colorspace, samp)
    rv, D, W = port.compile_dropon(image3, alpha3, dropon_cs, L, g["blockoffset_x"], g["blockoffset_y"],
                                   (g["crop_x"], g["crop_y"], g["crop_w"], g["crop_h"]))
    if rv != 0:
        return rv, g, None, None
    for c in range(len(samp)):
        port.compose_plane(planes[c], g["block_x"] * samp[c][0], g["block_y"] * samp[c][1], D[c], W[c], qtables[c])
    return 0, g, D, W
