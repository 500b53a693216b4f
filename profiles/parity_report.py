#!/usr/bin/env python
"""Parity report of the compositing path on the BASELINE.json configurations: the drop-in API (mj_compose on the
GPU) against the unmodified reference (oracle/_ref) on the same inputs -- differing-coefficient count, largest
difference in quantisation steps, and the PSNR between the two DECODED results (both written with libjpeg and
decoded with Pillow).  usage (GPU box): python profiles/parity_report.py > gpurun_out/parity_report.json"""
import io
import json
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import libmodjpeg_b200 as M  # noqa: E402
import util  # noqa: E402
from oracle import oracle_py as O  # noqa: E402  (the checker)


def psnr(a, b):
    d = a.astype(np.float64) - b.astype(np.float64)
    mse = float((d * d).mean())
    return None if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


def case(name, data, raw, cs, blend, align, ox=0, oy=0):
    ref = O.Reference()
    jr = ref.read_jpeg(data)
    dr = ref.dropon_from_raw(raw, cs, blend)
    assert jr.compose(dr, align, ox, oy) == 0
    want = jr.planes()
    out_ref = jr.write(0)
    j = M.Jpeg()
    assert j.read_jpeg_from_memory(data) == 0
    before = j.planes()
    d = M.Dropon()
    assert d.read_dropon_from_raw(raw, cs, blend) == 0
    assert j.compose(d, align, ox, oy) == 0
    got = j.planes()
    out_b200 = j.write_jpeg_to_memory(0)[1]
    n = differ = changed = 0
    mx = 0
    for a, b, c in zip(got, want, before):
        dd = a.astype(np.int32) - b.astype(np.int32)
        n += dd.size
        differ += int((dd != 0).sum())
        mx = max(mx, int(np.abs(dd).max()))
        changed += int((b != c).sum())
    pa, pb = np.array(Image.open(io.BytesIO(out_b200))), np.array(Image.open(io.BytesIO(out_ref)))
    p = psnr(pa, pb)
    return {"config": name, "coefficients": n, "changed_by_compose": changed, "differing_from_reference": differ,
            "differing_rate_of_changed": differ / max(1, changed), "max_abs_diff_in_quant_steps": mx,
            "decoded_psnr_db_vs_reference": "identical" if p is None else round(p, 2),
            "max_decoded_pixel_diff": int(np.abs(pa.astype(int) - pb.astype(int)).max())}


def main():
    g = os.path.join(ROOT, "tests", "golden")
    res = [case("c1: image.jpg 256x256 4:2:0 + dropon.png, top left", open(os.path.join(g, "image.jpg"), "rb").read(),
                np.array(Image.open(os.path.join(g, "dropon.png")).convert("RGBA")), M.CS_RGBA, 255, 4 | 1)]
    yy, xx = np.mgrid[0:1024, 0:1024]
    r = np.hypot(yy - 511.5, xx - 511.5)
    wm = np.zeros((1024, 1024, 4), np.uint8)
    wm[:, :, 0], wm[:, :, 1], wm[:, :, 2] = xx // 4, yy // 4, 128
    wm[:, :, 3] = np.clip((480 - r) / 96 * 255, 0, 255).astype(np.uint8)
    res.append(case("c2: 6000x4000 4:2:0 + 1024^2 radial-alpha watermark, centred", util.jpeg_bytes(6000, 4000, "420", 85, seed=2), wm, M.CS_RGBA, 255, 16))
    res.append(case("c3: 1920x1080 4:2:0 + full-frame tiled alpha logo", util.jpeg_bytes(1920, 1080, "420", 85, seed=100),
                    util.logo_rgba(1920, 1080, tile=256, radius=110), M.CS_RGBA, 255, 4 | 1))
    wavy = util.wavy_alpha_rgba(3840, 2160)
    res.append(case("c4: 3840x2160 4:4:4, full-frame non-uniform alpha (every block float-blended)", util.jpeg_bytes(3840, 2160, "444", 85, seed=4), wavy, M.CS_RGBA, 255, 4 | 1))
    res.append(case("c4: 3840x2160 grayscale, full-frame non-uniform alpha", util.jpeg_bytes(3840, 2160, "444", 85, seed=4, gray=True), wavy, M.CS_RGBA, 255, 4 | 1))
    rng = np.random.default_rng(6)
    rgb = np.ascontiguousarray(np.repeat(np.repeat(rng.integers(0, 256, size=(270, 480, 3), dtype=np.uint8), 8, 0), 8, 1))
    res.append(case("c4: 3840x2160 4:4:4, full-frame uniform alpha 128 (class U)", util.jpeg_bytes(3840, 2160, "444", 85, seed=4), rgb, M.CS_RGB, 128, 4 | 1))
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
