#!/usr/bin/env python
"""Latency of ONE mj_compose / mj_effect_* call through the drop-in API (host libjpeg arrays in, host arrays out;
K1 + staging + K2 + staging inside the call) next to the unmodified reference on the same host, single thread.
   c1  tests/golden/image.jpg (256x256 4:2:0) + dropon.png (160x50 RGBA), top left           (BASELINE configs[0])
   c2  6000x4000 4:2:0 + 1024^2 radial-alpha watermark, centred                               (configs[1])
usage (GPU box): python profiles/api_latency.py > gpurun_out/api_latency.json"""
import json
import os
import statistics
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import libmodjpeg_b200 as M  # noqa: E402
import util  # noqa: E402
from oracle import oracle_py as O  # noqa: E402  (timed as the baseline only)


def med(f, n):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        f()
        ts.append(time.perf_counter() - t0)
    return statistics.median(ts) * 1e3


def case(name, data, raw, align, reps):
    ref = O.Reference()
    out = {"config": name}
    d = M.Dropon()
    assert d.read_dropon_from_raw(raw, M.CS_RGBA, 255) == 0
    j = M.Jpeg()
    assert j.read_jpeg_from_memory(data) == 0
    j.compose(d, align, 0, 0)  # warm: context, pools
    jr = ref.read_jpeg(data)
    dr = ref.dropon_from_raw(raw, O.CS_RGBA, 255)

    def ours():
        assert j.compose(d, align, 0, 0) == 0

    def theirs():
        assert jr.compose(dr, align, 0, 0) == 0

    out["mj_compose_ms"] = {"b200": med(ours, reps), "reference_1_thread": med(theirs, max(3, reps // 4))}
    for eff, fo, fr in [("luminance", lambda: j.effect_luminance(3), lambda: jr.luminance(3)),
                        ("tint", lambda: j.effect_tint(2, -2), lambda: jr.tint(2, -2)),
                        ("pixelate", lambda: j.effect_pixelate(), lambda: jr.pixelate())]:
        out[f"mj_effect_{eff}_ms"] = {"b200": med(fo, reps), "reference_1_thread": med(fr, max(3, reps // 4))}
    # the writer: libjpeg's entropy encoder on the host (what mj_write_jpeg_to_memory and the reference do) against K4 behind
    # the same markers (mjx_write_jpeg_to_memory_device: planes staged to the device, six launches, the segment back)
    j.write_jpeg_to_memory_device(0)
    out["mj_write_jpeg_to_memory_ms"] = {"host_libjpeg": med(lambda: j.write_jpeg_to_memory(0), reps), "device_k4": med(lambda: j.write_jpeg_to_memory_device(0), reps)}
    return out


def main():
    from PIL import Image

    g = os.path.join(ROOT, "tests", "golden")
    res = [case("c1", open(os.path.join(g, "image.jpg"), "rb").read(), np.array(Image.open(os.path.join(g, "dropon.png")).convert("RGBA")), 4 | 1, 40)]
    yy, xx = np.mgrid[0:1024, 0:1024]
    r = np.hypot(yy - 511.5, xx - 511.5)
    wm = np.zeros((1024, 1024, 4), np.uint8)
    wm[:, :, 0], wm[:, :, 1], wm[:, :, 2] = xx // 4, yy // 4, 128
    wm[:, :, 3] = np.clip((480 - r) / 96 * 255, 0, 255).astype(np.uint8)
    res.append(case("c2", util.jpeg_bytes(6000, 4000, "420", 85, seed=2), wm, 16, 20))
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
