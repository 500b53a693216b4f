#!/usr/bin/env python
"""mj_compose_batch (K5 -> K1 -> K2 -> K4 on the device where the window allows it) against the per-image calls
(mj_read_jpeg_from_memory -> mj_compose -> mj_write_jpeg_to_memory, host libjpeg on both ends): random sizes, samplings,
qualities, optimised / progressive / damaged inputs, several batches with one geometry (the all-device path) and several mixed
ones.  Every output must be byte-identical.
usage (GPU box): python profiles/fuzz_batch.py [images] [seed] > gpurun_out/fuzz_batch.json"""
import io
import json
import os
import sys

import numpy as np
from PIL import Image, ImageFile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import libmodjpeg_b200 as M  # noqa: E402
import util  # noqa: E402
from libmodjpeg_b200 import capi  # noqa: E402

ImageFile.MAXBLOCK = 1 << 24


def make(rng, w, h, kind):
    img = util.photo(w, h, int(rng.integers(1 << 30)))
    if kind == "noise":
        img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    buf = io.BytesIO()
    subs = int(rng.integers(0, 3))
    q = int(rng.integers(30, 101))
    pil = Image.fromarray(img)
    if kind == "gray":
        pil = pil.convert("L")
        pil.save(buf, "JPEG", quality=q)
    elif kind == "progressive":
        pil.save(buf, "JPEG", quality=q, subsampling=subs, progressive=True)
    elif kind == "optimised":
        pil.save(buf, "JPEG", quality=q, subsampling=subs, optimize=True)
    else:
        pil.save(buf, "JPEG", quality=q, subsampling=subs)
    return buf.getvalue(), subs, q


def main():
    total = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
    report = {"batches": [], "images": 0, "mismatches": 0, "status_mismatches": 0}
    done = 0
    while done < total:
        uniform = bool(rng.integers(0, 2))
        n = int(rng.integers(3, 70))
        w0, h0 = int(rng.integers(16, 700)), int(rng.integers(16, 500))
        files, kinds = [], []
        subs0 = None
        for i in range(n):
            kind = rng.choice(["plain", "plain", "plain", "gray", "progressive", "optimised", "noise"]) if not uniform else rng.choice(["plain", "plain", "noise"])
            w, h = (w0, h0) if uniform else (int(rng.integers(16, 700)), int(rng.integers(16, 500)))
            data, subs, q = make(rng, w, h, kind)
            if uniform:  # one sampling for the whole batch
                while subs0 is not None and subs != subs0:
                    data, subs, q = make(rng, w, h, kind)
                subs0 = subs
            if rng.random() < 0.05:  # damage
                b = bytearray(data)
                off = len(b) // 2
                if rng.random() < 0.5:
                    b[off] ^= 0x10
                else:
                    del b[off:]
                data = bytes(b)
                kind += "+damaged"
            files.append(data)
            kinds.append(str(kind))
        lw, lh = int(rng.integers(8, 400)), int(rng.integers(8, 300))
        raw = util.logo_rgba(lw, lh, tile=int(rng.integers(8, 64)), radius=int(rng.integers(3, 40))) if rng.random() < 0.7 else util.noisy_rgba(lw, lh, int(rng.integers(1000)))
        d = M.Dropon()
        assert d.read_dropon_from_raw(raw, M.CS_RGBA, 255) == 0
        align = int(rng.choice([M.ALIGN_CENTER, M.ALIGN_TOP | M.ALIGN_LEFT, M.ALIGN_BOTTOM | M.ALIGN_RIGHT, M.ALIGN_TOP | M.ALIGN_RIGHT]))
        ox, oy = int(rng.integers(-40, 41)), int(rng.integers(-40, 41))
        rv, status, outs = capi.compose_batch(files, d, align, ox, oy, 0, nthreads=int(rng.integers(1, 9)))
        bad = sbad = 0
        for k, src in enumerate(files):
            j = M.Jpeg()
            r1 = j.read_jpeg_from_memory(src)
            want = None
            if r1 == 0:
                r1 = j.compose(d, align, ox, oy)
            if r1 == 0:
                r1, want = j.write_jpeg_to_memory(0)
            if "damaged" in kinds[k]:
                continue  # (libjpeg recovers from damage in its own ways; only that nothing crashes is asked here)
            if (status[k] == 0) != (r1 == 0):
                sbad += 1
            elif r1 == 0 and outs[k] != want:
                bad += 1
        report["batches"].append({"n": n, "uniform": uniform, "size": [w0, h0] if uniform else None, "rv": rv, "mismatches": bad, "status_mismatches": sbad,
                                  "kinds": {k: kinds.count(k) for k in sorted(set(kinds))}})
        report["images"] += n
        report["mismatches"] += bad
        report["status_mismatches"] += sbad
        done += n
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
