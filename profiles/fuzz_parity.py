#!/usr/bin/env python
"""Randomised parity sweep of the drop-in API against the unmodified reference (oracle/_ref), both run here on the
same inputs: random image sizes, samplings, qualities, dropon sizes / colourspaces / alpha shapes, alignments and
offsets (incl. partly and fully off-image), followed by a random chain of effects.  Reports how many quantised
coefficients differ and by how much; uniform-alpha / opaque dropons and the effects have to be bit-identical.
usage (GPU box): python profiles/fuzz_parity.py [cases] [seed] > gpurun_out/fuzz_parity.json"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import libmodjpeg_b200 as M  # noqa: E402
import util  # noqa: E402
from oracle import oracle_py as O  # noqa: E402  (the checker)

ALIGNS = [M.ALIGN_LEFT, M.ALIGN_RIGHT, M.ALIGN_CENTER]
VALIGNS = [M.ALIGN_TOP, M.ALIGN_BOTTOM, M.ALIGN_CENTER]


def random_dropon(rng, kind, aligned):
    w, h = int(rng.integers(1, 200)), int(rng.integers(1, 160))
    if aligned:  # whole MCUs: no partly covered edge blocks
        w, h = 16 * int(rng.integers(1, 12)), 16 * int(rng.integers(1, 10))
    rgb = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    if kind == "rgb_blend":  # uniform alpha through `blend`
        return rgb, M.CS_RGB, int(rng.integers(1, 256)), True
    yy, xx = np.mgrid[0:h, 0:w]
    if kind == "opaque":
        a = np.full((h, w), 255, np.uint8)
    elif kind == "binary":
        a = (((xx // 9 + yy // 7) % 2) * 255).astype(np.uint8)
    elif kind == "gradient":
        a = ((xx * 255) // max(1, w - 1)).astype(np.uint8)
    elif kind == "disc":
        r = np.hypot(xx - w / 2, yy - h / 2)
        a = np.clip((min(w, h) / 2 - r) * 12, 0, 255).astype(np.uint8)
    else:  # noise
        a = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
    return np.dstack([rgb, a]), M.CS_RGBA, 255, kind == "opaque"


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 2026
    rng = np.random.default_rng(seed)
    ref = O.Reference()
    tot = {"cases": 0, "composes_visible": 0, "coefficients_changed": 0, "coefficients_differing": 0, "max_abs_diff": 0,
           "exact_class_cases": 0, "exact_class_differing": 0, "effect_calls": 0, "effect_differing": 0, "return_code_mismatch": 0}
    worst = []
    for i in range(cases):
        W_, H_ = int(rng.integers(8, 420)), int(rng.integers(8, 330))
        gray = bool(rng.random() < 0.15)
        subs = str(rng.choice(["444", "422", "420"]))
        q = int(rng.integers(30, 99))
        data = util.jpeg_bytes(W_, H_, subs, q, seed=int(rng.integers(1 << 30)), gray=gray)
        kind = str(rng.choice(["rgb_blend", "opaque", "binary", "gradient", "disc", "noise"]))
        aligned = bool(rng.random() < 0.3)
        raw, cs, blend, exact = random_dropon(rng, kind, aligned)
        align = int(rng.choice(ALIGNS)) | int(rng.choice(VALIGNS))
        ox, oy = int(rng.integers(-220, 440)), int(rng.integers(-180, 340))
        if rng.random() < 0.5:
            ox, oy = int(rng.integers(-30, 30)), int(rng.integers(-30, 30))
        if aligned:
            align, ox, oy = M.ALIGN_TOP | M.ALIGN_LEFT, 16 * int(rng.integers(-3, 12)), 16 * int(rng.integers(-3, 10))
        # bit-exactness is claimed for blocks with uniform alpha: a uniform-alpha dropon has them everywhere only if
        # it covers whole MCUs (otherwise the padding around it makes the edge blocks non-uniform, i.e. float-blended)
        hf, vf = (1, 1) if gray or subs == "444" else ((2, 1) if subs == "422" else (2, 2))
        g = M.geometry(W_, H_, hf, vf, raw.shape[1], raw.shape[0], align, ox, oy)
        exact = exact and g["visible"] and g["blockoffset_x"] == 0 and g["blockoffset_y"] == 0 and \
            (g["crop_x"] + 0) % (8 * hf) == 0 and (g["crop_y"] + 0) % (8 * vf) == 0 and g["crop_w"] % (8 * hf) == 0 and g["crop_h"] % (8 * vf) == 0
        jr = ref.read_jpeg(data)
        dr = ref.dropon_from_raw(raw, cs, blend)
        rv_r = jr.compose(dr, align, ox, oy)
        j = M.Jpeg()
        assert j.read_jpeg_from_memory(data) == 0
        before = j.planes()
        d = M.Dropon()
        assert d.read_dropon_from_raw(raw, cs, blend) == 0
        rv = j.compose(d, align, ox, oy)
        tot["cases"] += 1
        if rv != rv_r:
            tot["return_code_mismatch"] += 1
            worst.append({"case": i, "rv": rv, "rv_reference": rv_r})
            continue
        got, want = j.planes(), jr.planes()
        changed = differ = mx = 0
        for a, b, c in zip(got, want, before):
            dd = a.astype(np.int32) - b.astype(np.int32)
            differ += int((dd != 0).sum())
            mx = max(mx, int(np.abs(dd).max()))
            changed += int((b != c).sum())
        tot["composes_visible"] += int(changed > 0)
        tot["coefficients_changed"] += changed
        tot["coefficients_differing"] += differ
        tot["max_abs_diff"] = max(tot["max_abs_diff"], mx)
        if exact:
            tot["exact_class_cases"] += 1
            tot["exact_class_differing"] += differ
        if differ:
            worst.append({"case": i, "size": [W_, H_], "subs": "gray" if gray else subs, "q": q, "kind": kind, "differing": differ, "changed": changed, "max": mx})
        # a random chain of effects on top; start both sides from the reference's planes so that the effects are
        # compared on identical inputs
        if differ == 0:
            for _ in range(int(rng.integers(0, 4))):
                fx = str(rng.choice(["grayscale", "pixelate", "tint", "luminance"]))
                if fx == "grayscale":
                    ra, rb = jr.grayscale(), j.effect_grayscale()
                elif fx == "pixelate":
                    ra, rb = jr.pixelate(), j.effect_pixelate()
                elif fx == "tint":
                    cb, cr = int(rng.integers(-300, 300)), int(rng.integers(-300, 300))
                    ra, rb = jr.tint(cb, cr), j.effect_tint(cb, cr)
                else:
                    v = int(rng.integers(-3000, 3000))
                    ra, rb = jr.luminance(v), j.effect_luminance(v)
                tot["effect_calls"] += 1
                bad = int(ra != rb) + sum(int((a != b).sum()) for a, b in zip(j.planes(), jr.planes()))
                tot["effect_differing"] += bad
                if bad:
                    worst.append({"case": i, "effect": fx, "differing": bad})
                    break
    tot["differing_rate_of_changed"] = tot["coefficients_differing"] / max(1, tot["coefficients_changed"])
    print(json.dumps({"seed": seed, "totals": tot, "cases_with_differences": worst[:40]}, indent=1))


if __name__ == "__main__":
    main()
