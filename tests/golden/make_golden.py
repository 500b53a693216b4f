"""Generate tests/golden/golden.npz from the UNMODIFIED reference build (oracle/_ref).

Run in the build container (needs /root/reference for oracle/build_ref.sh):
    python tests/golden/make_golden.py
The .npz holds inputs AND the reference's outputs, so the GPU box (no /root/reference) can
check the oracle port and the CUDA path against them.  image.jpg / dropon.png /
image_dropon.jpg are the reference's own README fixture (src/contrib/images).
"""
import io
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import oracle_py as O  # noqa: E402
import util  # noqa: E402

GEOMETRY_CASES = [(5, -20, -10), (10, 30, 20), (16, 0, 0), (5, 155, 120), (5, 160, 0), (5, -48, 0), (5, -47, 0),
                  (5, 3, 9), (6, -13, 7), (9, 5, -11), (16, 1000, 0), (16, -57, 49)]


def main():
    ref = O.Reference()
    out = {}
    rng = np.random.default_rng(2024)

    # ---- C1: the README fixture, TOP|LEFT, offset 0,0 -----------------------------------
    image = open(os.path.join(HERE, "image.jpg"), "rb").read()
    dropon = np.asarray(Image.open(os.path.join(HERE, "dropon.png")).convert("RGBA"))
    out["c1_dropon_rgba"] = dropon
    j = ref.read_jpeg(image)
    d = ref.dropon_from_raw(dropon, O.CS_RGBA, 255)
    before = j.planes()
    assert j.compose(d, O.ALIGN_TOP | O.ALIGN_LEFT, 0, 0) == 0
    after = j.planes()
    for c in range(3):
        out[f"c1_after_{c}"] = after[c]
        out[f"c1_changed_blocks_{c}"] = np.int64((after[c] != before[c]).any(-1).sum())
    gold = ref.read_jpeg(open(os.path.join(HERE, "image_dropon.jpg"), "rb").read())
    out["c1_readme_luma"] = gold.plane(0)
    out["c1_written_jpeg"] = np.frombuffer(j.write(0), np.uint8)

    # ---- geometry table on a 160x128 4:2:0 image with a 48x32 dropon ----------
    gj = util.jpeg_bytes(160, 128, "420", 85, seed=5)
    gd = util.noisy_rgba(48, 32, seed=11)
    out["geo_jpeg"] = np.frombuffer(gj, np.uint8)
    out["geo_dropon"] = gd
    out["geo_cases"] = np.array(GEOMETRY_CASES, np.int32)
    for i, (align, ox, oy) in enumerate(GEOMETRY_CASES):
        j = ref.read_jpeg(gj)
        d = ref.dropon_from_raw(gd, O.CS_RGBA, 255)
        assert j.compose(d, align, ox, oy) == 0
        for c, p in enumerate(j.planes()):
            out[f"geo_{i}_after_{c}"] = p

    # ---- compose KATs over layouts / dropon formats ---------------------------------------
    kat = []
    for name, subs, gray in [("420", "420", False), ("422", "422", False), ("444", "444", False), ("gray", "444", True)]:
        jb = util.jpeg_bytes(96, 80, subs, 85, seed=21, gray=gray)
        out[f"kat_{name}_jpeg"] = np.frombuffer(jb, np.uint8)
        for dn, raw, cs, blend in [("rgba_noise", util.noisy_rgba(40, 24, 31), O.CS_RGBA, 255),
                                   ("rgba_logo", util.logo_rgba(64, 48, tile=32, radius=13), O.CS_RGBA, 255),
                                   ("rgb_b77", util.noisy_rgba(40, 24, 32)[:, :, :3], O.CS_RGB, 77),
                                   ("rgb_b255", util.noisy_rgba(40, 24, 33)[:, :, :3], O.CS_RGB, 255),
                                   ("ycca", util.noisy_rgba(40, 24, 34), O.CS_YCCA, 255),
                                   ("graya", util.noisy_rgba(40, 24, 35)[:, :, :2], O.CS_GRAYA, 255)]:
            j = ref.read_jpeg(jb)
            d = ref.dropon_from_raw(raw, cs, blend)
            rv = j.compose(d, O.ALIGN_CENTER, 3, -2)
            out[f"kat_{name}_{dn}_raw"] = raw
            out[f"kat_{name}_{dn}_rv"] = np.int64(rv)
            if rv == 0:
                for c, p in enumerate(j.planes()):
                    out[f"kat_{name}_{dn}_after_{c}"] = p
            kat.append((name, dn, cs, blend))
    out["kat_index"] = np.array([f"{a}|{b}|{c}|{d}" for a, b, c, d in kat])

    # ---- compile KAT: mj_compile_dropon outputs -------------------------------------------
    raw = util.noisy_rgba(50, 37, 41)
    out["compile_raw"] = raw
    d = ref.dropon_from_raw(raw, O.CS_RGBA, 255)
    for name, cs, samp in [("420", 3, [(2, 2), (1, 1), (1, 1)]), ("422", 3, [(2, 1), (1, 1), (1, 1)]),
                           ("444", 3, [(1, 1), (1, 1), (1, 1)]), ("gray", 1, [(1, 1)]), ("rgb", 2, [(1, 1)] * 3),
                           ("411", 3, [(4, 1), (1, 1), (1, 1)]), ("440", 3, [(1, 2), (1, 1), (1, 1)])]:
        rv, img, alp = ref.compile_dropon(d, cs, samp, 3, 5, (2, 1, 45, 30))
        assert rv == 0
        for c in range(len(samp)):
            out[f"compile_{name}_D_{c}"] = img[c].astype(np.int16)
            out[f"compile_{name}_w_{c}"] = alp[c]  # float weights (dropon.c:548-566)

    # ---- effects on the README image ------------------------------------------------------
    for name, fn in [("luminance40", lambda j: j.luminance(40)), ("tint30m30", lambda j: j.tint(30, -30)),
                     ("grayscale", lambda j: j.grayscale()), ("pixelate", lambda j: j.pixelate()),
                     ("luminance_wrap", lambda j: j.luminance(2 ** 31 - 1)), ("tint_big", lambda j: j.tint(-5000, 70000))]:
        j = ref.read_jpeg(image)
        assert fn(j) == 0
        for c, p in enumerate(j.planes()):
            out[f"fx_{name}_{c}"] = p

    # ---- host I/O: read -> write with each option set; dropon ingest ----------------------
    for opt in (0, 1, 2, 3):
        j = ref.read_jpeg(image)
        out[f"rw_opt{opt}"] = np.frombuffer(j.write(opt), np.uint8)
    for dn, raw, cs, blend in [("rgba", util.noisy_rgba(9, 7, 51), O.CS_RGBA, 200), ("rgb", util.noisy_rgba(9, 7, 52)[:, :, :3], O.CS_RGB, 77),
                               ("ycc", util.noisy_rgba(9, 7, 53)[:, :, :3], O.CS_YCC, 300), ("ycca", util.noisy_rgba(9, 7, 54), O.CS_YCCA, 1),
                               ("gray", util.noisy_rgba(9, 7, 55)[:, :, 0], O.CS_GRAY, -5), ("graya", util.noisy_rgba(9, 7, 56)[:, :, :2], O.CS_GRAYA, 255)]:
        d = ref.dropon_from_raw(raw, cs, blend)
        out[f"ingest_{dn}_raw"] = raw
        out[f"ingest_{dn}_args"] = np.array([cs, blend], np.int64)
        out[f"ingest_{dn}_image3"] = d.image3()
        out[f"ingest_{dn}_alpha3"] = d.alpha3()
        out[f"ingest_{dn}_meta"] = np.array([d.width, d.height, d.colorspace, d.blend], np.int64)

    np.savez_compressed(os.path.join(HERE, "golden.npz"), **out)
    print("wrote golden.npz with", len(out), "arrays,", os.path.getsize(os.path.join(HERE, "golden.npz")), "bytes")


if __name__ == "__main__":
    main()
