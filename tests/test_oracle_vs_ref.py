"""CPU: the oracle port against the UNMODIFIED reference run live (oracle/_ref), on seeded random
inputs beyond the committed golden vectors.  Skipped where oracle/_ref is absent."""
import numpy as np

import util


def test_convolve_bitexact(ref, port):
    rng = np.random.default_rng(0)
    for k in range(8):
        for l in range(8):
            for _ in range(6):
                x = rng.integers(-2000, 2000, 64).astype(np.float32)
                y0 = rng.normal(0, 100, 64).astype(np.float32)
                w = float(np.float32(rng.normal(0, 0.1)))
                ya, yb = y0.copy(), y0.copy()
                ref.convolve(x, ya, w, k, l)
                port.convolve(x, yb, w, k, l)
                assert np.array_equal(ya.view(np.uint32), yb.view(np.uint32)), (k, l)


def test_compile_dropon_bitexact(ref, port):
    from oracle import oracle_py as O

    layouts = [(3, [(2, 2), (1, 1), (1, 1)]), (3, [(2, 1), (1, 1), (1, 1)]), (3, [(1, 1)] * 3), (1, [(1, 1)]),
               (2, [(1, 1)] * 3), (3, [(4, 1), (1, 1), (1, 1)]), (3, [(1, 2), (1, 1), (1, 1)]),
               (3, [(2, 2), (2, 1), (1, 2)]), (3, [(4, 2), (1, 1), (2, 1)]), (3, [(3, 1), (1, 1), (1, 1)])]
    for tcs, samp in layouts:
        for (w, h, boff, crop) in [(48, 32, (0, 0), None), (50, 37, (3, 5), None), (64, 64, (7, 1), (5, 3, 40, 50))]:
            for dcs, nch in [(O.CS_RGBA, 4), (O.CS_RGB, 3), (O.CS_YCCA, 4), (O.CS_GRAY, 1), (O.CS_GRAYA, 2)]:
                raw = util.noisy_rgba(w, h, seed=w + h + dcs)[:, :, :nch]
                d = ref.dropon_from_raw(raw if nch > 1 else raw[:, :, 0], dcs, 200)
                rv, img, alp = ref.compile_dropon(d, tcs, samp, boff[0], boff[1], crop)
                rv2, D, W = port.compile_dropon(d.image3(), d.alpha3(), d.colorspace, O.make_layout(tcs, samp), boff[0],
                                                boff[1], crop)
                assert rv == rv2, (tcs, samp, dcs)
                if rv != 0:
                    continue
                for c in range(len(samp)):
                    assert np.array_equal(img[c], D[c].astype(np.float32))
                    wp = np.stack([port.alpha_weights(b) for b in W[c].reshape(-1, 64)]).reshape(W[c].shape)
                    assert np.array_equal(alp[c].view(np.uint32), wp.view(np.uint32))


def test_compose_and_effects_bitexact(ref, port):
    from oracle import oracle_py as O

    for subs, gray in [("420", False), ("422", False), ("444", False), ("444", True)]:
        data = util.jpeg_bytes(176, 144, subs, 85, seed=9, gray=gray)
        for (align, ox, oy) in [(5, -20, -10), (16, 0, 0), (10, 7, 3), (6, -13, 7)]:
            for raw, cs, blend in [(util.noisy_rgba(56, 40, 3), O.CS_RGBA, 255), (util.logo_rgba(64, 48, 32, 13), O.CS_RGBA, 255),
                                   (util.noisy_rgba(56, 40, 4)[:, :, :3], O.CS_RGB, 99)]:
                j = ref.read_jpeg(data)
                d = ref.dropon_from_raw(raw, cs, blend)
                info, samp = j.info(), j.sampling()
                planes = j.planes()
                q = [j.qtable(c) for c in range(info["ncomp"])]
                rv, _, _, _ = util.oracle_compose(port, planes, q, info["width"], info["height"], info["colorspace"], samp,
                                                  d.image3(), d.alpha3(), d.colorspace, d.blend, align, ox, oy)
                assert j.compose(d, align, ox, oy) == rv == 0
                for c, p in enumerate(j.planes()):
                    assert np.array_equal(p, planes[c]), (subs, gray, align, c)
        # effects
        for fx in ("lum", "tint", "gray", "pix"):
            j = ref.read_jpeg(data)
            info = j.info()
            planes = j.planes()
            ycc = info["colorspace"] == 3
            if fx == "lum":
                j.luminance(-77)
                if ycc:
                    ci = j.comp_info(0)
                    port.effect_add_dc(planes[0], ci["wreal"], ci["hreal"], j.qtable(0)[0], -77)
            elif fx == "tint":
                j.tint(15, -2000)
                if ycc:
                    for c, v in ((1, 15), (2, -2000)):
                        ci = j.comp_info(c)
                        port.effect_add_dc(planes[c], ci["wreal"], ci["hreal"], j.qtable(c)[0], v)
            elif fx == "gray":
                j.grayscale()
                if ycc:
                    for c in (1, 2):
                        ci = j.comp_info(c)
                        port.effect_zero(planes[c], ci["wreal"], ci["hreal"])
            else:
                j.pixelate()
                for c in range(info["ncomp"]):
                    ci = j.comp_info(c)
                    port.effect_pixelate(planes[c], ci["wreal"], ci["hreal"])
            for c, p in enumerate(j.planes()):
                assert np.array_equal(p, planes[c]), (subs, gray, fx, c)
