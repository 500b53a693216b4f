"""CPU: the oracle port (oracle/mj_oracle.c) against the committed golden vectors, which were
produced by the UNMODIFIED reference build (tests/golden/make_golden.py).  This is what pins the
oracle on a box without /root/reference."""
import os

import numpy as np
import pytest
from PIL import Image

import util

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "golden.npz"))


def _planes_from_jpeg(lib_cls, data):
    j = lib_cls()
    assert j.read_jpeg_from_memory(bytes(data)) == 0
    return j


def _decode(built, data):
    """coefficient planes + tables of a JPEG through the product's host library (libjpeg only, no GPU)"""
    from libmodjpeg_b200 import Jpeg

    j = Jpeg()
    assert j.read_jpeg_from_memory(bytes(data)) == 0
    info = j.info()
    return j, info, j.sampling(), j.planes(), [j.qtable(c) for c in range(info["ncomp"])]


def test_ingest_matches_reference():
    for dn in ("rgba", "rgb", "ycc", "ycca", "gray", "graya"):
        cs, blend = G[f"ingest_{dn}_args"]
        i3, a3, stored, b = util.ingest_raw(G[f"ingest_{dn}_raw"], int(cs), int(blend))
        assert np.array_equal(i3, G[f"ingest_{dn}_image3"]), dn
        assert np.array_equal(a3, G[f"ingest_{dn}_alpha3"]), dn
        w, h, scs, sb = G[f"ingest_{dn}_meta"]
        assert (stored, b) == (scs, sb) and i3.shape[:2] == (h, w), dn


def test_c1_readme_fixture(built, port):
    """config 1: image.jpg + dropon.png, TOP|LEFT -- port == reference build bit for bit, and the
    luma plane equals the reference's own README result image_dropon.jpg."""
    data = open(os.path.join(HERE, "golden", "image.jpg"), "rb").read()
    j, info, samp, planes, q = _decode(built, data)
    before = [p.copy() for p in planes]
    i3, a3, cs, blend = util.ingest_raw(G["c1_dropon_rgba"], 2, 255)
    rv, g, D, W = util.oracle_compose(port, planes, q, info["width"], info["height"], info["colorspace"], samp, i3, a3, cs,
                                      blend, 4 | 1, 0, 0)
    assert rv == 0
    for c in range(3):
        assert np.array_equal(planes[c], G[f"c1_after_{c}"]), c
        assert int((planes[c] != before[c]).any(-1).sum()) == int(G[f"c1_changed_blocks_{c}"])
    assert [int(G[f"c1_changed_blocks_{c}"]) for c in range(3)] == [93, 30, 26]
    assert np.array_equal(planes[0], G["c1_readme_luma"])


def test_geometry_table(built, port):
    data = G["geo_jpeg"].tobytes()
    i3, a3, cs, blend = util.ingest_raw(G["geo_dropon"], 2, 255)
    expected_luma_changed = {(5, -20, -10): 12, (10, 30, 20): 6, (16, 0, 0): 24, (5, 155, 120): 1, (5, 160, 0): 0,
                             (5, -48, 0): 0, (5, -47, 0): 4}
    for i, (align, ox, oy) in enumerate(G["geo_cases"]):
        j, info, samp, planes, q = _decode(built, data)
        before = planes[0].copy()
        rv, g, _, _ = util.oracle_compose(port, planes, q, info["width"], info["height"], info["colorspace"], samp, i3, a3,
                                          cs, blend, int(align), int(ox), int(oy))
        assert rv == 0
        for c in range(3):
            assert np.array_equal(planes[c], G[f"geo_{i}_after_{c}"]), (i, c)
        key = (int(align), int(ox), int(oy))
        if key in expected_luma_changed:
            assert int((planes[0] != before).any(-1).sum()) == expected_luma_changed[key], key


def test_compose_kats(built, port):
    for entry in G["kat_index"]:
        name, dn, cs, blend = str(entry).split("|")
        data = G[f"kat_{name}_jpeg"].tobytes()
        j, info, samp, planes, q = _decode(built, data)
        i3, a3, scs, sblend = util.ingest_raw(G[f"kat_{name}_{dn}_raw"], int(cs), int(blend))
        rv, g, _, _ = util.oracle_compose(port, planes, q, info["width"], info["height"], info["colorspace"], samp, i3, a3,
                                          scs, sblend, 16, 3, -2)
        assert rv == int(G[f"kat_{name}_{dn}_rv"]), (name, dn)
        if rv == 0:
            for c in range(info["ncomp"]):
                assert np.array_equal(planes[c], G[f"kat_{name}_{dn}_after_{c}"]), (name, dn, c)


def test_compile_kats(port):
    from oracle import oracle_py as O

    i3, a3, cs, blend = util.ingest_raw(G["compile_raw"], 2, 255)
    for name, tcs, samp in [("420", 3, [(2, 2), (1, 1), (1, 1)]), ("422", 3, [(2, 1), (1, 1), (1, 1)]),
                            ("444", 3, [(1, 1), (1, 1), (1, 1)]), ("gray", 1, [(1, 1)]), ("rgb", 2, [(1, 1)] * 3),
                            ("411", 3, [(4, 1), (1, 1), (1, 1)]), ("440", 3, [(1, 2), (1, 1), (1, 1)])]:
        rv, D, W = port.compile_dropon(i3, a3, cs, O.make_layout(tcs, samp), 3, 5, (2, 1, 45, 30))
        assert rv == 0
        for c in range(len(samp)):
            assert np.array_equal(D[c], G[f"compile_{name}_D_{c}"]), (name, c)
            w = np.stack([port.alpha_weights(b) for b in W[c].reshape(-1, 64)]).reshape(W[c].shape)
            assert np.array_equal(w.view(np.uint32), G[f"compile_{name}_w_{c}"].view(np.uint32)), (name, c)


def test_effect_kats(built, port):
    data = open(os.path.join(HERE, "golden", "image.jpg"), "rb").read()
    cases = {"luminance40": [(0, 40)], "tint30m30": [(1, 30), (2, -30)], "luminance_wrap": [(0, 2 ** 31 - 1)],
             "tint_big": [(1, -5000), (2, 70000)]}
    for name, adds in cases.items():
        j, info, samp, planes, q = _decode(built, data)
        for c, v in adds:
            ci = j.comp_info(c)
            port.effect_add_dc(planes[c], ci["wreal"], ci["hreal"], q[c][0], v)
        for c in range(3):
            assert np.array_equal(planes[c], G[f"fx_{name}_{c}"]), (name, c)
    j, info, samp, planes, q = _decode(built, data)
    for c in (1, 2):
        ci = j.comp_info(c)
        port.effect_zero(planes[c], ci["wreal"], ci["hreal"])
    for c in range(3):
        assert np.array_equal(planes[c], G[f"fx_grayscale_{c}"])
    j, info, samp, planes, q = _decode(built, data)
    for c in range(3):
        ci = j.comp_info(c)
        port.effect_pixelate(planes[c], ci["wreal"], ci["hreal"])
        assert np.array_equal(planes[c], G[f"fx_pixelate_{c}"])
