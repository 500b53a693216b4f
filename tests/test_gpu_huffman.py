"""GPU: K4, the Huffman coding of baseline scans on the device (k4_huffman.cu), against libjpeg and the oracle.

Replaces the entropy encoder behind the reference's mj_write_jpeg_to_memory (reference: src/image.c:120-209 ->
libjpeg jpeg_write_coefficients / jchuff.c).  Bar: byte-identical files.
"""
import os
import sys

import numpy as np
import pytest

import libmodjpeg_b200 as M
import util
from libmodjpeg_b200 import capi
from libmodjpeg_b200.batch import DeviceBatch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import huff_oracle as H  # noqa: E402
from test_huffman_oracle import oracle_segment  # noqa: E402

pytestmark = pytest.mark.gpu

CASES = [(64, 48, "420", False, 85), (67, 45, "420", False, 90), (120, 72, "422", False, 75), (50, 50, "444", False, 95), (33, 17, "444", True, 60),
         (17, 9, "420", False, 100), (1920, 1080, "420", False, 85), (1000, 601, "444", False, 97), (641, 479, "422", False, 50), (8, 8, "444", True, 85)]


@pytest.mark.parametrize("w,h,subs,gray,quality", CASES)
def test_file_written_on_device_equals_libjpeg(engine, w, h, subs, gray, quality):
    j = M.Jpeg()
    assert j.read_jpeg_from_memory(util.jpeg_bytes(w, h, subs, quality, seed=w + h, gray=gray)) == 0
    rv, want = j.write_jpeg_to_memory(0)
    assert rv == 0
    rv, got = j.write_jpeg_to_memory_device(0)
    assert rv == 0
    assert got == want


def test_file_with_markers_and_after_compose(engine):
    """saved COM / APPn markers are re-emitted by libjpeg itself on both paths; coefficients changed by mj_compose"""
    from PIL import Image
    import io

    img = Image.fromarray(util.photo(300, 200, 7))
    buf = io.BytesIO()
    img.save(buf, "JPEG", quality=88, subsampling=2, comment=b"hello, marker", dpi=(72, 72))
    j = M.Jpeg()
    assert j.read_jpeg_from_memory(buf.getvalue()) == 0
    d = M.Dropon()
    assert d.read_dropon_from_raw(util.logo_rgba(96, 64, 32, 13), capi.CS_RGBA, 255) == 0
    assert j.compose(d, capi.ALIGN_BOTTOM | capi.ALIGN_RIGHT, -3, -5) == 0
    rv, want = j.write_jpeg_to_memory(0)
    rv2, got = j.write_jpeg_to_memory_device(0)
    assert rv == 0 and rv2 == 0
    assert got == want
    # and what does not apply is refused, not mis-coded
    for opt in (capi.OPTION_OPTIMIZE, capi.OPTION_PROGRESSIVE):
        assert j.write_jpeg_to_memory_device(opt)[0] != 0


def test_uncodable_coefficient_falls_back_to_libjpeg(engine, monkeypatch):
    j = M.Jpeg()
    assert j.read_jpeg_from_memory(util.jpeg_bytes(64, 64, "444", 85, seed=2)) == 0
    p = j.plane(0).copy()
    p[1, 1, 7] = 1024  # needs 11 bits: no baseline AC code
    j.set_plane(0, p)
    assert j.write_jpeg_to_memory_device(0)[0] != 0
    # a DC step of more than 11 bits likewise
    p[1, 1, 7] = 0
    p[2, 2, 0], p[2, 3, 0] = -2000, 2047
    j.set_plane(0, p)
    assert j.write_jpeg_to_memory_device(0)[0] != 0


def test_batch_device_segments_equal_oracle_and_libjpeg(engine):
    """mjx_huffman_encode_batch_device: many images per call, planes resident in HBM, one output slab"""
    W_, H_, n = 203, 117, 37
    dec, segs = [], []
    for i in range(5):
        j = M.Jpeg()
        assert j.read_jpeg_from_memory(util.jpeg_bytes(W_, H_, "420", 60 + 8 * i, seed=40 + i)) == 0
        rv, data = j.write_jpeg_to_memory(0)
        assert rv == 0
        seg = H.split_jpeg(data)[1]
        if i == 0:
            assert oracle_segment(j) == seg
        dec.append(j)
        segs.append(seg)
    info, samp = dec[0].info(), dec[0].sampling()
    planes0 = dec[0].planes()
    shapes = [p.shape[:2] for p in planes0]
    real = [(dec[0].comp_info(c)["wreal"], dec[0].comp_info(c)["hreal"]) for c in range(info["ncomp"])]
    batch = DeviceBatch(engine, shapes, n, real_dims=real)
    batch.set_descs(np.stack([np.stack([dec[i % 5].qtable(c) for c in range(info["ncomp"])]) for i in range(n)]))
    for i in range(n):
        batch.upload_image(i, dec[i % 5].planes())
    scan = capi.standard_scan(W_, H_, samp)
    cap = 64 * 1024
    out_dev = engine.device_alloc(n * cap)
    sizes_dev = engine.device_alloc(n * 4)
    try:
        engine.huffman_encode_batch_device(batch.descs_dev, n, scan, out_dev, cap, sizes_dev)
        sizes = np.zeros(n, np.uint32)
        engine.copy_d2h(sizes, sizes_dev)
        out = np.zeros(n * cap, np.uint8)
        engine.copy_d2h(out, out_dev)
        engine.sync()
        for i in range(n):
            assert sizes[i] == len(segs[i % 5]), i
            assert out[i * cap:i * cap + sizes[i]].tobytes() == segs[i % 5], i
        # a slab too small for some of the images: those come back as "not coded", the others are untouched
        small = min(len(s) for s in segs) + 8
        engine.huffman_encode_batch_device(batch.descs_dev, n, scan, out_dev, small, sizes_dev)
        engine.copy_d2h(sizes, sizes_dev)
        engine.copy_d2h(out, out_dev)
        engine.sync()
        for i in range(n):
            if len(segs[i % 5]) <= small:
                assert out[i * small:i * small + sizes[i]].tobytes() == segs[i % 5]
            else:
                assert sizes[i] == 0xFFFFFFFF
    finally:
        engine.device_free(out_dev)
        engine.device_free(sizes_dev)


def test_write_path_env_switch(engine, tmp_path):
    """MJX_GPU_HUFFMAN=1 makes mj_write_jpeg_to_memory itself take the device path (read once per process: a fresh one)"""
    import subprocess

    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import libmodjpeg_b200 as M, util\n"
        "j = M.Jpeg(); assert j.read_jpeg_from_memory(util.jpeg_bytes(333, 222, '420', 85, seed=1)) == 0\n"
        "rv, a = j.write_jpeg_to_memory(0); assert rv == 0\n"
        "rv, b = j.write_jpeg_to_memory(1); assert rv == 0\n"  # optimised: libjpeg, whatever the switch says
        "open(%r, 'wb').write(a); open(%r, 'wb').write(b)\n"
    )
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    outs = {}
    for flag in ("0", "1"):
        pa, pb = str(tmp_path / f"a{flag}.jpg"), str(tmp_path / f"b{flag}.jpg")
        env = dict(os.environ, MJX_GPU_HUFFMAN=flag)
        subprocess.run([sys.executable, "-c", code % (root, os.path.dirname(os.path.abspath(__file__)), pa, pb)], check=True, env=env)
        outs[flag] = (open(pa, "rb").read(), open(pb, "rb").read())
    assert outs["0"] == outs["1"]


def test_compose_batch_device_resident_pipeline(engine, tmp_path):
    """mj_compose_batch with MJX_GPU_HUFFMAN=1: whole planes up, K2 + K4 in HBM, only the segments come back -- the files
    equal the ones of the ordinary pipeline (host libjpeg encode), image by image; mixed geometries (4:2:0, 4:4:4 with an odd
    size, grayscale smaller than the dropon) and an undecodable input in one batch"""
    import pickle
    import subprocess

    code = (
        "import sys, pickle; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import numpy as np, libmodjpeg_b200 as M, util\n"
        "from libmodjpeg_b200 import capi\n"
        "d = M.Dropon(); assert d.read_dropon_from_raw(util.logo_rgba(200, 120, tile=64, radius=27), M.CS_RGBA, 255) == 0\n"
        "jp = [util.jpeg_bytes(320, 240, '420', 85, seed=300 + i) for i in range(5)] + [util.jpeg_bytes(275, 203, '444', 90, seed=310 + i) for i in range(3)]\n"
        "jp += [b'definitely not a jpeg', util.jpeg_bytes(320, 240, '420', 100, seed=77), util.jpeg_bytes(96, 64, '444', 85, seed=5, gray=True)]\n"
        "order = [0, 5, 1, 8, 6, 2, 9, 3, 7, 10, 4]\n"
        "rv, status, outs = capi.compose_batch([jp[i] for i in order], d, M.ALIGN_CENTER, 7, -5, 0, nthreads=4)\n"
        "pickle.dump((rv, list(status), outs), open(%r, 'wb'))\n"
    )
    here = os.path.dirname(os.path.abspath(__file__))
    res = {}
    for flag in ("0", "1"):
        path = str(tmp_path / f"r{flag}.pkl")
        subprocess.run([sys.executable, "-c", code % (os.path.join(here, ".."), here, path)], check=True, env=dict(os.environ, MJX_GPU_HUFFMAN=flag))
        res[flag] = pickle.load(open(path, "rb"))
    assert res["0"][0] == 0 and res["1"][0] == 0
    assert res["0"][1] == res["1"][1]
    assert sum(1 for s in res["1"][1] if s == 0) == 10
    for a, b in zip(res["0"][2], res["1"][2]):
        assert a == b
