"""GPU: K5, the Huffman decoding of baseline scans on the device (k5_huffman_decode.cu), against libjpeg.

Replaces the entropy decoder behind the reference's mj_read_jpeg_from_memory (reference: src/image.c:33-118 -> libjpeg
jpeg_read_coefficients / jdhuff.c).  Bar: the planes equal libjpeg's coefficient arrays block for block, padding included.
"""
import io

import numpy as np
import pytest
from PIL import Image

import libmodjpeg_b200 as M
import util
from libmodjpeg_b200 import capi
from libmodjpeg_b200.batch import DeviceBatch

pytestmark = pytest.mark.gpu


def _decode_on_device(engine, files):
    """files: JPEG byte strings of ONE geometry.  Returns (planes per image, status, libjpeg's planes per image)."""
    want = []
    for d in files:
        j = M.Jpeg()
        assert j.read_jpeg_from_memory(d) == 0
        want.append(j.planes())
    shapes = [p.shape[:2] for p in want[0]]
    n = len(files)
    scans = [capi.scan_from_jpeg(d) for d in files]
    segs = [d[off:] for d, (_, off, _) in zip(files, scans)]  # to the end of the file: what follows the last MCU is ignored
    offsets = np.concatenate([[0], np.cumsum([len(s) + 13 for s in segs])[:-1]]).astype(np.uint64)  # odd gaps: no alignment is assumed
    lengths = np.array([len(s) for s in segs], np.uint32)
    blob = np.zeros(int(offsets[-1]) + len(segs[-1]) + 16, np.uint8)
    for o, s in zip(offsets, segs):
        blob[int(o):int(o) + len(s)] = np.frombuffer(s, np.uint8)
    data_dev = engine.device_alloc(blob.size)
    status_dev = engine.device_alloc(4 * n)
    batch = DeviceBatch(engine, shapes, n)
    batch.set_descs(np.ones((len(shapes), 64), np.uint16))
    junk = [np.full(p.shape, 0x5a5a, np.int16) for p in want[0]]
    for i in range(n):
        batch.upload_image(i, junk)  # the decoder owes every coefficient, zeros included
    try:
        engine.copy_h2d(data_dev, blob)
        engine.huffman_decode_batch_device(data_dev, offsets, lengths, n, scans[0][0], batch.descs_dev, status_dev)
        status = np.zeros(n, np.uint32)
        engine.copy_d2h(status, status_dev)
        engine.sync()
        got = [batch.download_image(i) for i in range(n)]
    finally:
        engine.device_free(data_dev)
        engine.device_free(status_dev)
    return got, status, want


CASES = [(64, 48, "420", False, 85), (67, 45, "420", False, 90), (120, 72, "422", False, 75), (50, 50, "444", False, 95), (33, 17, "444", True, 60),
         (17, 9, "420", False, 100), (1920, 1080, "420", False, 85), (1000, 601, "444", False, 97), (641, 479, "422", False, 50), (8, 8, "444", True, 85)]


@pytest.mark.parametrize("w,h,subs,gray,quality", CASES)
def test_planes_decoded_on_device_equal_libjpeg(engine, w, h, subs, gray, quality):
    files = [util.jpeg_bytes(w, h, subs, quality, seed=w + h + i, gray=gray) for i in range(3)]
    got, status, want = _decode_on_device(engine, files)
    assert not status.any()
    for i in range(len(files)):
        for c in range(len(want[i])):
            assert np.array_equal(got[i][c], want[i][c]), (i, c)


def test_optimised_tables_and_noise(engine):
    """the file's own (optimised) Huffman tables; incompressible content: long codes, many 0xFF bytes to un-stuff"""
    from PIL import ImageFile

    ImageFile.MAXBLOCK = max(ImageFile.MAXBLOCK, 1 << 22)  # optimize=True writes the scan in one piece
    rng = np.random.default_rng(4)
    files = []
    for i in range(4):
        img = Image.fromarray(rng.integers(0, 256, size=(211, 333, 3), dtype=np.uint8))
        buf = io.BytesIO()
        img.save(buf, "JPEG", quality=98, subsampling=0, optimize=True)
        files.append(buf.getvalue())
    # one scan description per call: the test feeds each file with its own tables
    for d in files:
        got, status, want = _decode_on_device(engine, [d])
        assert not status.any()
        for c in range(3):
            assert np.array_equal(got[0][c], want[0][c])


def test_corrupt_stream_is_handed_back(engine):
    d = bytearray(util.jpeg_bytes(320, 240, "420", 85, seed=3))
    _, off, _ = capi.scan_from_jpeg(bytes(d))
    good = bytes(d)
    del d[off + 2000:]  # truncated: fewer blocks than the frame announces
    d += b"\xff\xd9"
    got, status, _ = _decode_on_device(engine, [good])
    assert status[0] == 0
    # truncated file alone (libjpeg only warns about it; the device decoder must not pretend)
    j = M.Jpeg()
    j.read_jpeg_from_memory(bytes(d))
    shapes = [p.shape[:2] for p in j.planes()]
    scan, off2, _ = capi.scan_from_jpeg(bytes(d))
    seg = np.frombuffer(bytes(d)[off2:], np.uint8)
    data_dev = engine.device_alloc(seg.size + 16)
    status_dev = engine.device_alloc(4)
    batch = DeviceBatch(engine, shapes, 1)
    batch.set_descs(np.ones((3, 64), np.uint16))
    try:
        engine.copy_h2d(data_dev, seg.copy())
        engine.huffman_decode_batch_device(data_dev, np.array([0], np.uint64), np.array([seg.size], np.uint32), 1, scan, batch.descs_dev, status_dev)
        st = np.zeros(1, np.uint32)
        engine.copy_d2h(st, status_dev)
        engine.sync()
        assert st[0] != 0
    finally:
        engine.device_free(data_dev)
        engine.device_free(status_dev)


def test_compose_batch_window_never_leaves_the_device(engine):
    """mj_compose_batch on a window of one geometry: JPEG bytes up, K5 -> K2 -> K4 in HBM, JPEG bytes back -- the files equal
    read -> mj_compose -> write through the per-image API (host libjpeg on both ends), byte for byte"""
    raw = util.logo_rgba(200, 120, tile=64, radius=27)
    d = M.Dropon()
    assert d.read_dropon_from_raw(raw, M.CS_RGBA, 255) == 0
    jpegs = [util.jpeg_bytes(333, 250, "420", 70 + (i % 3) * 10, seed=500 + i) for i in range(37)]
    jpegs[11] = b"definitely not a jpeg"
    rv, status, outs = capi.compose_batch(jpegs, d, M.ALIGN_BOTTOM | M.ALIGN_RIGHT, -9, -6, 0, nthreads=4)
    assert rv == 0
    for k, src in enumerate(jpegs):
        if k == 11:
            assert status[k] == 5 and outs[k] is None
            continue
        assert status[k] == 0, k
        j = M.Jpeg()
        assert j.read_jpeg_from_memory(src) == 0
        assert j.compose(d, M.ALIGN_BOTTOM | M.ALIGN_RIGHT, -9, -6) == 0
        rvw, want = j.write_jpeg_to_memory(0)
        assert rvw == 0
        assert outs[k] == want, k


def test_damaged_files_do_not_bring_the_pipeline_down(engine):
    """bit flips and truncations in the entropy-coded data of a window that takes the device path: every image gets a status
    and (when it is 0) a file libjpeg can read; the intact neighbours are untouched"""
    raw = util.logo_rgba(96, 64, 32, 13)
    d = M.Dropon()
    assert d.read_dropon_from_raw(raw, M.CS_RGBA, 255) == 0
    rng = np.random.default_rng(11)
    good = [util.jpeg_bytes(256, 192, "420", 85, seed=600 + i) for i in range(24)]
    files = list(good)
    damaged = set()
    for i in range(0, 24, 3):
        b = bytearray(files[i])
        off = capi.scan_from_jpeg(files[i])[1]
        if i % 2 == 0:
            for _ in range(5):
                b[int(rng.integers(off, len(b) - 2))] ^= 1 << int(rng.integers(0, 8))
        else:
            del b[off + (len(b) - off) // 2:]
            b += b"\xff\xd9"
        files[i] = bytes(b)
        damaged.add(i)
    rv, status, outs = capi.compose_batch(files, d, M.ALIGN_CENTER, 0, 0, 0, nthreads=4)
    assert rv == 0
    for i in range(24):
        if i in damaged:
            if status[i] == 0:
                j = M.Jpeg()
                assert j.read_jpeg_from_memory(outs[i]) == 0
            continue
        assert status[i] == 0
        j = M.Jpeg()
        assert j.read_jpeg_from_memory(good[i]) == 0 and j.compose(d, M.ALIGN_CENTER, 0, 0) == 0
        assert j.write_jpeg_to_memory(0)[1] == outs[i], i


def test_batch_pipeline_random_sweep(engine):
    """profiles/fuzz_batch.py in small: random sizes / samplings / qualities, batches of one geometry (all-device path) and mixed
    ones (progressive, optimised, grayscale, damaged inputs) -- every file mj_compose_batch writes equals the per-image calls'"""
    import json
    import os
    import subprocess
    import sys

    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    out = subprocess.run([sys.executable, os.path.join(root, "profiles", "fuzz_batch.py"), "160", "31"], check=True, capture_output=True, text=True).stdout
    rep = json.loads(out)
    assert rep["images"] >= 160 and rep["mismatches"] == 0 and rep["status_mismatches"] == 0
    assert any(b["uniform"] for b in rep["batches"]) and any(not b["uniform"] for b in rep["batches"])
