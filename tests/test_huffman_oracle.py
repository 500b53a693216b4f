"""CPU: the Huffman-encoder oracle (oracle/huff_oracle.py) against the real libjpeg.

mj_write_jpeg_to_memory of the drop-in library's HOST path is libjpeg's jpeg_write_coefficients (reference:
src/image.c:120-209); its files are the fixture the oracle is pinned on.  The GPU encoder (k4_huffman.cu) is then
compared with both in tests/test_gpu_huffman.py.
"""
import os
import sys

import numpy as np
import pytest

import libmodjpeg_b200 as M
import util
from libmodjpeg_b200 import capi

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import huff_oracle as H  # noqa: E402

CASES = [(64, 48, "420", False, 85), (67, 45, "420", False, 90), (120, 72, "422", False, 75), (50, 50, "444", False, 95), (33, 17, "444", True, 60),
         (17, 9, "420", False, 100), (200, 120, "420", False, 30)]


def std_tables():
    dc = [H.derive(*capi.STD_DC_LUMA), H.derive(*capi.STD_DC_CHROMA)]
    ac = [H.derive(*capi.STD_AC_LUMA), H.derive(*capi.STD_AC_CHROMA)]
    return dc, ac


def oracle_segment(j):
    info, samp = j.info(), j.sampling()
    planes = j.planes()
    real = [(j.comp_info(c)["wreal"], j.comp_info(c)["hreal"]) for c in range(info["ncomp"])]
    dc, ac = std_tables()
    return H.entropy_segment(planes, real, samp, info["width"], info["height"], dc, ac, [0] + [1] * (info["ncomp"] - 1))


@pytest.mark.parametrize("w,h,subs,gray,quality", CASES)
def test_oracle_segment_equals_libjpeg(built, w, h, subs, gray, quality):
    j = M.Jpeg()
    assert j.read_jpeg_from_memory(util.jpeg_bytes(w, h, subs, quality, seed=w + h, gray=gray)) == 0
    rv, data = j.write_jpeg_to_memory(0)
    assert rv == 0
    head, seg, tail = H.split_jpeg(data)
    assert oracle_segment(j) == seg


def test_oracle_segment_after_coefficient_edits(built):
    """planes edited in place (what a compose leaves behind), including the padding blocks libjpeg regenerates"""
    j = M.Jpeg()
    assert j.read_jpeg_from_memory(util.jpeg_bytes(75, 41, "420", 85, seed=5)) == 0
    rng = np.random.default_rng(9)
    for c, p in enumerate(j.planes()):
        p = p.copy()
        p[:, :, 0] += rng.integers(-40, 40, size=p.shape[:2], dtype=np.int16)
        p[:, :, 1:] = np.where(rng.random(p[:, :, 1:].shape) < 0.2, rng.integers(-300, 300, size=p[:, :, 1:].shape), p[:, :, 1:]).astype(np.int16)
        j.set_plane(c, p)
    rv, data = j.write_jpeg_to_memory(0)
    assert rv == 0
    assert oracle_segment(j) == H.split_jpeg(data)[1]


def test_oracle_refuses_what_libjpeg_refuses(built):
    j = M.Jpeg()
    assert j.read_jpeg_from_memory(util.jpeg_bytes(32, 32, "444", 85, seed=2)) == 0
    p = j.plane(0).copy()
    p[1, 1, 7] = 1024  # needs 11 bits: no baseline AC code
    j.set_plane(0, p)
    with pytest.raises(H.NotCodable):
        oracle_segment(j)
    # (classic libjpeg raises JERR_BAD_DCT_COEF here; libjpeg-turbo's encoder does not check and writes an undecodable
    # stream -- either way the GPU encoder must not be the one to code this block: it hands the image back, see
    # tests/test_gpu_huffman.py::test_uncodable_coefficient_falls_back_to_libjpeg)


def test_marker_parser_agrees_with_libjpeg(built):
    """capi.scan_from_jpeg (what a batch host hands to K5): tables, sampling, MCU grid and the start of the entropy-coded
    segment of files libjpeg wrote"""
    for (w, h, subs, gray, quality) in CASES:
        data = util.jpeg_bytes(w, h, subs, quality, seed=w, gray=gray)
        scan, off, frame = capi.scan_from_jpeg(data)
        head, seg, tail = H.split_jpeg(data)
        assert off == len(head) and frame["width"] == w and frame["height"] == h
        j = M.Jpeg()
        assert j.read_jpeg_from_memory(data) == 0
        ref = capi.standard_scan(w, h, j.sampling())
        # Pillow writes the Annex K tables unless asked to optimise -- the ones the components use (a grayscale file has no chroma tables)
        assert (scan.ncomp, scan.mcus_per_row, scan.mcu_rows) == (ref.ncomp, ref.mcus_per_row, ref.mcu_rows)
        for c in range(scan.ncomp):
            assert (scan.h_samp[c], scan.v_samp[c], scan.dc_tbl[c], scan.ac_tbl[c]) == (ref.h_samp[c], ref.v_samp[c], ref.dc_tbl[c], ref.ac_tbl[c])
            assert bytes(scan.dc[scan.dc_tbl[c]]) == bytes(ref.dc[ref.dc_tbl[c]]) and bytes(scan.ac[scan.ac_tbl[c]]) == bytes(ref.ac[ref.ac_tbl[c]])
    with pytest.raises(ValueError):
        import io

        from PIL import Image

        buf = io.BytesIO()
        Image.fromarray(util.photo(64, 64, 1)).save(buf, "JPEG", progressive=True)
        capi.scan_from_jpeg(buf.getvalue())
