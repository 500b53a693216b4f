"""CPU: the C-ABI libraries load and export every symbol the headers in include/ declare; struct
layouts match the reference's; behaviour without a GPU is a loud error, never a CPU fallback."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header, prefix):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"\w+)\s*\(", txt)))


def test_libmjx_exports_every_declared_symbol(built):
    lib = C.CDLL(built["libmjx"])
    names = _declared("mjx.h", "mjx_")
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n


def test_libmodjpeg_exports_reference_api(built):
    C.CDLL(built["libmjx"], mode=C.RTLD_GLOBAL)
    lib = C.CDLL(built["libmodjpeg"])
    names = _declared("libmodjpeg.h", "mj_")
    # the reference's 16 public functions (reference: src/libmodjpeg.h:129-149) + the additive batch pipeline and coalescer
    additive = ["mj_compose_batch", "mj_batch_set_devices", "mj_coalesce_configure", "mj_coalesce_stats"]
    assert all(a in names for a in additive)
    names = [n for n in names if n not in additive]
    assert names == sorted(["mj_init_dropon", "mj_read_dropon_from_raw", "mj_read_dropon_from_memory", "mj_read_dropon_from_file",
                            "mj_init_jpeg", "mj_read_jpeg_from_memory", "mj_read_jpeg_from_file", "mj_compose",
                            "mj_write_jpeg_to_memory", "mj_write_jpeg_to_file", "mj_free_jpeg", "mj_free_dropon",
                            "mj_effect_grayscale", "mj_effect_pixelate", "mj_effect_tint", "mj_effect_luminance"])
    for n in names + additive + _declared("mjx_host.h", "mjx_"):
        assert hasattr(lib, n), n


def test_struct_layout_matches_libjpeg_abi62(built):
    """the hand-written jpeglib.h must match the runtime: libjpeg itself verifies 632/520 in
    jpeg_Create*; here the derived sizes the reference's callers rely on."""
    from oracle import oracle_py as O

    h = C.CDLL(O.HARNESS_SO)
    assert h.mjh_sizeof_decompress() == 632
    assert h.mjh_sizeof_compress() == 520
    assert h.mjh_sizeof_component() == 96
    assert h.mjh_sizeof_error_mgr() == 168
    assert h.mjh_sizeof_jpeg() == 696
    assert h.mjh_sizeof_dropon() == 32
    assert (h.mjh_offsetof_coef(), h.mjh_offsetof_width(), h.mjh_offsetof_sampling()) == (632, 640, 648)


def test_error_codes_without_touching_the_gpu(built):
    import libmodjpeg_b200 as M

    j = M.Jpeg()
    d = M.Dropon()
    assert j.compose(None, 0) == 2  # MJ_ERR_NULL_DATA
    assert j.effect_pixelate() == 2 and j.effect_grayscale() == 2 and j.effect_tint(1, 1) == 2 and j.effect_luminance(1) == 2
    assert j.read_jpeg_from_memory(b"") == 2
    assert j.read_jpeg_from_memory(b"not a jpeg at all") == 5  # MJ_ERR_DECODE_JPEG
    assert j.read_jpeg_from_file("/nonexistent/file.jpg") == 7  # MJ_ERR_FILEIO
    assert d.read_dropon_from_raw(np.zeros((2, 2, 3), np.uint8), 99, 255) == 4  # MJ_ERR_UNSUPPORTED_COLORSPACE
    assert d.read_dropon_from_memory(b"\x89PNG\r\n\x1a\n0000") == 7  # truncated PNG: MJ_ERR_FILEIO (reference: src/dropon.c:170)
    assert d.read_dropon_from_memory(b"GIF89a0000000000") == 9  # MJ_ERR_UNSUPPORTED_FILETYPE
    assert d.read_dropon_from_memory(b"short") == 2
    rv, out = j.write_jpeg_to_memory(0)
    assert rv == 2


def test_no_cpu_fallback(built):
    """without a device the engine refuses to construct and mj_compose returns MJ_ERR_DEVICE"""
    import libmodjpeg_b200 as M
    from libmodjpeg_b200 import capi
    import util

    if capi.load_mjx().mjx_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(M.MjxError):
        M.Engine(0)
    j = M.Jpeg()
    assert j.read_jpeg_from_memory(util.jpeg_bytes(64, 48, "420", 85, 1)) == 0
    d = M.Dropon()
    assert d.read_dropon_from_raw(util.noisy_rgba(16, 16, 1), M.CS_RGBA, 255) == 0
    before = j.planes()
    assert j.compose(d, M.ALIGN_CENTER) == 10  # MJ_ERR_DEVICE
    assert j.effect_luminance(5) == 10
    for a, b in zip(before, j.planes()):
        assert np.array_equal(a, b)
    # blend 0 and off-image dropons are still no-ops that return MJ_OK (reference: compose.c:38,136)
    d0 = M.Dropon()
    assert d0.read_dropon_from_raw(util.noisy_rgba(16, 16, 1)[:, :, :3], M.CS_RGB, 0) == 0
    assert j.compose(d0, M.ALIGN_CENTER) == 0
    assert j.compose(d, M.ALIGN_TOP | M.ALIGN_LEFT, 64, 0) == 0
