"""BASELINE config 1 as a literal drop-in: the reference's command line tool (src/contrib/modjpeg.c, compiled
UNMODIFIED against the reference's own header by oracle/build_cli.sh) linked once against the reference
library and once against this repo's libmodjpeg.so.  Same program, same arguments, other library."""
import os
import subprocess

import numpy as np
import pytest

import libmodjpeg_b200 as M

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_CLI = os.path.join(ROOT, "oracle", "_ref", "modjpeg_ref")
B200_CLI = os.path.join(ROOT, "oracle", "_ref", "modjpeg_b200")
IMAGE = os.path.join(ROOT, "tests", "golden", "image.jpg")
DROPON = os.path.join(ROOT, "tests", "golden", "dropon.png")

needs_cli = pytest.mark.skipif(not (os.path.exists(REF_CLI) and os.path.exists(B200_CLI)),
                               reason="oracle/_ref CLI binaries not built (need /root/reference once)")


def _run(cli, *args):
    return subprocess.run([cli, *args], capture_output=True, text=True, timeout=120)


def _planes(path):
    j = M.Jpeg()
    assert j.read_jpeg_from_file(path) == 0
    return [p.copy() for p in j.planes()]


@needs_cli
def test_cli_links_against_the_dropin_and_host_path_is_byte_identical(built, tmp_path):
    # no compose / effect requested: pure host libjpeg path, must reproduce the reference's bytes
    a, b = str(tmp_path / "ref.jpg"), str(tmp_path / "b200.jpg")
    for opts in ([], ["-O"], ["-P"]):
        assert _run(REF_CLI, "-i", IMAGE, *opts, "-o", a).returncode == 0
        assert _run(B200_CLI, "-i", IMAGE, *opts, "-o", b).returncode == 0
        assert open(a, "rb").read() == open(b, "rb").read(), opts


@needs_cli
def test_cli_fails_loudly_without_a_gpu(built, tmp_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = _run(B200_CLI, "-i", IMAGE, "-d", DROPON, "-o", str(tmp_path / "x.jpg"))
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


def test_png_dropon_ingest_matches_pillow(built):
    from PIL import Image

    d = M.Dropon()
    assert d.read_dropon_from_file(DROPON) == 0  # libpng simplified API (reference: src/dropon.c:163-201)
    rgba = np.array(Image.open(DROPON).convert("RGBA"))
    assert (d.width, d.height, d.colorspace, d.blend) == (160, 50, M.CS_RGB, -1)
    assert np.array_equal(d.image3(), rgba[:, :, :3]) and np.array_equal(d.alpha3()[:, :, 0], rgba[:, :, 3])
    assert d.read_dropon_from_memory(b"GIF89a__________") == 9  # MJ_ERR_UNSUPPORTED_FILETYPE


@needs_cli
@pytest.mark.gpu
@pytest.mark.parametrize("args", [
    ["-d", DROPON],                                              # README example: top left
    ["-p", "br", "-m", "-5,-7", "-d", DROPON],                   # bottom right with offset
    ["-p", "c", "-d", DROPON, "-y", "25", "-b", "-12", "-r", "9"],  # compose, then luminance and tint
    ["-d", DROPON, "-x"],                                        # compose, then pixelate
    ["-g", "-p", "tr", "-d", DROPON, "-O"],                      # grayscale first, optimised output
])
def test_cli_config1_same_program_other_library(engine, tmp_path, args):
    a, b = str(tmp_path / "ref.jpg"), str(tmp_path / "b200.jpg")
    ra = _run(REF_CLI, "-i", IMAGE, *args, "-o", a)
    rb = _run(B200_CLI, "-i", IMAGE, *args, "-o", b)
    assert ra.returncode == 0, ra.stderr
    assert rb.returncode == 0, rb.stderr
    before = _planes(IMAGE)
    n = bad = changed = 0
    for pa, pb, p0 in zip(_planes(a), _planes(b), before):
        d = pa.astype(np.int32) - pb.astype(np.int32)
        assert np.abs(d).max() <= 1, args  # float-blended blocks: within one quantisation step
        n += d.size
        bad += int((d != 0).sum())
        changed += int((pa != p0).sum())
    assert changed > 0
    assert bad <= max(3, int(n * 2e-4)), (args, bad, n)
