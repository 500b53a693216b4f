"""CPU: the host boundary's libjpeg side (no GPU involved): reading, writing and dropon ingest
produce what the reference build produced (golden vectors)."""
import os

import numpy as np

import libmodjpeg_b200 as M
import util

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "golden.npz"))


def test_read_write_bytes_identical_to_reference(built):
    image = open(os.path.join(HERE, "golden", "image.jpg"), "rb").read()
    for opt in (0, 1, 2, 3):
        j = M.Jpeg()
        assert j.read_jpeg_from_memory(image) == 0
        rv, out = j.write_jpeg_to_memory(opt)
        assert rv == 0
        assert np.array_equal(np.frombuffer(out, np.uint8), G[f"rw_opt{opt}"]), opt


def test_write_after_import_roundtrip(built, tmp_path):
    data = util.jpeg_bytes(120, 72, "422", 90, seed=3)
    j = M.Jpeg()
    assert j.read_jpeg_from_memory(data) == 0
    planes = j.planes()
    planes[0][:, :, 5] += 1
    j.set_plane(0, planes[0])
    path = str(tmp_path / "o.jpg")
    assert j.write_jpeg_to_file(path, M.capi.OPTION_OPTIMIZE) == 0
    k = M.Jpeg()
    assert k.read_jpeg_from_file(path) == 0
    for c, (a, b) in enumerate(zip(planes, k.planes())):
        ci = k.comp_info(c)  # blocks beyond the real dims are MCU padding the encoder regenerates
        assert np.array_equal(a[:ci["hreal"], :ci["wreal"]], b[:ci["hreal"], :ci["wreal"]])
    assert k.info()["width"] == 120 and k.comp_info(1)["h"] == 1 and k.comp_info(0)["h"] == 2


def test_max_pixel_and_colorspace_gates(built):
    import io
    from PIL import Image

    data = util.jpeg_bytes(64, 64, "420", 85, 1)
    j = M.Jpeg()
    assert j.read_jpeg_from_memory(data, max_pixel=100) == 8  # MJ_ERR_IMAGE_SIZE
    assert j.read_jpeg_from_memory(data, max_pixel=64 * 64) == 0
    b = io.BytesIO()
    Image.fromarray(util.photo(32, 32, 1)).convert("CMYK").save(b, "JPEG")
    assert j.read_jpeg_from_memory(b.getvalue()) == 4  # MJ_ERR_UNSUPPORTED_COLORSPACE (reference: src/image.c:84-92)


def test_dropon_ingest_matches_reference(built):
    for dn in ("rgba", "rgb", "ycc", "ycca", "gray", "graya"):
        cs, blend = G[f"ingest_{dn}_args"]
        d = M.Dropon()
        assert d.read_dropon_from_raw(G[f"ingest_{dn}_raw"], int(cs), int(blend)) == 0
        assert np.array_equal(d.image3(), G[f"ingest_{dn}_image3"]), dn
        assert np.array_equal(d.alpha3(), G[f"ingest_{dn}_alpha3"]), dn
        assert [d.width, d.height, d.colorspace, d.blend] == list(G[f"ingest_{dn}_meta"]), dn


def test_dropon_from_jpeg_with_mask(built):
    import io
    from PIL import Image

    rgb = util.photo(40, 24, 5)
    b = io.BytesIO()
    Image.fromarray(rgb).save(b, "JPEG", quality=95)
    m = io.BytesIO()
    Image.fromarray(rgb[:, :, 0]).save(m, "JPEG", quality=95)
    d = M.Dropon()
    assert d.read_dropon_from_memory(b.getvalue(), m.getvalue(), 255) == 0
    assert (d.width, d.height, d.colorspace, d.blend) == (40, 24, M.CS_RGB, -1)
    a3 = d.alpha3()
    assert (a3[:, :, 0] == a3[:, :, 1]).all() and a3.std() > 1
    assert d.read_dropon_from_memory(b.getvalue(), None, 128) == 0
    assert d.blend == 128 and (d.alpha3() == 128).all()
    big = io.BytesIO()
    Image.fromarray(util.photo(48, 24, 5)[:, :, 0]).save(big, "JPEG")
    assert d.read_dropon_from_memory(b.getvalue(), big.getvalue(), 255) == 3  # MJ_ERR_DROPON_DIMENSIONS


def test_geometry_matches_oracle(built, port):
    rng = np.random.default_rng(0)
    for _ in range(3000):
        W, H = int(rng.integers(1, 400)), int(rng.integers(1, 400))
        dw, dh = int(rng.integers(1, 300)), int(rng.integers(1, 300))
        hf, vf = int(rng.choice([8, 16, 24, 32])), int(rng.choice([8, 16, 32]))
        align = int(rng.integers(0, 32))
        ox, oy = int(rng.integers(-450, 450)), int(rng.integers(-450, 450))
        a = M.geometry(W, H, hf, vf, dw, dh, align, ox, oy)
        b = port.geometry(W, H, hf, vf, dw, dh, align, ox, oy)
        if not b["visible"]:
            assert not a["visible"]
        else:
            assert a == b, (W, H, dw, dh, hf, vf, align, ox, oy)
