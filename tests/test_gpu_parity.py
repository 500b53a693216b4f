"""GPU: parity of the CUDA path (through the C-ABI) against the oracle.

Bar (north_star): quantised coefficients bit-exact for untouched blocks, opaque-replace blocks,
uniform-alpha blocks, the dropon compile and all integer effects; float-blended (class G) blocks
within +-1 quantisation step, with the differing-coefficient count reported and bounded.
"""
import os

import numpy as np
import pytest

import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "golden.npz"))
G_RATE = 2e-4  # allowed fraction of class-G coefficients off by one step (measured ~1e-6..1e-5)

LAYOUTS = [("420", 3, [(2, 2), (1, 1), (1, 1)]), ("422", 3, [(2, 1), (1, 1), (1, 1)]), ("444", 3, [(1, 1)] * 3),
           ("gray", 1, [(1, 1)]), ("rgb", 2, [(1, 1)] * 3), ("411", 3, [(4, 1), (1, 1), (1, 1)]),
           ("440", 3, [(1, 2), (1, 1), (1, 1)]), ("mixed", 3, [(2, 2), (2, 1), (1, 2)]), ("h3", 3, [(3, 1), (1, 1), (1, 1)])]


def _expected_classes(W):
    ac = (W.reshape(W.shape[0], W.shape[1], 64)[:, :, 1:] != 0).any(-1)
    dc = W[:, :, 0]
    return np.where(ac, 3, np.where(dc == 0, 0, np.where(dc == 2040, 2, 1))).astype(np.uint8)


# ---------------------------------------------------------------------------------------------
# K1: dropon compile
# ---------------------------------------------------------------------------------------------


@pytest.mark.parametrize("name,tcs,samp", LAYOUTS)
def test_k1_compile_bitexact_vs_oracle(engine, port, name, tcs, samp):
    from libmodjpeg_b200 import Layout
    from oracle import oracle_py as O

    for (w, h, boff, crop) in [(48, 32, (0, 0), None), (50, 37, (3, 5), None), (64, 64, (7, 1), (5, 3, 40, 50)),
                               (200, 120, (15, 9), (0, 0, 200, 97))]:
        for cs, nch in [(2, 4), (1, 3), (6, 4), (3, 1), (4, 2)]:
            raw = util.noisy_rgba(w, h, seed=w * 7 + cs)[:, :, :nch]
            if (w, cs) == (200, 2):
                raw = util.logo_rgba(w, h, 64, 27)
            i3, a3, scs, blend = util.ingest_raw(raw, cs, 200)
            rv, D, W = port.compile_dropon(i3, a3, scs, O.make_layout(tcs, samp), boff[0], boff[1], crop)
            L = Layout.make(tcs, samp)
            if rv != 0:
                from libmodjpeg_b200 import MjxError

                with pytest.raises(MjxError) as e:
                    engine.dropon_compile(i3, a3, scs, L, boff, crop)
                assert e.value.code == 6
                continue
            cd = engine.dropon_compile(i3, a3, scs, L, boff, crop)
            assert cd.ncomp == len(samp)
            counts = {"T": 0, "U": 0, "OPAQUE": 0, "G": 0}
            for c in range(len(samp)):
                Dg, Wg, cls = cd.download(c)
                assert np.array_equal(Dg, D[c]), (name, w, cs, c)
                assert np.array_equal(Wg, W[c]), (name, w, cs, c)
                exp = _expected_classes(W[c])
                assert np.array_equal(cls, exp)
                for k, v in zip(("T", "U", "OPAQUE", "G"), np.bincount(exp.reshape(-1), minlength=4)):
                    counts[k] += int(v)
            assert cd.class_counts() == counts
            cd.free()


def test_k1_compile_golden_from_reference(engine):
    """the compile KAT the reference build produced (golden.npz)"""
    from libmodjpeg_b200 import Layout

    i3, a3, cs, blend = util.ingest_raw(G["compile_raw"], 2, 255)
    for name, tcs, samp in LAYOUTS[:7]:
        cd = engine.dropon_compile(i3, a3, cs, Layout.make(tcs, samp), (3, 5), (2, 1, 45, 30))
        for c in range(len(samp)):
            Dg, Wg, _ = cd.download(c)
            assert np.array_equal(Dg, G[f"compile_{name}_D_{c}"]), (name, c)
        cd.free()


# ---------------------------------------------------------------------------------------------
# K2: masked blend on device-resident planes
# ---------------------------------------------------------------------------------------------


def _decode(data):
    from libmodjpeg_b200 import Jpeg

    j = Jpeg()
    assert j.read_jpeg_from_memory(bytes(data)) == 0
    info = j.info()
    return j, info, j.sampling(), j.planes(), [j.qtable(c) for c in range(info["ncomp"])]


def _check_planes(got, want, before, cls_maps, origin, samp, tag):
    """bit-exact outside class G; class G within +-1 step at a bounded rate. Returns (nG, ndiffG)."""
    nG = nbad = 0
    for c in range(len(got)):
        d = got[c].astype(np.int32) - want[c].astype(np.int32)
        assert np.abs(d).max() <= 1, (tag, c, int(np.abs(d).max()))
        gmask = np.zeros(got[c].shape[:2], bool)
        if cls_maps is not None:
            hb, wb = cls_maps[c].shape
            y0, x0 = origin[1] * samp[c][1], origin[0] * samp[c][0]
            gmask[y0:y0 + hb, x0:x0 + wb] = cls_maps[c] == 3
            tmask = np.ones(got[c].shape[:2], bool)
            tmask[y0:y0 + hb, x0:x0 + wb] = cls_maps[c] == 0
            assert np.array_equal(got[c][tmask], before[c][tmask]), (tag, c, "untouched/transparent blocks changed")
        assert not (d[~gmask] != 0).any(), (tag, c, "non-generic block differs")
        nG += int(gmask.sum()) * 64
        nbad += int((d[gmask] != 0).sum())
    assert nbad <= max(3, nG * G_RATE), (tag, nbad, nG)
    return nG, nbad


@pytest.mark.parametrize("subs,gray,quality", [("420", False, 85), ("422", False, 85), ("444", False, 95), ("444", True, 85), ("420", False, 50)])
def test_k2_device_batch_vs_oracle(engine, port, subs, gray, quality):
    from libmodjpeg_b200 import Layout
    from libmodjpeg_b200.batch import DeviceBatch

    W_, H_ = 272, 208
    datas = [util.jpeg_bytes(W_, H_, subs, quality, seed=40 + i, gray=gray) for i in range(3)]
    dec = [_decode(d) for d in datas]
    info, samp = dec[0][1], dec[0][2]
    shapes = [p.shape[:2] for p in dec[0][3]]
    report = []
    for name, raw, cs, blend, align, ox, oy in [("logo", util.logo_rgba(200, 150, 64, 27), 2, 255, 16, 0, 0),
                                                ("noise", util.noisy_rgba(120, 90, 8), 2, 255, 4 | 1, 37, 21),
                                                ("uniform", util.noisy_rgba(120, 90, 9)[:, :, :3], 1, 128, 8 | 2, -5, -3),
                                                ("opaque", util.noisy_rgba(120, 90, 10)[:, :, :3], 1, 255, 4 | 1, -30, -20),
                                                ("fullframe", util.wavy_alpha_rgba(W_, H_), 2, 255, 4 | 1, 0, 0)]:
        if gray and cs == 1 and False:
            continue
        i3, a3, scs, sblend = util.ingest_raw(raw, cs, blend)
        batch = DeviceBatch(engine, shapes, len(dec))
        batch.set_descs(np.stack([np.stack(d[4]) for d in dec]))
        want, before = [], []
        for i, (j, inf, sp, planes, q) in enumerate(dec):
            batch.upload_image(i, planes)
            exp = [p.copy() for p in planes]
            rv, g, D, Wc = util.oracle_compose(port, exp, q, inf["width"], inf["height"], inf["colorspace"], sp, i3, a3, scs,
                                               sblend, align, ox, oy)
            assert rv == 0 and g["visible"]
            want.append(exp)
            before.append(planes)
        cd = engine.dropon_compile(i3, a3, scs, Layout.make(info["colorspace"], samp), (g["blockoffset_x"], g["blockoffset_y"]),
                                   (g["crop_x"], g["crop_y"], g["crop_w"], g["crop_h"]))
        cls_maps = [cd.download(c)[2] for c in range(info["ncomp"])]
        engine.compose_batch_device(batch.descs_dev, batch.n, cd, g["block_x"], g["block_y"])
        engine.sync()
        nG = nbad = 0
        for i in range(len(dec)):
            got = batch.download_image(i)
            a, b = _check_planes(got, want[i], before[i], cls_maps, (g["block_x"], g["block_y"]), samp, (subs, gray, name, i))
            nG += a
            nbad += b
        report.append((name, cd.class_counts(), nG, nbad))
        cd.free()
        batch.free()
    print("\nK2 parity", subs, "gray" if gray else "", f"q{quality}:", report)


def test_k2_device_batch_mixed_quant_tables(engine, port):
    """one launch over images that all carry DIFFERENT quantisation tables (qualities 35..97), more images than a
    warp handles per work item: the per-image table path of both K2 kernels"""
    from libmodjpeg_b200 import Layout
    from libmodjpeg_b200.batch import DeviceBatch

    W_, H_ = 208, 144
    # 130 images > 96 per item: two chunks per tile.  Runs of one, two and three images with the SAME table next to
    # images with different ones: the OPAQUE/U kernel reuses an opaque row only while the table stays the same
    quals = [35 + 2 * i for i in range(32)] * 2 + [35 + 2 * (i // 3) for i in range(48)] + [50, 50, 75, 75, 75, 50] * 3
    base = {q: _decode(util.jpeg_bytes(W_, H_, "420", q, seed=500 + q)) for q in sorted(set(quals))}
    dec = [base[q] for q in quals]
    info, samp = dec[0][1], dec[0][2]
    shapes = [p.shape[:2] for p in dec[0][3]]
    raw = util.logo_rgba(160, 112, 64, 27)
    i3, a3, scs, sblend = util.ingest_raw(raw, 2, 255)
    batch = DeviceBatch(engine, shapes, len(dec))
    batch.set_descs(np.stack([np.stack(d[4]) for d in dec]))
    want = {}
    g = None
    for q, (j, inf, sp, planes, qt) in base.items():
        exp = [p.copy() for p in planes]
        rv, g, D, Wc = util.oracle_compose(port, exp, qt, inf["width"], inf["height"], inf["colorspace"], sp, i3, a3, scs, sblend, 16, 3, -2)
        assert rv == 0 and g["visible"]
        want[q] = exp
    for i, d in enumerate(dec):
        batch.upload_image(i, d[3])
    cd = engine.dropon_compile(i3, a3, scs, Layout.make(info["colorspace"], samp), (g["blockoffset_x"], g["blockoffset_y"]),
                               (g["crop_x"], g["crop_y"], g["crop_w"], g["crop_h"]))
    cls_maps = [cd.download(c)[2] for c in range(info["ncomp"])]
    engine.compose_batch_device(batch.descs_dev, batch.n, cd, g["block_x"], g["block_y"])
    engine.sync()
    nG = nbad = 0
    for i, q in enumerate(quals):
        a, b = _check_planes(batch.download_image(i), want[q], base[q][3], cls_maps, (g["block_x"], g["block_y"]), samp, ("mixedq", q, i))
        nG += a
        nbad += b
    print(f"\nK2 mixed quant tables: {len(quals)} images, {nG} generic coefficients, {nbad} differ by one step")
    # the batch is large enough for the two K2 kernels to run side by side (the default); one after the other must
    # give the same bytes
    side_by_side = [batch.download_image(i) for i in range(len(dec))]
    for i, d in enumerate(dec):
        batch.upload_image(i, d[3])
    engine.set_overlap(False)
    try:
        engine.compose_batch_device(batch.descs_dev, batch.n, cd, g["block_x"], g["block_y"])
        engine.sync()
    finally:
        engine.set_overlap(True)
    for i in range(len(dec)):
        for a, b in zip(side_by_side[i], batch.download_image(i)):
            assert np.array_equal(a, b), ("overlap on/off", i)
    cd.free()
    batch.free()


@pytest.mark.parametrize("n_images", [20, 76])
def test_k2_device_batch_mixed_geometry_and_tables(engine, port, n_images):
    """one launch over images of TWO plane geometries (the smaller image ends inside the dropon, so part of the
    dropon's blocks do not lie on it) and two quantisation tables, interleaved in runs: the kernels skip absent
    blocks per image and may reuse offsets / opaque rows only while geometry / table stay the same.
    20 images take the one-after-the-other launch, 76 the side-by-side one."""
    from libmodjpeg_b200 import Layout, capi
    from libmodjpeg_b200.batch import DeviceBatch

    big = {q: _decode(util.jpeg_bytes(320, 240, "420", q, seed=900 + q)) for q in (60, 90)}
    small = {q: _decode(util.jpeg_bytes(208, 144, "420", q, seed=950 + q)) for q in (60, 90)}
    info, samp = big[60][1], big[60][2]
    shapes_big = [p.shape[:2] for p in big[60][3]]
    shapes_small = [p.shape[:2] for p in small[60][3]]
    raw = util.logo_rgba(288, 208, 64, 27)  # reaches beyond the small image on both axes
    i3, a3, scs, sblend = util.ingest_raw(raw, 2, 255)

    def expect(planes, qt, shapes):
        """oracle on the planes zero-padded to the big geometry (blocks are independent), cropped back"""
        pad = [np.zeros((sb[0], sb[1], 64), np.int16) for sb in shapes_big]
        for a, b in zip(pad, planes):
            a[:b.shape[0], :b.shape[1]] = b
        rv, g, D, Wc = util.oracle_compose(port, pad, qt, info["width"], info["height"], info["colorspace"], samp, i3, a3, scs, sblend, 5, 0, 0)
        assert rv == 0 and g["visible"]
        return [a[:r, :w_].copy() for a, (r, w_) in zip(pad, shapes)], g

    want_big = {q: expect([p.copy() for p in big[q][3]], big[q][4], shapes_big) for q in (60, 90)}
    want_small = {q: expect([p.copy() for p in small[q][3]], small[q][4], shapes_small) for q in (60, 90)}
    g = want_big[60][1]
    # pattern of (geometry, table) in runs of one to three
    pat = [("b", 60), ("b", 60), ("s", 60), ("s", 90), ("s", 90), ("b", 90), ("s", 60), ("b", 60), ("b", 90), ("b", 90), ("b", 90), ("s", 90)]
    seq = [pat[i % len(pat)] for i in range(n_images)]
    nb, ns = sum(1 for k, _ in seq if k == "b"), sum(1 for k, _ in seq if k == "s")
    bb, bs = DeviceBatch(engine, shapes_big, nb), DeviceBatch(engine, shapes_small, ns)
    ptr_rows, ib, is_ = [], 0, 0
    where = []
    for kind, q in seq:
        if kind == "b":
            bb.upload_image(ib, big[q][3])
            d = capi.make_image_descs([[bb.plane_ptr(ib, c) for c in range(3)]], [s_ for _, s_ in shapes_big], [r for r, _ in shapes_big], np.stack(big[q][4]))
            where.append((bb, ib))
            ib += 1
        else:
            bs.upload_image(is_, small[q][3])
            d = capi.make_image_descs([[bs.plane_ptr(is_, c) for c in range(3)]], [s_ for _, s_ in shapes_small], [r for r, _ in shapes_small], np.stack(small[q][4]))
            where.append((bs, is_))
            is_ += 1
        ptr_rows.append(d)
    descs = np.concatenate(ptr_rows)
    descs_dev = engine.device_alloc(descs.nbytes)
    engine.copy_h2d(descs_dev, descs.view(np.uint8).reshape(-1))
    engine.sync()
    cd = engine.dropon_compile(i3, a3, scs, Layout.make(info["colorspace"], samp), (g["blockoffset_x"], g["blockoffset_y"]),
                               (g["crop_x"], g["crop_y"], g["crop_w"], g["crop_h"]))
    cls_maps = [cd.download(c)[2] for c in range(3)]
    engine.compose_batch_device(descs_dev, n_images, cd, g["block_x"], g["block_y"])
    engine.sync()
    nG = nbad = 0
    for i, (kind, q) in enumerate(seq):
        batch, idx = where[i]
        got = batch.download_image(idx)
        if kind == "b":
            a, b = _check_planes(got, want_big[q][0], big[q][3], cls_maps, (g["block_x"], g["block_y"]), samp, ("mixed geometry", i, kind, q))
        else:
            # the class maps reach beyond the small planes: crop them to the plane for the masks
            maps = [m[:got[c].shape[0] - g["block_y"] * samp[c][1], :got[c].shape[1] - g["block_x"] * samp[c][0]] for c, m in enumerate(cls_maps)]
            a, b = _check_planes(got, want_small[q][0], small[q][3], maps, (g["block_x"], g["block_y"]), samp, ("mixed geometry", i, kind, q))
        nG += a
        nbad += b
    print(f"\nK2 mixed geometry ({n_images} images): {nG} generic coefficients, {nbad} differ by one step")
    engine.device_free(descs_dev)
    cd.free()
    bb.free()
    bs.free()


def test_k2_with_oracle_compiled_dropon(engine, port):
    """K2 in isolation: the dropon coefficients come from the oracle (mjx_dropon_from_coefficients)"""
    from libmodjpeg_b200 import Layout
    from oracle import oracle_py as O

    j, info, samp, planes, q = _decode(util.jpeg_bytes(160, 128, "420", 85, seed=5))
    raw = util.logo_rgba(96, 64, 32, 13)
    i3, a3, scs, blend = util.ingest_raw(raw, 2, 255)
    rv, D, Wc = port.compile_dropon(i3, a3, scs, O.make_layout(3, samp))
    cd = engine.dropon_from_coefficients(Layout.make(3, samp), D, Wc)
    want = [p.copy() for p in planes]
    for c in range(3):
        port.compose_plane(want[c], 1 * samp[c][0], 2 * samp[c][1], D[c], Wc[c], q[c])
    got = [p.copy() for p in planes]
    engine.compose_planes_host(got, q, cd, 1, 2)
    cls_maps = [cd.download(c)[2] for c in range(3)]
    _check_planes(got, want, planes, cls_maps, (1, 2), samp, "oracle-compiled")
    cd.free()


def test_batch_host_equals_batch_device(engine):
    from libmodjpeg_b200 import Layout, capi
    from libmodjpeg_b200.batch import DeviceBatch

    dec = [_decode(util.jpeg_bytes(200, 136, "420", 85, seed=60 + i)) for i in range(7)]
    info, samp = dec[0][1], dec[0][2]
    raw = util.logo_rgba(128, 96, 64, 27)
    i3, a3, scs, blend = util.ingest_raw(raw, 2, 255)
    g = capi.geometry(info["width"], info["height"], 16, 16, 128, 96, 16, 5, 3)
    cd = engine.dropon_compile(i3, a3, scs, Layout.make(3, samp), (g["blockoffset_x"], g["blockoffset_y"]),
                               (g["crop_x"], g["crop_y"], g["crop_w"], g["crop_h"]))
    batch = DeviceBatch(engine, [p.shape[:2] for p in dec[0][3]], len(dec))
    batch.set_descs(np.stack([np.stack(d[4]) for d in dec]))
    for i, d in enumerate(dec):
        batch.upload_image(i, d[3])
    engine.compose_batch_device(batch.descs_dev, batch.n, cd, g["block_x"], g["block_y"])
    engine.sync()
    host_planes = [[p.copy() for p in d[3]] for d in dec]
    items = (capi.HostImage * len(dec))()
    keep = []
    for i, d in enumerate(dec):
        it, k = capi.make_host_image(host_planes[i], d[4])
        items[i] = it
        keep.append(k)
    engine.compose_batch_host(items, len(dec), cd, g["block_x"], g["block_y"])
    for i in range(len(dec)):
        for a, b in zip(batch.download_image(i), host_planes[i]):
            assert np.array_equal(a, b)
    cd.free()
    batch.free()


def test_batch_host_zero_copy_equals_staged(engine):
    """page-locked planes take the zero-copy path (K2 works on host memory, one launch for the batch);
    it must give exactly what the staged path gives, and leave every block outside the dropon alone"""
    from libmodjpeg_b200 import Layout, capi

    dec = [_decode(util.jpeg_bytes(200, 136, "420", 85, seed=80 + i)) for i in range(5)]
    info, samp = dec[0][1], dec[0][2]
    raw = util.logo_rgba(128, 96, 64, 27)
    i3, a3, scs, blend = util.ingest_raw(raw, 2, 255)
    g = capi.geometry(info["width"], info["height"], 16, 16, 128, 96, 16, 5, 3)
    cd = engine.dropon_compile(i3, a3, scs, Layout.make(3, samp), (g["blockoffset_x"], g["blockoffset_y"]),
                               (g["crop_x"], g["crop_y"], g["crop_w"], g["crop_h"]))
    staged = [[p.copy() for p in d[3]] for d in dec]
    items = (capi.HostImage * len(dec))()
    keep = []
    for i, d in enumerate(dec):
        items[i], k = capi.make_host_image(staged[i], d[4])
        keep.append(k)
    engine.compose_batch_host(items, len(dec), cd, g["block_x"], g["block_y"])  # pageable numpy memory: staged

    sizes = [p.nbytes for p in dec[0][3]]
    per_image = sum(sizes)
    pinned = engine.host_alloc(per_image * len(dec))
    zitems = (capi.HostImage * len(dec))()
    views = []
    for i, d in enumerate(dec):
        off = i * per_image
        vs = []
        for c, p in enumerate(d[3]):
            v = pinned[off:off + p.nbytes].view(np.int16).reshape(p.shape)
            v[...] = p
            vs.append(v)
            off += p.nbytes
        zitems[i], k = capi.make_host_image(vs, d[4])
        keep.append(k)
        views.append(vs)
    l0 = engine.kernel_launches
    engine.compose_batch_host(zitems, len(dec), cd, g["block_x"], g["block_y"])
    assert engine.kernel_launches - l0 <= 3, "zero-copy path is one K2 launch sequence for the whole batch"
    changed = 0
    for i in range(len(dec)):
        for a, b, before in zip(views[i], staged[i], dec[i][3]):
            assert np.array_equal(a, b)
            changed += int((a != before).sum())
    assert changed > 0
    engine.set_zero_copy(False)
    try:
        for i, d in enumerate(dec):
            for v, p in zip(views[i], d[3]):
                v[...] = p
        l0 = engine.kernel_launches
        engine.compose_batch_host(zitems, len(dec), cd, g["block_x"], g["block_y"])
        assert engine.kernel_launches - l0 >= len(dec)  # staged: per-image launches
        for i in range(len(dec)):
            for a, b in zip(views[i], staged[i]):
                assert np.array_equal(a, b)
    finally:
        engine.set_zero_copy(True)
    engine.host_free(pinned)
    cd.free()


# ---------------------------------------------------------------------------------------------
# the public API end to end: mj_compose / mj_effect_* of libmodjpeg.so
# ---------------------------------------------------------------------------------------------


def test_api_compose_golden_c1(engine):
    """config 1 through the drop-in API: the README fixture"""
    import libmodjpeg_b200 as M

    image = open(os.path.join(HERE, "golden", "image.jpg"), "rb").read()
    j = M.Jpeg()
    assert j.read_jpeg_from_memory(image) == 0
    before = j.planes()
    d = M.Dropon()
    assert d.read_dropon_from_raw(G["c1_dropon_rgba"], M.CS_RGBA, 255) == 0
    assert j.compose(d, M.ALIGN_TOP | M.ALIGN_LEFT, 0, 0) == 0
    got = j.planes()
    nbad = 0
    for c in range(3):
        dd = got[c].astype(np.int32) - G[f"c1_after_{c}"].astype(np.int32)
        assert np.abs(dd).max() <= 1
        nbad += int((dd != 0).sum())
        # the count of changed blocks is the golden one; only a block holding one of the (at most 3) one-step
        # differences may flip between changed and unchanged
        off_blocks = int((dd != 0).any(-1).sum())
        assert abs(int((got[c] != before[c]).any(-1).sum()) - int(G[f"c1_changed_blocks_{c}"])) <= off_blocks
    assert nbad <= 3
    # alpha of dropon.png is {0, 64, 255}: every block away from the glyph edges is T, U or OPAQUE -> mostly exact
    rv, out = j.write_jpeg_to_memory(0)
    assert rv == 0 and out[:2] == b"\xff\xd8"
    if nbad == 0:
        assert np.array_equal(np.frombuffer(out, np.uint8), G["c1_written_jpeg"])


def test_api_compose_geometry_table_golden(engine):
    import libmodjpeg_b200 as M

    data = G["geo_jpeg"].tobytes()
    tot = bad = 0
    for i, (align, ox, oy) in enumerate(G["geo_cases"]):
        j = M.Jpeg()
        assert j.read_jpeg_from_memory(data) == 0
        d = M.Dropon()
        assert d.read_dropon_from_raw(G["geo_dropon"], M.CS_RGBA, 255) == 0
        assert j.compose(d, int(align), int(ox), int(oy)) == 0
        for c, p in enumerate(j.planes()):
            dd = p.astype(np.int32) - G[f"geo_{i}_after_{c}"].astype(np.int32)
            assert np.abs(dd).max() <= 1, (i, c)
            tot += int((G[f"geo_{i}_after_{c}"] != 0).size)
            bad += int((dd != 0).sum())
    assert bad <= 5, (bad, tot)


def test_api_compose_kats_golden(engine):
    import libmodjpeg_b200 as M

    bad = 0
    for entry in G["kat_index"]:
        name, dn, cs, blend = str(entry).split("|")
        j = M.Jpeg()
        assert j.read_jpeg_from_memory(G[f"kat_{name}_jpeg"].tobytes()) == 0
        d = M.Dropon()
        assert d.read_dropon_from_raw(G[f"kat_{name}_{dn}_raw"], int(cs), int(blend)) == 0
        before = j.planes()
        rv = j.compose(d, M.ALIGN_CENTER, 3, -2)
        assert rv == int(G[f"kat_{name}_{dn}_rv"]), (name, dn, rv)
        got = j.planes()
        if rv != 0:
            for a, b in zip(before, got):
                assert np.array_equal(a, b)
            continue
        for c, p in enumerate(got):
            dd = p.astype(np.int32) - G[f"kat_{name}_{dn}_after_{c}"].astype(np.int32)
            assert np.abs(dd).max() <= 1, (name, dn, c)
            if dn in ("rgb_b77", "rgb_b255"):
                assert not dd.any(), (name, dn, c)  # uniform alpha: bit-exact
            bad += int((dd != 0).sum())
    assert bad <= 10, bad


def test_api_compose_vs_reference_live(engine, ref):
    """the product library against the reference library, same inputs, both run here"""
    import libmodjpeg_b200 as M
    from oracle import oracle_py as O

    stats = []
    for subs, gray in [("420", False), ("444", False), ("444", True)]:
        data = util.jpeg_bytes(320, 240, subs, 85, seed=77, gray=gray)
        for raw, cs, blend, align, ox, oy in [(util.logo_rgba(256, 192, 64, 27), 2, 255, 16, 0, 0),
                                              (util.noisy_rgba(90, 70, 3), 2, 255, 10, -8, -4),
                                              (util.noisy_rgba(90, 70, 4)[:, :, :3], 1, 180, 5, -40, 100)]:
            a = ref.read_jpeg(data)
            da = ref.dropon_from_raw(raw, cs, blend)
            assert a.compose(da, align, ox, oy) == 0
            b = M.Jpeg()
            assert b.read_jpeg_from_memory(data) == 0
            db = M.Dropon()
            assert db.read_dropon_from_raw(raw, cs, blend) == 0
            assert b.compose(db, align, ox, oy) == 0
            n = bad = 0
            for pa, pb in zip(a.planes(), b.planes()):
                dd = pa.astype(np.int32) - pb.astype(np.int32)
                assert np.abs(dd).max() <= 1
                n += dd.size
                bad += int((dd != 0).sum())
            if cs == 1:
                assert bad == 0
            stats.append((subs, gray, cs, n, bad))
            assert bad <= max(3, n * G_RATE)
    print("\nAPI vs live reference:", stats)


def test_api_random_sweep_vs_reference_live(engine, ref):
    """80 random compose + effect cases through the drop-in API against the reference library run here
    (profiles/fuzz_parity.py): same return codes, uniform-alpha cases and effects bit-identical, float-blended
    coefficients within one step at a bounded rate"""
    import json
    import subprocess
    import sys

    out = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "fuzz_parity.py"), "80", "7"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    t = json.loads(out.stdout)["totals"]
    print("\nrandom sweep:", t)
    assert t["cases"] == 80 and t["composes_visible"] > 20
    assert t["return_code_mismatch"] == 0 and t["max_abs_diff"] <= 1
    assert t["exact_class_differing"] == 0 and t["effect_differing"] == 0
    assert t["coefficients_differing"] <= G_RATE * t["coefficients_changed"] + 2


def test_api_effects_golden_and_oracle(engine, port):
    import libmodjpeg_b200 as M

    image = open(os.path.join(HERE, "golden", "image.jpg"), "rb").read()
    for name, fn in [("luminance40", lambda j: j.effect_luminance(40)), ("tint30m30", lambda j: j.effect_tint(30, -30)),
                     ("grayscale", lambda j: j.effect_grayscale()), ("pixelate", lambda j: j.effect_pixelate()),
                     ("luminance_wrap", lambda j: j.effect_luminance(2 ** 31 - 1)), ("tint_big", lambda j: j.effect_tint(-5000, 70000))]:
        j = M.Jpeg()
        assert j.read_jpeg_from_memory(image) == 0
        assert fn(j) == 0
        for c, p in enumerate(j.planes()):
            assert np.array_equal(p, G[f"fx_{name}_{c}"]), (name, c)
    # odd sizes (MCU padding blocks must stay untouched), grayscale JPEG gates
    for subs, gray in [("420", False), ("444", True)]:
        data = util.jpeg_bytes(150, 93, subs, 85, seed=6, gray=gray)
        for fx in ("lum", "tint", "gray", "pix"):
            j = M.Jpeg()
            assert j.read_jpeg_from_memory(data) == 0
            info = j.info()
            want = j.planes()
            q = [j.qtable(c) for c in range(info["ncomp"])]
            ycc = info["colorspace"] == 3
            if fx == "lum":
                assert j.effect_luminance(-300) == 0
                if ycc:
                    ci = j.comp_info(0)
                    port.effect_add_dc(want[0], ci["wreal"], ci["hreal"], q[0][0], -300)
            elif fx == "tint":
                assert j.effect_tint(0, 55) == 0
                if ycc:
                    ci = j.comp_info(2)
                    port.effect_add_dc(want[2], ci["wreal"], ci["hreal"], q[2][0], 55)
            elif fx == "gray":
                assert j.effect_grayscale() == 0
                if ycc:
                    for c in (1, 2):
                        ci = j.comp_info(c)
                        port.effect_zero(want[c], ci["wreal"], ci["hreal"])
            else:
                assert j.effect_pixelate() == 0
                for c in range(info["ncomp"]):
                    ci = j.comp_info(c)
                    port.effect_pixelate(want[c], ci["wreal"], ci["hreal"])
            for c, p in enumerate(j.planes()):
                assert np.array_equal(p, want[c]), (subs, gray, fx, c)


def test_k3_device_batch_fused_pipeline(engine, port):
    """several effects in one pass over a device-resident batch == the oracle applying them in order"""
    from libmodjpeg_b200 import capi
    from libmodjpeg_b200.batch import DeviceBatch

    dec = [_decode(util.jpeg_bytes(150, 93, "420", 85, seed=80 + i)) for i in range(4)]
    j0 = dec[0][0]
    real = [(j0.comp_info(c)["wreal"], j0.comp_info(c)["hreal"]) for c in range(3)]
    batch = DeviceBatch(engine, [p.shape[:2] for p in dec[0][3]], len(dec), real_dims=real)
    batch.set_descs(np.stack([np.stack(d[4]) for d in dec]))
    for i, d in enumerate(dec):
        batch.upload_image(i, d[3])
    ops = [(capi.FX_ADD_DC, 0, 40), (capi.FX_PIXELATE, 0, 0), (capi.FX_ADD_DC, 0, -3000), (capi.FX_ADD_DC, 1, 30),
           (capi.FX_ZERO, 2, 0), (capi.FX_ADD_DC, 2, 17)]
    engine.effects_batch_device(batch.descs_dev, batch.n, 3, ops)
    engine.sync()
    for i, d in enumerate(dec):
        want = [p.copy() for p in d[3]]
        q = d[4]
        for op, c, v in ops:
            w, h = real[c]
            if op == capi.FX_ADD_DC:
                port.effect_add_dc(want[c], w, h, q[c][0], v)
            elif op == capi.FX_PIXELATE:
                port.effect_pixelate(want[c], w, h)
            else:
                port.effect_zero(want[c], w, h)
        for c, p in enumerate(batch.download_image(i)):
            assert np.array_equal(p, want[c]), (i, c)
    batch.free()


def test_api_threads_are_independent(engine):
    """concurrent calls on different structs are safe (one engine context per thread)"""
    import threading

    import libmodjpeg_b200 as M

    data = util.jpeg_bytes(160, 128, "420", 85, seed=5)
    raw = util.logo_rgba(96, 64, 32, 13)
    results = [None] * 6

    def work(k):
        j = M.Jpeg()
        d = M.Dropon()
        assert j.read_jpeg_from_memory(data) == 0
        assert d.read_dropon_from_raw(raw, M.CS_RGBA, 255) == 0
        for _ in range(3):
            j2 = M.Jpeg()
            assert j2.read_jpeg_from_memory(data) == 0
            assert j2.compose(d, M.ALIGN_CENTER, k % 2, 0) == 0
            results[k] = [p.copy() for p in j2.planes()]

    ts = [threading.Thread(target=work, args=(k,)) for k in range(6)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for k in range(2, 6):
        for a, b in zip(results[k], results[k % 2]):
            assert np.array_equal(a, b)


def test_compose_batch_over_two_devices(engine):
    """mj_compose_batch with the batch cut into one slice per GPU (one process, a group of host threads and a compiled dropon
    per device): byte-identical outputs to the single-device run"""
    import libmodjpeg_b200 as M
    from libmodjpeg_b200 import capi

    if capi.load_mjx().mjx_device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    datas = [util.jpeg_bytes(320, 240, "420", 85, seed=80 + i) for i in range(5)]
    batch = [datas[i % 5] for i in range(37)]
    d = M.Dropon()
    assert d.read_dropon_from_raw(util.logo_rgba(160, 120, 32, 13), M.CS_RGBA, 255) == 0
    rv1, st1, out1 = capi.compose_batch(batch, d, M.ALIGN_CENTER, 3, 5, 0, nthreads=8)
    capi.batch_set_devices(2)
    try:
        rv2, st2, out2 = capi.compose_batch(batch, d, M.ALIGN_CENTER, 3, 5, 0, nthreads=8)
    finally:
        capi.batch_set_devices(1)
    assert rv1 == 0 and rv2 == 0 and not any(st1) and not any(st2)
    assert all(a == b for a, b in zip(out1, out2))


def test_coalesced_compose_equals_unbatched(engine):
    """request coalescer: 8 threads composing different images with ONE dropon, at two placements and on
    two image sizes (different keys must not share a batch) -- byte-identical to the unbatched calls, fewer launches than
    requests"""
    import threading

    import libmodjpeg_b200 as M
    from libmodjpeg_b200 import capi

    datas = [util.jpeg_bytes(320, 240, "420", 85, seed=60 + i) for i in range(6)] + [util.jpeg_bytes(208, 144, "420", 85, seed=70 + i) for i in range(2)]
    raw = util.logo_rgba(160, 120, 32, 13)
    d = M.Dropon()
    assert d.read_dropon_from_raw(raw, M.CS_RGBA, 255) == 0
    jobs = [(i, datas[i % len(datas)], (M.ALIGN_CENTER, 0, 0) if i % 3 else (M.ALIGN_TOP | M.ALIGN_LEFT, 8, 16)) for i in range(48)]

    def run_all(nthreads):
        out = [None] * len(jobs)

        def work(t):
            for i, data, (al, ox, oy) in jobs[t::nthreads]:
                j = M.Jpeg()
                assert j.read_jpeg_from_memory(data) == 0
                assert j.compose(d, al, ox, oy) == 0
                out[i] = [p.copy() for p in j.planes()]

        ts = [threading.Thread(target=work, args=(t,)) for t in range(nthreads)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        return out

    plain = run_all(8)
    b0, r0 = capi.coalesce_stats()
    capi.coalesce_configure(True, 8, 3000)
    try:
        merged = run_all(8)
    finally:
        capi.coalesce_configure(False)
    b1, r1 = capi.coalesce_stats()
    assert r1 - r0 == len(jobs) and 0 < b1 - b0 < len(jobs), (b1 - b0, r1 - r0)
    for a, b in zip(plain, merged):
        for pa, pb in zip(a, b):
            assert np.array_equal(pa, pb)
    print(f"\ncoalescer: {r1 - r0} requests in {b1 - b0} launches")


def test_k2_strict_mode_reproduces_int16_wraparound(engine, port):
    """adversarial coefficients (|I*q| far outside int16): the strict kernel follows the reference's
    wrap-around bit for bit on T/U/OPAQUE blocks and to +-1 step on G blocks."""
    from libmodjpeg_b200 import Layout
    from oracle import oracle_py as O

    rng = np.random.default_rng(5)
    samp = [(1, 1)] * 3
    planes = [rng.integers(-2047, 2048, (12, 16, 64)).astype(np.int16) for _ in range(3)]
    q = [rng.integers(20, 256, 64).astype(np.uint16) for _ in range(3)]
    try:
        engine.set_strict(True)
        clear = util.noisy_rgba(96, 64, 4)
        clear[:, :, 3] = 0  # alpha 0 everywhere: class T, which the reference still dequantises and requantises in place
        for name, raw, cs, blend in [("uniform", util.noisy_rgba(96, 64, 1)[:, :, :3], 1, 100), ("opaque", util.noisy_rgba(96, 64, 2)[:, :, :3], 1, 255),
                                     ("transparent", clear, 2, 255), ("generic", util.noisy_rgba(96, 64, 3), 2, 255)]:
            i3, a3, scs, sblend = util.ingest_raw(raw, cs, blend)
            rv, D, Wc = port.compile_dropon(i3, a3, scs, O.make_layout(3, samp))
            assert rv == 0
            cd = engine.dropon_compile(i3, a3, scs, Layout.make(3, samp))
            want = [p.copy() for p in planes]
            for c in range(3):
                port.compose_plane(want[c], 2, 1, D[c], Wc[c], q[c])
            got = [p.copy() for p in planes]
            engine.compose_planes_host(got, q, cd, 2, 1)
            n = bad = 0
            for c in range(3):
                dd = got[c].astype(np.int32) - want[c].astype(np.int32)
                if name != "generic":
                    assert not dd.any(), (name, c)
                n += dd.size
                bad += int((dd != 0).sum())
            assert bad <= n * 1e-3, (name, bad, n)
            cd.free()
    finally:
        engine.set_strict(False)


def test_k2_fast_equals_strict_on_encoder_produced_jpegs(engine):
    """the fast kernels (work lists, thread-per-block, fp32-pipe requantisation) and the strict
    kernel agree bit for bit outside class G and to +-1 step inside, on real JPEG coefficients"""
    from libmodjpeg_b200 import Layout, capi

    for subs in ("420", "444"):
        j, info, samp, planes, q = _decode(util.jpeg_bytes(400, 304, subs, 85, seed=91))
        raw = util.logo_rgba(320, 240, 64, 27)
        i3, a3, scs, blend = util.ingest_raw(raw, 2, 255)
        g = capi.geometry(info["width"], info["height"], info["max_h"] * 8, info["max_v"] * 8, 320, 240, 16, 7, 5)
        cd = engine.dropon_compile(i3, a3, scs, Layout.make(3, samp), (g["blockoffset_x"], g["blockoffset_y"]),
                                   (g["crop_x"], g["crop_y"], g["crop_w"], g["crop_h"]))
        fast = [p.copy() for p in planes]
        engine.compose_planes_host(fast, q, cd, g["block_x"], g["block_y"])
        try:
            engine.set_strict(True)
            strict = [p.copy() for p in planes]
            engine.compose_planes_host(strict, q, cd, g["block_x"], g["block_y"])
        finally:
            engine.set_strict(False)
        cls_maps = [cd.download(c)[2] for c in range(3)]
        _check_planes(fast, strict, planes, cls_maps, (g["block_x"], g["block_y"]), samp, ("fast-vs-strict", subs))
        cd.free()


# ---------------------------------------------------------------------------------------------
# the batch pipeline: mj_compose_batch == read + mj_compose + write, image by image
# ---------------------------------------------------------------------------------------------


def test_compose_batch_pipeline_equals_per_image_api(engine):
    import libmodjpeg_b200 as M
    from libmodjpeg_b200 import capi

    raw = util.logo_rgba(200, 120, tile=64, radius=27)
    d = M.Dropon()
    assert d.read_dropon_from_raw(raw, M.CS_RGBA, 255) == 0
    # two geometries and one undecodable input in one batch
    jpegs = [util.jpeg_bytes(320, 240, "420", 85, seed=300 + i) for i in range(5)] + \
            [util.jpeg_bytes(272, 208, "444", 90, seed=310 + i) for i in range(3)] + [b"definitely not a jpeg"]
    order = [0, 5, 1, 8, 6, 2, 3, 7, 4]
    batch = [jpegs[i] for i in order]
    rv, status, outs = capi.compose_batch(batch, d, M.ALIGN_CENTER, 7, -5, 0, nthreads=4)
    assert rv == 0
    for k, i in enumerate(order):
        if i == 8:
            assert status[k] == 5 and outs[k] is None  # MJ_ERR_DECODE_JPEG, the rest of the batch is unaffected
            continue
        assert status[k] == 0
        j = M.Jpeg()
        assert j.read_jpeg_from_memory(jpegs[i]) == 0
        assert j.compose(d, M.ALIGN_CENTER, 7, -5) == 0
        rvw, want = j.write_jpeg_to_memory(0)
        assert rvw == 0
        assert outs[k] == want, (k, i)  # same kernels, same libjpeg: byte-identical files
    # blend == 0: the pipeline is a pure transcode
    d0 = M.Dropon()
    assert d0.read_dropon_from_raw(raw[:, :, :3].copy(), M.CS_RGB, 0) == 0
    rv, status, outs = capi.compose_batch(jpegs[:2], d0, M.ALIGN_CENTER, 0, 0, 0, nthreads=2)
    assert rv == 0 and status == [0, 0]
    for o, src in zip(outs, jpegs[:2]):
        j = M.Jpeg()
        assert j.read_jpeg_from_memory(src) == 0
        assert j.write_jpeg_to_memory(0)[1] == o


def test_reciprocal_tables_divide_exactly_on_device(engine):
    """K2 builds its per-image 1/q tables with MUFU.RCP; trunc(a * rq) must equal a / q for every 16-bit quantiser
    value and every dequantised magnitude the fast path can see -- checked exhaustively by a device kernel"""
    assert engine.selftest_reciprocal() == 0


def test_compiled_dropon_cache_opt_in(engine, tmp_path):
    """MJX_DROPON_CACHE=1: repeated mj_compose calls with one dropon reuse the compiled dropon (fewer kernel
    launches), give identical results, and re-reading the dropon invalidates the entry"""
    import subprocess
    import sys

    code = r'''
import sys, ctypes, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import libmodjpeg_b200 as M, util
from libmodjpeg_b200 import capi
def launches():
    return capi.load_mjx().mjx_ctx_kernel_launches(ctypes.c_void_p(capi.load_modjpeg().mjx_host_ctx()))
raw = util.logo_rgba(96, 64, 32, 13)
d = M.Dropon(); assert d.read_dropon_from_raw(raw, M.CS_RGBA, 255) == 0
outs, per_call = [], []
for i in range(4):
    j = M.Jpeg(); assert j.read_jpeg_from_memory(util.jpeg_bytes(160, 128, "420", 85, seed=700)) == 0
    l0 = launches() if i else 0
    assert j.compose(d, M.ALIGN_CENTER, 0, 0) == 0
    per_call.append(launches() - l0)
    outs.append([p.copy() for p in j.planes()])
assert all(np.array_equal(a, b) for o in outs[1:] for a, b in zip(o, outs[0]))
raw2 = raw.copy(); raw2[:, :, 0] = 255 - raw2[:, :, 0]
assert d.read_dropon_from_raw(raw2, M.CS_RGBA, 255) == 0      # new pixels: the cached entry must not be used
j = M.Jpeg(); assert j.read_jpeg_from_memory(util.jpeg_bytes(160, 128, "420", 85, seed=700)) == 0
assert j.compose(d, M.ALIGN_CENTER, 0, 0) == 0
changed = any(not np.array_equal(a, b) for a, b in zip(j.planes(), outs[0]))
print("RESULT", repr((per_call, changed)))
''' % (ROOT, os.path.join(ROOT, "tests"))
    env = dict(os.environ)
    res = {}
    for on in ("0", "1"):
        env["MJX_DROPON_CACHE"] = on
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")][0]
        res[on] = eval(line[len("RESULT"):])
    (off_calls, off_changed), (on_calls, on_changed) = res["0"], res["1"]
    assert off_changed and on_changed
    assert off_calls[1] == off_calls[2] == off_calls[3]          # reference behaviour: recompile per call
    assert on_calls[1] < off_calls[1] and on_calls[1] == on_calls[3]  # cached: only K2 launches remain
