"""GPU: the tensor-core G kernel of large batches (k2_generic_op.cu) against the oracle and against the fp32 G kernel.

The kernel replaces mj_compose_with_mask / mj_convolve for blocks with a non-uniform mask (reference:
src/compose.c:277-336, src/convolve.c:29-1099) by one tcgen05 UMMA per dropon block and 128 images: the per-block blend
is linear in the image block, its 64 x 64 operator is built once per (dropon, quantisation tables).  Bar: like the fp32
kernel, within +-1 quantisation step of the oracle on class G blocks at a bounded rate, bit-exact elsewhere.  Whatever the
kernel does not serve -- coefficients outside the baseline range [-1024, 1023], images whose tables differ from the first
image's, quantiser values above 255 -- must come out exactly as the fp32 kernel computes it (handed over by the redo mask).
"""
import numpy as np
import pytest

import util
from test_gpu_parity import _check_planes, _decode

pytestmark = pytest.mark.gpu

OP_MIN_IMAGES = 256  # libmodjpeg_b200/csrc/mjx_internal.cuh: kOpMinImages


@pytest.fixture(autouse=True)
def _op_kernel_from_256_images(engine):
    """the kernel serves batches from 256 images on; by default it is only CHOSEN from 1025 on (smaller batches leave its warp
    groups idle and the fp32 kernel is faster) -- the tests want it on their small batches"""
    engine.set_tensor_core_min_images(OP_MIN_IMAGES)
    yield
    engine.set_tensor_core_min_images(1025)


def _run(engine, batch, dec_planes, cd, g, mode, expect_op=None):
    for i, planes in enumerate(dec_planes):
        batch.upload_image(i, planes)
    engine.set_tensor_core(mode)
    try:
        l0 = engine.kernel_launches
        engine.compose_batch_device(batch.descs_dev, batch.n, cd, g["block_x"], g["block_y"])
        engine.sync()
        launches = engine.kernel_launches - l0
    finally:
        engine.set_tensor_core(1)
    if expect_op is not None:
        # fp32 path: at most OPAQUE/U + G kernel; tensor-core path adds prepare, build, the kernel itself and the redo pass
        assert (launches >= 4) == expect_op, f"{launches} launches: the {'tensor-core' if expect_op else 'fp32'} path did not run"
    return [batch.download_image(i) for i in range(batch.n)]


def _compile(engine, port, uniq, raw, align=16, ox=5, oy=3):
    from libmodjpeg_b200 import Layout

    i3, a3, scs, sblend = util.ingest_raw(raw, 2, 255)
    want, g = [], None
    for (j, inf, sp, planes, q) in uniq:
        exp = [p.copy() for p in planes]
        rv, g, D, Wc = util.oracle_compose(port, exp, q, inf["width"], inf["height"], inf["colorspace"], sp, i3, a3, scs, sblend, align, ox, oy)
        assert rv == 0 and g["visible"]
        want.append(exp)
    info, samp = uniq[0][1], uniq[0][2]
    cd = engine.dropon_compile(i3, a3, scs, Layout.make(info["colorspace"], samp), (g["blockoffset_x"], g["blockoffset_y"]),
                               (g["crop_x"], g["crop_y"], g["crop_w"], g["crop_h"]))
    return cd, g, want


@pytest.mark.parametrize("subs,gray,quality,nimg,pieces", [("420", False, 85, 300, 2), ("444", False, 95, 257, 2), ("444", True, 60, 256, 2),
                                                           ("422", False, 100, 384, 2), ("420", False, 100, 520, 2)])
def test_operator_kernel_vs_oracle_and_fp32(built, port, subs, gray, quality, nimg, pieces):
    from libmodjpeg_b200 import Engine
    from libmodjpeg_b200.batch import DeviceBatch

    engine = Engine(0)  # own ctx: the operator cache of a dropon belongs to the first ctx that uses it
    engine.set_tensor_core_min_images(OP_MIN_IMAGES)
    engine.set_operator_pieces(pieces)
    W_, H_ = 208, 144
    uniq = [_decode(util.jpeg_bytes(W_, H_, subs, quality, seed=700 + i, gray=gray)) for i in range(5)]
    dec = [uniq[i % 5] for i in range(nimg)]
    info, samp = dec[0][1], dec[0][2]
    shapes = [p.shape[:2] for p in dec[0][3]]
    for name, raw in [("logo", util.logo_rgba(176, 128, 64, 27)), ("wavy", util.wavy_alpha_rgba(W_, H_))]:
        cd, g, want = _compile(engine, port, uniq, raw)
        cls_maps = [cd.download(c)[2] for c in range(info["ncomp"])]
        batch = DeviceBatch(engine, shapes, nimg)
        batch.set_descs(np.stack([np.stack(d[4]) for d in dec]))
        planes_in = [d[3] for d in dec]
        out = {m: _run(engine, batch, planes_in, cd, g, m, expect_op=m != 0) for m in (0, 2, 1, 1)}
        nG = nbad = ndiff = 0
        for i in range(nimg):
            a, b = _check_planes(out[2][i], want[i % 5], dec[i][3], cls_maps, (g["block_x"], g["block_y"]), samp, ("op", name, i))
            nG += a
            nbad += b
            for c in range(len(shapes)):
                assert np.array_equal(out[1][i][c], out[2][i][c]), "range check on/off must not change in-range results"
                d = out[2][i][c].astype(np.int32) - out[0][i][c].astype(np.int32)
                assert np.abs(d).max() <= 1
                ndiff += int((d != 0).sum())
        print(f"\ntensor-core G kernel {subs} q{quality} {name} ({pieces} pieces): {nG} generic coefficients, {nbad} differ from the oracle "
              f"by one step, {ndiff} from the fp32 kernel")
        cd.free()
        batch.free()
    engine.close()


def test_operator_kernel_out_of_range_blocks_go_to_fp32(engine, port):
    """coefficients outside [-1024, 1023] (possible after mj_effect_luminance on a q=1 table, or in a hostile stream):
    mode 1 must give exactly what the fp32 kernel gives for the affected blocks, and the oracle's result within +-1 step
    everywhere"""
    from libmodjpeg_b200.batch import DeviceBatch

    W_, H_ = 208, 144
    nimg = 288
    uniq = [_decode(util.jpeg_bytes(W_, H_, "420", 90, seed=800 + i)) for i in range(3)]
    info, samp = uniq[0][1], uniq[0][2]
    shapes = [p.shape[:2] for p in uniq[0][3]]
    r = np.random.default_rng(5)
    planes_in, qs = [], []
    for i in range(nimg):
        planes = [p.copy() for p in uniq[i % 3][3]]
        if i % 4 == 1:  # every DC far outside the baseline range (|DC * q| stays inside int16)
            for p in planes:
                p[:, :, 0] = r.integers(1100, 1900, p.shape[:2]) * r.choice([-1, 1], p.shape[:2])
        elif i % 4 == 2:  # a few scattered low-frequency coefficients just outside (small q there: |I*q| stays inside int16)
            for p in planes:
                ys, xs = r.integers(0, p.shape[0], 6), r.integers(0, p.shape[1], 6)
                p[ys, xs, r.choice([0, 1, 8], 6)] = r.choice([-1025, 1024, 1500, -2047], 6)
        planes_in.append(planes)
        qs.append(uniq[i % 3][4])
    raw = util.wavy_alpha_rgba(W_, H_)
    i3, a3, scs, sblend = util.ingest_raw(raw, 2, 255)
    want = []
    for i in range(nimg):
        exp = [p.copy() for p in planes_in[i]]
        rv, g, D, Wc = util.oracle_compose(port, exp, qs[i], info["width"], info["height"], info["colorspace"], samp, i3, a3, scs, sblend, 5, 0, 0)
        assert rv == 0
        want.append(exp)
    cd, g, _ = _compile(engine, port, uniq[:1], raw, 5, 0, 0)
    batch = DeviceBatch(engine, shapes, nimg)
    batch.set_descs(np.stack([np.stack(q) for q in qs]))
    out0 = _run(engine, batch, planes_in, cd, g, 0, expect_op=False)
    out1 = _run(engine, batch, planes_in, cd, g, 1, expect_op=True)
    nbad = n = 0
    for i in range(nimg):
        for c in range(3):
            d = out1[i][c].astype(np.int32) - want[i][c].astype(np.int32)
            # the oracle wraps int16 like the reference; these inputs keep |I*q| and the blend inside int16
            assert np.abs(d).max() <= 1, (i, c, int(np.abs(d).max()))
            n += d.size
            nbad += int((d != 0).sum())
            if i % 4 == 1:
                assert np.array_equal(out1[i][c], out0[i][c]), "every block of this image is out of range: fp32 results expected"
    assert nbad <= max(3, n * 2e-4), (nbad, n)
    cd.free()
    batch.free()


def test_operator_kernel_mixed_tables_and_rebuild(engine, port):
    """images whose quantisation tables differ from the first image's go to the fp32 kernel (bit-identical to mode 0); a
    second batch with other tables rebuilds the cached operator; going back rebuilds it again"""
    from libmodjpeg_b200.batch import DeviceBatch

    W_, H_ = 208, 144
    nimg = 300
    uq = {q: [_decode(util.jpeg_bytes(W_, H_, "420", q, seed=900 + 7 * q + i)) for i in range(3)] for q in (85, 60)}
    info, samp = uq[85][0][1], uq[85][0][2]
    shapes = [p.shape[:2] for p in uq[85][0][3]]
    raw = util.logo_rgba(176, 128, 64, 27)
    cd, g, want85 = _compile(engine, port, uq[85], raw)
    _, _, want60 = _compile(engine, port, uq[60], raw)
    batch = DeviceBatch(engine, shapes, nimg)

    def check(pick):
        dec = [uq[pick(i)][i % 3] for i in range(nimg)]
        want = [(want85 if pick(i) == 85 else want60)[i % 3] for i in range(nimg)]
        batch.set_descs(np.stack([np.stack(d[4]) for d in dec]))
        planes_in = [d[3] for d in dec]
        out0 = _run(engine, batch, planes_in, cd, g, 0, expect_op=False)
        out1 = _run(engine, batch, planes_in, cd, g, 1, expect_op=True)
        first = pick(0)
        nbad = n = 0
        for i in range(nimg):
            for c in range(3):
                d = out1[i][c].astype(np.int32) - want[i][c].astype(np.int32)
                assert np.abs(d).max() <= 1, (i, c)
                n += d.size
                nbad += int((d != 0).sum())
                if pick(i) != first:
                    assert np.array_equal(out1[i][c], out0[i][c]), "other tables than image 0: fp32 results expected"
        assert nbad <= max(3, n * 2e-4), (nbad, n)

    check(lambda i: 60 if i % 5 == 3 else 85)  # mixed: every fifth image has other tables
    check(lambda i: 60)                        # all q60: the operator is rebuilt for the new tables
    check(lambda i: 85 if i else 60)           # image 0 alone decides: everything else is redone in fp32
    check(lambda i: 85)                        # and back
    cd.free()
    batch.free()


def test_operator_kernel_16bit_tables_fall_back(engine, port):
    """quantiser values above 255 (16-bit tables): the component is not served by the tensor-core kernel, every block of it
    goes through the redo mask to the fp32 kernel -- results identical to mode 0"""
    from libmodjpeg_b200.batch import DeviceBatch

    W_, H_ = 208, 144
    nimg = 260
    base = _decode(util.jpeg_bytes(W_, H_, "444", 75, seed=880))
    info, samp = base[1], base[2]
    shapes = [p.shape[:2] for p in base[3]]
    tabs = [np.clip((np.asarray(q, np.int64) * 30), 1, 4000).astype(np.uint16) for q in base[4]]
    tabs[1] = np.asarray(base[4][1], np.uint16)  # one component keeps its 8-bit table: served by the tensor cores
    planes = [np.clip(p.astype(np.int64) // 8, -1024, 1023).astype(np.int16) for p in base[3]]
    planes[1] = base[3][1].copy()
    raw = util.wavy_alpha_rgba(W_, H_)
    i3, a3, scs, sblend = util.ingest_raw(raw, 2, 255)
    exp = [p.copy() for p in planes]
    rv, g, D, Wc = util.oracle_compose(port, exp, tabs, info["width"], info["height"], info["colorspace"], samp, i3, a3, scs, sblend, 5, 0, 0)
    assert rv == 0
    cd, g, _ = _compile(engine, port, [base], raw, 5, 0, 0)
    batch = DeviceBatch(engine, shapes, nimg)
    batch.set_descs(np.stack([np.stack(tabs)] * nimg))
    out0 = _run(engine, batch, [planes] * nimg, cd, g, 0, expect_op=False)
    out1 = _run(engine, batch, [planes] * nimg, cd, g, 1, expect_op=True)
    nbad = n = 0
    for i in range(nimg):
        for c in range(3):
            d = out1[i][c].astype(np.int32) - exp[c].astype(np.int32)
            assert np.abs(d).max() <= 1, (i, c, int(np.abs(d).max()))
            n += d.size
            nbad += int((d != 0).sum())
            if c != 1:
                assert np.array_equal(out1[i][c], out0[i][c])
    assert nbad <= max(3, n * 2e-4), (nbad, n)
    cd.free()
    batch.free()


def test_operator_kernel_partial_overlap_and_small_batches(engine, port):
    """dropon hanging over the right/bottom image edge (absent blocks), image count not a multiple of 128, and a batch
    below the threshold (fp32 path) -- same results from both kernels within +-1 step, oracle parity"""
    from libmodjpeg_b200.batch import DeviceBatch

    W_, H_ = 160, 112
    uniq = [_decode(util.jpeg_bytes(W_, H_, "420", 80, seed=950 + i)) for i in range(4)]
    info, samp = uniq[0][1], uniq[0][2]
    shapes = [p.shape[:2] for p in uniq[0][3]]
    raw = util.wavy_alpha_rgba(120, 96)
    cd, g, want = _compile(engine, port, uniq, raw, 8 | 2, 30, 20)  # BOTTOM | RIGHT, pushed partly off the image
    cls_maps = [cd.download(c)[2] for c in range(3)]
    for nimg in (OP_MIN_IMAGES - 1, 259, 401):
        dec = [uniq[i % 4] for i in range(nimg)]
        batch = DeviceBatch(engine, shapes, nimg)
        batch.set_descs(np.stack([np.stack(d[4]) for d in dec]))
        out = _run(engine, batch, [d[3] for d in dec], cd, g, 1, expect_op=nimg >= OP_MIN_IMAGES)
        nG = nbad = 0
        for i in range(nimg):
            a, b = _check_planes(out[i], want[i % 4], dec[i][3], cls_maps, (g["block_x"], g["block_y"]), samp, ("edge", nimg, i))
            nG += a
            nbad += b
        assert nbad <= max(3, nG * 2e-4), (nbad, nG)
        batch.free()
    cd.free()
