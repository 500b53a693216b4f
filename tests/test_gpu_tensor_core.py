"""GPU: the tensor-core G kernel (k2_generic_tc.cu) against the oracle and against the fp32 G kernel.

The kernel replaces the inverse transform of mj_compose_with_mask / mj_convolve (reference: src/compose.c:277-336,
src/convolve.c:29-1099) by a tcgen05 UMMA on exact fp16 integers.  Bar: like the fp32 kernel, within +-1 quantisation
step of the oracle on class G blocks at a bounded rate, bit-exact elsewhere; blocks whose coefficients leave the baseline
range [-1024, 1023] must come out exactly as the fp32 kernel computes them (mode 1 hands them over).
"""
import numpy as np
import pytest

import util
from test_gpu_parity import _check_planes, _decode

pytestmark = pytest.mark.gpu


def _run(engine, batch, dec_planes, cd, g, mode):
    for i, planes in enumerate(dec_planes):
        batch.upload_image(i, planes)
    engine.set_tensor_core(mode)
    try:
        engine.compose_batch_device(batch.descs_dev, batch.n, cd, g["block_x"], g["block_y"])
        engine.sync()
    finally:
        engine.set_tensor_core(1)
    return [batch.download_image(i) for i in range(batch.n)]


@pytest.mark.parametrize("subs,gray,quality,nimg", [("420", False, 85, 40), ("444", False, 95, 30), ("444", True, 60, 26), ("422", False, 100, 50)])
def test_tensor_core_kernel_vs_oracle_and_fp32(engine, port, subs, gray, quality, nimg):
    from libmodjpeg_b200 import Layout
    from libmodjpeg_b200.batch import DeviceBatch

    W_, H_ = 208, 144
    uniq = [_decode(util.jpeg_bytes(W_, H_, subs, quality, seed=700 + i, gray=gray)) for i in range(5)]
    dec = [uniq[i % 5] for i in range(nimg)]
    info, samp = dec[0][1], dec[0][2]
    shapes = [p.shape[:2] for p in dec[0][3]]
    for name, raw in [("logo", util.logo_rgba(176, 128, 64, 27)), ("wavy", util.wavy_alpha_rgba(W_, H_))]:
        i3, a3, scs, sblend = util.ingest_raw(raw, 2, 255)
        want = []
        g = None
        for (j, inf, sp, planes, q) in uniq:
            exp = [p.copy() for p in planes]
            rv, g, D, Wc = util.oracle_compose(port, exp, q, inf["width"], inf["height"], inf["colorspace"], sp, i3, a3, scs, sblend, 16, 5, 3)
            assert rv == 0 and g["visible"]
            want.append(exp)
        cd = engine.dropon_compile(i3, a3, scs, Layout.make(info["colorspace"], samp), (g["blockoffset_x"], g["blockoffset_y"]),
                                   (g["crop_x"], g["crop_y"], g["crop_w"], g["crop_h"]))
        cls_maps = [cd.download(c)[2] for c in range(info["ncomp"])]
        batch = DeviceBatch(engine, shapes, nimg)
        batch.set_descs(np.stack([np.stack(d[4]) for d in dec]))
        planes_in = [d[3] for d in dec]
        out = {m: _run(engine, batch, planes_in, cd, g, m) for m in (0, 2, 1)}
        nG = nbad = ndiff = 0
        for i in range(nimg):
            a, b = _check_planes(out[2][i], want[i % 5], dec[i][3], cls_maps, (g["block_x"], g["block_y"]), samp, ("tc", name, i))
            nG += a
            nbad += b
            for c in range(len(shapes)):
                assert np.array_equal(out[1][i][c], out[2][i][c]), "range check on/off must not change in-range results"
                d = out[2][i][c].astype(np.int32) - out[0][i][c].astype(np.int32)
                assert np.abs(d).max() <= 1
                ndiff += int((d != 0).sum())
        print(f"\ntensor-core G kernel {subs} q{quality} {name}: {nG} generic coefficients, {nbad} differ from the oracle by one step, "
              f"{ndiff} from the fp32 kernel")
        cd.free()
        batch.free()


def test_tensor_core_kernel_out_of_range_blocks_go_to_fp32(engine, port):
    """coefficients outside [-1024, 1023] (possible after mj_effect_luminance on a q=1 table, or in a hostile stream):
    mode 1 must give exactly what the fp32 kernel gives for the affected (tile, image) pairs, and the oracle's result
    within +-1 step everywhere"""
    from libmodjpeg_b200 import Layout
    from libmodjpeg_b200.batch import DeviceBatch

    W_, H_ = 208, 144
    nimg = 36
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
    want, g = [], None
    for i in range(nimg):
        exp = [p.copy() for p in planes_in[i]]
        rv, g, D, Wc = util.oracle_compose(port, exp, qs[i], info["width"], info["height"], info["colorspace"], samp, i3, a3, scs, sblend, 5, 0, 0)
        assert rv == 0
        want.append(exp)
    cd = engine.dropon_compile(i3, a3, scs, Layout.make(info["colorspace"], samp), (g["blockoffset_x"], g["blockoffset_y"]),
                               (g["crop_x"], g["crop_y"], g["crop_w"], g["crop_h"]))
    batch = DeviceBatch(engine, shapes, nimg)
    batch.set_descs(np.stack([np.stack(q) for q in qs]))
    out0 = _run(engine, batch, planes_in, cd, g, 0)
    out1 = _run(engine, batch, planes_in, cd, g, 1)
    nbad = n = 0
    for i in range(nimg):
        for c in range(3):
            d = out1[i][c].astype(np.int32) - want[i][c].astype(np.int32)
            # the oracle wraps int16 like the reference; these inputs keep |I*q| and the blend inside int16
            assert np.abs(d).max() <= 1, (i, c, int(np.abs(d).max()))
            n += d.size
            nbad += int((d != 0).sum())
            if i % 4 == 1:
                assert np.array_equal(out1[i][c], out0[i][c]), "all tiles of this image are out of range: fp32 results expected"
    assert nbad <= max(3, n * 2e-4), (nbad, n)
    cd.free()
    batch.free()


def test_tensor_core_kernel_16bit_tables(engine, port):
    """quantiser values up to 4000: B is rescaled per table (fp16 range) and q is split into two pieces"""
    from libmodjpeg_b200 import Layout
    from libmodjpeg_b200.batch import DeviceBatch

    W_, H_ = 208, 144
    nimg = 30
    base = _decode(util.jpeg_bytes(W_, H_, "444", 75, seed=880))
    info, samp = base[1], base[2]
    shapes = [p.shape[:2] for p in base[3]]
    r = np.random.default_rng(9)
    tabs = []
    for k in range(3):
        t = [np.clip((np.asarray(q, np.int64) * [1, 30, 200][k]), 1, 4000).astype(np.uint16) for q in base[4]]
        tabs.append(t)
    planes_in, qs = [], []
    for i in range(nimg):
        k = i % 3
        planes = [np.clip(p.astype(np.int64) // [1, 8, 40][k], -1024, 1023).astype(np.int16) for p in base[3]]
        planes_in.append(planes)
        qs.append(tabs[k])
    raw = util.wavy_alpha_rgba(W_, H_)
    i3, a3, scs, sblend = util.ingest_raw(raw, 2, 255)
    want, g = [], None
    for i in range(3):
        exp = [p.copy() for p in planes_in[i]]
        rv, g, D, Wc = util.oracle_compose(port, exp, qs[i], info["width"], info["height"], info["colorspace"], samp, i3, a3, scs, sblend, 5, 0, 0)
        assert rv == 0
        want.append(exp)
    cd = engine.dropon_compile(i3, a3, scs, Layout.make(info["colorspace"], samp), (g["blockoffset_x"], g["blockoffset_y"]),
                               (g["crop_x"], g["crop_y"], g["crop_w"], g["crop_h"]))
    batch = DeviceBatch(engine, shapes, nimg)
    batch.set_descs(np.stack([np.stack(q) for q in qs]))
    out = _run(engine, batch, planes_in, cd, g, 1)
    nbad = n = 0
    for i in range(nimg):
        for c in range(3):
            d = out[i][c].astype(np.int32) - want[i % 3][c].astype(np.int32)
            assert np.abs(d).max() <= 1, (i, c, int(np.abs(d).max()))
            n += d.size
            nbad += int((d != 0).sum())
    assert nbad <= max(3, n * 2e-4), (nbad, n)
    cd.free()
    batch.free()
