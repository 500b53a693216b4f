#!/usr/bin/env python
"""K2 on the other BASELINE.json configurations (parity cases in tests/, timed here for the record):
   c2    6000x4000 4:2:0 + 1024^2 radial-alpha watermark, centred
   c4i   7680x4320 4:4:4, full-frame overlay, uniform alpha 128   (every block class U)
   c4ii  7680x4320 4:4:4, full-frame overlay, non-uniform alpha   (every block class G: worst case)
   c4g   7680x4320 grayscale, non-uniform alpha
Each image is replicated in HBM until the batch exceeds the 126 MB L2; device time by CUDA events.
The *_n1 lines are the configurations LITERALLY (one image, n = 1): L2 is flushed before every timed launch (a 512 MB buffer
is overwritten outside the event pair), the compiled dropon is part of the traffic.
usage (GPU box): python profiles/configs_perf.py > gpurun_out/configs_perf.json"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import libmodjpeg_b200 as M  # noqa: E402
import util  # noqa: E402
from libmodjpeg_b200 import capi  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


_FLUSH = None


def run(name, W, H, subs, gray, raw, cs, blend, align, copies, engine, stream, dev, flush=False):
    global _FLUSH
    j = M.Jpeg()
    assert j.read_jpeg_from_memory(util.jpeg_bytes(W, H, subs, 85, seed=11, gray=gray)) == 0
    info, samp = j.info(), j.sampling()
    planes = j.planes()
    nc = info["ncomp"]
    q = np.stack([j.qtable(c) for c in range(nc)])
    shapes = [p.shape[:2] for p in planes]
    flat = np.concatenate([p.reshape(-1).view(np.uint8) for p in planes])
    slab = torch.empty((copies, flat.size), dtype=torch.uint8, device=dev)
    slab[:] = torch.from_numpy(flat).to(dev)
    off = np.concatenate([[0], np.cumsum([p.nbytes for p in planes])[:-1]])
    ptrs = [[slab.data_ptr() + i * flat.size + int(off[c]) for c in range(nc)] for i in range(copies)]
    descs = capi.make_image_descs(ptrs, [s for _, s in shapes], [r for r, _ in shapes], q)
    descs_dev = torch.from_numpy(descs.view(np.uint8).reshape(-1).copy()).to(dev)
    i3, a3, scs, sblend = util.ingest_raw(raw, cs, blend)
    g = capi.geometry(info["width"], info["height"], info["max_h"] * 8, info["max_v"] * 8, raw.shape[1], raw.shape[0], align, 0, 0)
    cd = engine.dropon_compile(i3, a3, scs, M.Layout.make(info["colorspace"], samp), (g["blockoffset_x"], g["blockoffset_y"]),
                               (g["crop_x"], g["crop_y"], g["crop_w"], g["crop_h"]))
    cnt = cd.class_counts()

    def step():
        engine.compose_batch_device(descs_dev.data_ptr(), copies, cd, g["block_x"], g["block_y"])

    for _ in range(3):
        step()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    if flush:
        if _FLUSH is None:
            _FLUSH = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
        ms = 0.0
        for _ in range(reps):
            _FLUSH.add_(1)  # 512 MB read + written: nothing of the image or the dropon is left in L2
            e0.record(stream)
            step()
            e1.record(stream)
            torch.cuda.synchronize(dev)
            ms += e0.elapsed_time(e1) / reps
    else:
        e0.record(stream)
        for _ in range(reps):
            step()
        e1.record(stream)
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / reps
    alg = copies * (cnt["OPAQUE"] * 128 + (cnt["U"] + cnt["G"]) * 256) + cd.blocks * 4 + (cnt["OPAQUE"] + cnt["U"]) * 128 + cnt["G"] * 256
    out = {"config": name, "image": f"{W}x{H} {'gray' if gray else subs}", "copies_in_hbm": copies, "resident_bytes": int(copies * flat.size),
           "dropon_blocks": cd.blocks, "classes": cnt, "ms": ms, "mblocks_per_s": copies * cd.blocks / ms / 1e3,
           "algorithmic_bytes": alg, "achieved_gbs": alg / ms / 1e6, "frac_of_measured_copy_peak": alg / ms / 1e6 / PEAK,
           "l2": "flushed before every timed launch" if flush else "batch larger than L2"}
    if flush:
        # what the launch really moves for ONE image: the G kernel reads the compiled dropon as fp32 (A and Ds, 512 B per G block)
        moved = alg + cnt["G"] * 256
        out["moved_bytes"] = moved
        out["moved_gbs"] = moved / ms / 1e6
        out["moved_frac_of_measured_copy_peak"] = moved / ms / 1e6 / PEAK
    cd.free()
    del slab
    torch.cuda.empty_cache()
    return out


def main():
    dev = torch.device("cuda", 0)
    engine = M.Engine(0)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    engine.set_stream(stream.cuda_stream)
    res = []
    yy, xx = np.mgrid[0:1024, 0:1024]
    r = np.hypot(yy - 511.5, xx - 511.5)
    wm = np.zeros((1024, 1024, 4), np.uint8)
    wm[:, :, 0], wm[:, :, 1], wm[:, :, 2] = xx // 4, yy // 4, 128
    wm[:, :, 3] = np.clip((480 - r) / 96 * 255, 0, 255).astype(np.uint8)
    res.append(run("c2", 6000, 4000, "420", False, wm, 2, 255, 16, 256, engine, stream, dev))
    res.append(run("c2_n1", 6000, 4000, "420", False, wm, 2, 255, 16, 1, engine, stream, dev, flush=True))
    rng = np.random.default_rng(3)
    small = rng.integers(0, 256, size=(4320 // 8, 7680 // 8, 3), dtype=np.uint8)
    rgb = np.ascontiguousarray(np.repeat(np.repeat(small, 8, 0), 8, 1))
    res.append(run("c4i", 7680, 4320, "444", False, rgb, 1, 128, 4 | 1, 8, engine, stream, dev))
    wavy = util.wavy_alpha_rgba(7680, 4320)
    res.append(run("c4ii", 7680, 4320, "444", False, wavy, 2, 255, 4 | 1, 8, engine, stream, dev))
    res.append(run("c4i_n1", 7680, 4320, "444", False, rgb, 1, 128, 4 | 1, 1, engine, stream, dev, flush=True))
    res.append(run("c4ii_n1", 7680, 4320, "444", False, wavy, 2, 255, 4 | 1, 1, engine, stream, dev, flush=True))
    res.append(run("c4g", 7680, 4320, "444", True, wavy, 2, 255, 4 | 1, 16, engine, stream, dev))
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
