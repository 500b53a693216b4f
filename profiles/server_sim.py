#!/usr/bin/env python
"""Request-style load in the manner of the nginx filter the reference's README points to (README.md:406-408; SURVEY
8f rank 3): T host threads, each serving requests "JPEG bytes in -> one shared logo composed on -> JPEG bytes out"
through the unchanged public API (mj_read_jpeg_from_memory, mj_compose, mj_write_jpeg_to_memory), one image per
call -- no batching across requests.  Every thread owns its mj_jpeg_t; the mj_dropon_t is shared and read-only.
The same loop runs against the drop-in library (one mjx_ctx / CUDA stream per thread, created on first use) and
against the unmodified reference (oracle/_ref) on the same host threads.
   small   tests/golden/image.jpg 256x256 4:2:0 + dropon.png 160x50 (BASELINE configs[0])
   photo   1920x1080 4:2:0 q85 + 256x256 logo with a soft disc of alpha, bottom right
usage (GPU box): python profiles/server_sim.py [threads] > gpurun_out/server_sim.json
(set MJX_DROPON_CACHE=1 to let each thread keep its compiled copy of the logo between requests)"""
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import libmodjpeg_b200 as M  # noqa: E402
import util  # noqa: E402
from oracle import oracle_py as O  # noqa: E402  (timed as the baseline only)


def run(threads, seconds, make_request):
    """requests per second over `threads` threads, each looping make_request() for `seconds`"""
    counts = [0] * threads
    stop = threading.Event()
    ready = threading.Barrier(threads + 1)

    def worker(t):
        req = make_request()
        req()  # warm: per-thread context, pools
        ready.wait()
        n = 0
        while not stop.is_set():
            req()
            n += 1
        counts[t] = n

    ts = [threading.Thread(target=worker, args=(t,)) for t in range(threads)]
    for t in ts:
        t.start()
    ready.wait()
    t0 = time.perf_counter()
    time.sleep(seconds)
    stop.set()
    for t in ts:
        t.join()
    return sum(counts) / (time.perf_counter() - t0)


def case(name, data, raw, align, ox, oy, threads, seconds):
    d = M.Dropon()
    assert d.read_dropon_from_raw(raw, M.CS_RGBA, 255) == 0
    ref = O.Reference()
    dr = ref.dropon_from_raw(raw, O.CS_RGBA, 255)

    def ours():
        def req():
            j = M.Jpeg()
            assert j.read_jpeg_from_memory(data) == 0
            assert j.compose(d, align, ox, oy) == 0
            rv, out = j.write_jpeg_to_memory(0)
            assert rv == 0 and len(out) > 100
        return req

    def theirs():
        def req():
            j = ref.read_jpeg(data)
            assert j.compose(dr, align, ox, oy) == 0
            assert len(j.write(0)) > 100
            j.free()
        return req

    def compose_only():
        # the compose call alone on an already decoded image (what the coalescer can change; entropy coding is libjpeg's)
        def req():
            if not hasattr(req, "j"):
                req.j = M.Jpeg()
                assert req.j.read_jpeg_from_memory(data) == 0
            assert req.j.compose(d, align, ox, oy) == 0
        return req

    from libmodjpeg_b200 import capi

    res = {"config": name, "threads": threads, "seconds": seconds, "dropon_cache": os.environ.get("MJX_DROPON_CACHE", "0")}
    for label, n in (("1_thread", 1), ("all_threads", threads)):
        res[label] = {"b200_requests_per_s": run(n, seconds, ours), "reference_requests_per_s": run(n, seconds, theirs)}
    res["all_threads"]["b200_compose_calls_per_s"] = run(threads, seconds, compose_only)
    # the same load with the request coalescer on (mj_coalesce_configure: concurrent calls share one launch and one compiled dropon)
    b0, r0 = capi.coalesce_stats()
    capi.coalesce_configure(True, threads, 150)
    try:
        res["all_threads"]["b200_coalesced_requests_per_s"] = run(threads, seconds, ours)
        res["all_threads"]["b200_coalesced_compose_calls_per_s"] = run(threads, seconds, compose_only)
    finally:
        capi.coalesce_configure(False)
    b1, r1 = capi.coalesce_stats()
    res["all_threads"]["coalescer"] = {"requests": r1 - r0, "launches": b1 - b0, "max_batch": threads, "wait_us": 150}
    return res


def main():
    from PIL import Image

    threads = int(sys.argv[1]) if len(sys.argv) > 1 else len(os.sched_getaffinity(0))
    g = os.path.join(ROOT, "tests", "golden")
    small = open(os.path.join(g, "image.jpg"), "rb").read()
    logo_small = np.array(Image.open(os.path.join(g, "dropon.png")).convert("RGBA"))
    photo = util.jpeg_bytes(1920, 1080, "420", 85, seed=11)
    yy, xx = np.mgrid[0:256, 0:256]
    rgb = np.dstack([(xx % 256), (yy % 256), ((xx + yy) // 2 % 256)]).astype(np.uint8)
    alpha = np.clip((120 - np.hypot(xx - 128, yy - 128)) * 8, 0, 255).astype(np.uint8)
    logo = np.dstack([rgb, alpha])
    res = [case("small: 256x256 + 160x50 logo, top left", small, logo_small, M.ALIGN_TOP | M.ALIGN_LEFT, 0, 0, threads, 3.0),
           case("photo: 1080p + 256x256 soft-disc logo, bottom right", photo, logo, M.ALIGN_BOTTOM | M.ALIGN_RIGHT, -16, -16, threads, 4.0)]
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
