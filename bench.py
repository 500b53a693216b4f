#!/usr/bin/env python
"""bench.py -- headline benchmark of the DCT-domain compositing hot path (BASELINE.json).

Workload (config.workload "c3"): BASELINE.json configs[2] -- a batch of synthetic 1920x1080
4:2:0 q85 JPEGs with a full-frame tiled alpha logo, sharded by image over the GPUs (weak
scaling: 1250 images per GPU, so 8 GPUs hold the 10 000-image batch).  It is the configuration
the metric ("composited Mblocks/s + images/s at 1/2/4/8 B200") is quoted on; configs[1] (one
24 MP image + 1024^2 watermark = 25 k blocks, ~3 MB) is a single microsecond-scale launch and is
covered as a parity case instead.

One "step" = one pass of K2 (mjx_compose_batch_device) over the rank's whole batch, planes
resident in HBM (7.8 GB per GPU >> 126 MB L2, so no L2 flush is needed between steps).
`value`  = composited blocks / s over all ranks, device-timed with CUDA events, max over ranks.
`e2e`    = the same metric through the C-ABI call that takes HOST planes
           (mjx_compose_batch_host): pinned host memory -> H2D -> K2 -> D2H inside the timed region.
`roofline` = algorithmic HBM bytes of one launch / its CUDA-event duration vs the
           measured copy bandwidth in MEASURED_PEAKS.json.
`cpu_baseline` = the unmodified reference (oracle/_ref, mj_compose incl. its dropon compile) on
           the box's host cores, on a bounded sample of the same images.
`roofline.alternatives` = the G class on the range-vouched tensor-core kernel and on the fp32 kernel, same run.
`parity`  = 8 images of the timed batch against oracle/_ref (outside every timed region).
`other_kernels` = K1, K3, K4 (Huffman coding) and K5 (Huffman decoding) on the same resident batch.
`e2e_files` = mj_compose_batch: JPEG bytes in -> JPEG bytes out (K5, K1, K2, K4 on the device), and the unmodified
           reference's read -> compose -> write loop on the same host threads.
`--impl reference` times only that CPU arm and prints the same line shape.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

W_IMG, H_IMG = 1920, 1080
N_BASES = 16
ALIGN_TOP_LEFT = 4 | 1
# both arms name the workload with the same words
WORKLOAD = "c3: batch of 1920x1080 4:2:0 q85 JPEGs + full-frame tiled alpha logo, 48960 blocks/image (BASELINE.json configs[2])"


def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# stdout carries exactly ONE line, the JSON result: file descriptor 1 is pointed at stderr for the whole run (NCCL prints
# its version banner with printf, libjpeg and the reference print warnings) and the result goes to the saved descriptor.
_RESULT_OUT = None


def capture_stdout() -> None:
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _RESULT_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


# ------------------------------------------------------------------------------------------
# synthetic inputs: 16 distinct 1080p 4:2:0 q85 bases cycled over the batch
# ------------------------------------------------------------------------------------------


def make_inputs(n_bases: int):
    import util

    jpegs = [util.jpeg_bytes(W_IMG, H_IMG, "420", 85, seed=100 + i) for i in range(n_bases)]
    logo = util.logo_rgba(W_IMG, H_IMG, tile=256, radius=110)
    return jpegs, logo


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [l for t, l in self.lines if t0 - 0.05 <= t <= t1 + 0.15] or [l for _, l in self.lines]
        sm, mx, reasons = [], [], set()
        for l in rows:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# the reference's CPU path (oracle/_ref): used by cpu_baseline and by --impl reference
# ------------------------------------------------------------------------------------------


class CpuArm:
    """mj_compose of the unmodified reference build on host threads.  Each thread owns decoded
    copies of the base images and one dropon; a 'step' composes one image per thread."""

    def __init__(self, jpegs, logo, threads: int):
        from oracle import oracle_py as O

        self.O = O
        if O.have_reference():
            self.lib = O.Reference()
            self.kind = "reference"
        else:
            raise RuntimeError("oracle/_ref/libmodjpeg_ref.so is missing (build it with oracle/build_ref.sh where /root/reference exists)")
        self.threads = threads
        self.jpegs = jpegs
        self.state = []
        for t in range(threads):
            j = self.lib.read_jpeg(jpegs[t % len(jpegs)])
            d = self.lib.dropon_from_raw(logo, O.CS_RGBA, 255)
            self.state.append((j, d))
        info = self.state[0][0].info()
        samp = self.state[0][0].sampling()
        g = O.OraclePort().geometry(info["width"], info["height"], info["max_h"] * 8, info["max_v"] * 8, logo.shape[1],
                                    logo.shape[0], ALIGN_TOP_LEFT, 0, 0)
        dims = O.OraclePort().compiled_dims(O.make_layout(info["colorspace"], samp), g["blockoffset_x"], g["blockoffset_y"],
                                            g["crop_w"], g["crop_h"])
        self.blocks_per_image = sum(w * h for w, h in dims)
        self.geom = g
        self.image_blocks = sum(self.state[0][0].comp_info(c)["wreal"] * self.state[0][0].comp_info(c)["hreal"] for c in range(info["ncomp"]))

    def _run_threads(self, work, threads: int) -> float:
        errs = []

        def guarded(t):
            try:
                work(t)
            except Exception as e:  # noqa: BLE001
                errs.append(e)

        ts = [threading.Thread(target=guarded, args=(t,)) for t in range(threads)]
        t0 = time.perf_counter()
        [t.start() for t in ts]
        [t.join() for t in ts]
        dt = time.perf_counter() - t0
        if errs:
            raise RuntimeError(f"reference arm failed: {errs[:3]}")
        return dt

    def step(self, images_per_thread: int = 1, threads: int | None = None) -> float:
        """mj_compose (dropon compile + blend, src/compose.c:33) of images_per_thread images on every thread; wall seconds"""

        def work(t):
            j, d = self.state[t]
            for _ in range(images_per_thread):
                rv = j.compose(d, ALIGN_TOP_LEFT, 0, 0)
                if rv != 0:
                    raise RuntimeError(f"mj_compose -> {rv}")

        return self._run_threads(work, threads or self.threads)

    def split(self, images_per_thread: int, threads: int) -> dict:
        """mj_compile_dropon (src/dropon.c:325) and mj_compose_with_mask (src/compose.c:237) timed separately, as
        mj_compose calls them; per-thread seconds are summed per phase, rates are summed over the threads"""
        g = self.geom
        info, samp = self.state[0][0].info(), self.state[0][0].sampling()
        comp_s, blend_s = [0.0] * threads, [0.0] * threads

        def work(t):
            j, d = self.state[t]
            for _ in range(images_per_thread):
                t0 = time.perf_counter()
                rv, cd = self.lib.compile_handle(d, info["colorspace"], samp, g["blockoffset_x"], g["blockoffset_y"],
                                                 (g["crop_x"], g["crop_y"], g["crop_w"], g["crop_h"]))
                t1 = time.perf_counter()
                if rv != 0:
                    raise RuntimeError(f"mj_compile_dropon -> {rv}")
                rv = self.lib.compose_with_mask(j, cd, g["block_x"], g["block_y"])
                t2 = time.perf_counter()
                self.lib.free_compiled(cd)
                if rv != 0:
                    raise RuntimeError(f"mj_compose_with_mask -> {rv}")
                comp_s[t] += t1 - t0
                blend_s[t] += t2 - t1

        wall = self._run_threads(work, threads)
        blocks = images_per_thread * self.blocks_per_image
        return {"threads": threads, "images": images_per_thread * threads, "wall_s": wall,
                "compile_ms_per_image": 1e3 * sum(comp_s) / (images_per_thread * threads),
                "blend_ms_per_image": 1e3 * sum(blend_s) / (images_per_thread * threads),
                "blend_only_mblocks_per_s": sum(blocks / b for b in blend_s if b > 0) / 1e6,
                "compile_and_blend_mblocks_per_s": images_per_thread * threads * self.blocks_per_image / wall / 1e6}

    def effects(self, threads: int, reps: int = 2) -> dict:
        """each mj_effect_* (src/effect.c:28-222) on the decoded 1080p images, per-thread seconds summed"""
        out = {}
        nblocks = self.image_blocks
        for name, call in (("luminance_40", lambda j: j.luminance(40)), ("tint_30_m30", lambda j: j.tint(30, -30)),
                           ("pixelate", lambda j: j.pixelate()), ("grayscale", lambda j: j.grayscale())):
            secs = [0.0] * threads

            def work(t):
                j = self.state[t][0]
                for _ in range(reps):
                    t0 = time.perf_counter()
                    rv = call(j)
                    secs[t] += time.perf_counter() - t0
                    if rv != 0:
                        raise RuntimeError(f"mj_effect {name} -> {rv}")

            self._run_threads(work, threads)
            out[name] = {"ms_per_image": 1e3 * sum(secs) / (reps * threads), "mblocks_per_s": sum(reps * nblocks / x for x in secs if x > 0) / 1e6}
        return out

    def baseline_report(self, reps: int = 2) -> dict:
        """what BASELINE.md 3 / SURVEY 8d ask for: whole mj_compose, its two halves separately, each effect; one thread
        and every host thread"""
        T = self.threads
        self.step(1, T)  # warm: page in, first-touch
        rep = {"cpu_model": cpu_model(), "host_threads": T}
        for label, th in (("1_thread", 1), (f"{T}_threads", T)):
            dt = self.step(reps, th)
            rep[label] = {"mj_compose_mblocks_per_s": th * reps * self.blocks_per_image / dt / 1e6,
                          "mj_compose_images_per_s": th * reps / dt, "split": self.split(reps, th)}
        rep["effects_1_thread"] = self.effects(1)
        rep[f"effects_{T}_threads"] = self.effects(T)  # last: grayscale zeroes the chroma planes of the arm's images
        return rep


def reference_file_pipeline(jpegs, logo, threads: int, images: int) -> float:
    """decode -> mj_compose -> encode with the unmodified reference, `images` JPEGs over `threads` host threads
    (one image at a time per thread, as a caller of the reference would); returns wall seconds"""
    from oracle import oracle_py as O

    lib = O.Reference()
    drops = [lib.dropon_from_raw(logo, O.CS_RGBA, 255) for _ in range(threads)]
    errs = []

    def work(t):
        for i in range(t, images, threads):
            j = lib.read_jpeg(jpegs[i % len(jpegs)])
            if j.compose(drops[t], ALIGN_TOP_LEFT, 0, 0) != 0:
                errs.append(i)
            j.write(0)

    ts = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    [t.start() for t in ts]
    [t.join() for t in ts]
    dt = time.perf_counter() - t0
    if errs:
        raise RuntimeError(f"reference pipeline failed on images {errs[:3]}")
    return dt


def host_threads() -> int:
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    return max(1, min(n, 64))


def run_reference_arm(args, rank: int, world: int):
    if rank != 0:
        return
    jpegs, logo = make_inputs(min(N_BASES, 4))
    threads = host_threads()
    arm = CpuArm(jpegs, logo, threads)
    for _ in range(args.warmup):
        arm.step(1)
    total = 0.0
    for _ in range(args.steps):
        total += arm.step(1)
    images = threads * args.steps
    blocks = images * arm.blocks_per_image
    mbps = blocks / total / 1e6
    detail = None
    if not args.no_cpu_baseline:
        detail = arm.baseline_report()
    line = {
        "impl": "reference", "metric": "composited_mblocks_per_s", "value": mbps, "unit": "Mblocks/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int16+fp32", "data": "synthetic",
        "images_per_s": images / total,
        "config": {"workload": WORKLOAD,
                   "sample": f"{threads} images per step (one per host thread)", "inputs": "host-resident decoded coefficients"},
        "cpu_baseline": {"value": mbps, "unit": "Mblocks/s", "cores": threads, "kind": arm.kind, "cpu_model": cpu_model(),
                         "sample": f"{images} x mj_compose (dropon compile + blend) over {threads} threads",
                         "detail": detail},
        "e2e": {"value": mbps, "unit": "Mblocks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------
# the B200 arm
# ------------------------------------------------------------------------------------------


def kernel_source_sha() -> str:
    """fingerprint of the K2 sources: profiles/k2_traffic.json carries the one its ncu capture was taken from"""
    import hashlib

    h = hashlib.sha256()
    d = os.path.join(ROOT, "libmodjpeg_b200", "csrc")
    for name in ("k2_compose.cu", "k2_generic_op.cu", "k2_common.cuh", "k2_umma.cuh", "mjx_math.cuh"):
        with open(os.path.join(d, name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def parity_block(engine, slab, base_flat, jpegs, logo, shapes, plane_bytes, comp_off, lo, k, step, dev) -> dict:
    """north_star: differing-coefficient count and decoded-pixel PSNR, reported.  The first k images of the resident batch
    get their pristine planes back, ONE launch of the timed path runs over the whole batch, and those k images are
    compared with what the unmodified reference (oracle/_ref, the checker) makes of the same inputs."""
    import io

    import torch
    from PIL import Image

    from oracle import oracle_py as O

    ref = O.Reference()
    dref = ref.dropon_from_raw(logo, O.CS_RGBA, 255)
    for i in range(k):
        slab[i] = torch.from_numpy(base_flat[(lo + i) % N_BASES]).to(dev)
    step()
    torch.cuda.synchronize(dev)
    got_all = slab[:k].cpu().numpy()
    n = changed = differ = 0
    mx = 0
    psnrs = []
    for i in range(k):
        jb = jpegs[(lo + i) % N_BASES]
        jr = ref.read_jpeg(jb)
        before = jr.planes()
        if jr.compose(dref, ALIGN_TOP_LEFT, 0, 0) != 0:
            raise RuntimeError("reference mj_compose failed")
        want = jr.planes()
        jg = ref.read_jpeg(jb)
        for c, (r, st) in enumerate(shapes):
            got = got_all[i, int(comp_off[c]):int(comp_off[c]) + plane_bytes[c]].view(np.int16).reshape(r, st, 64)
            d = got.astype(np.int32) - want[c].astype(np.int32)
            n += d.size
            differ += int((d != 0).sum())
            mx = max(mx, int(np.abs(d).max()))
            changed += int((want[c] != before[c]).sum())
            jg.set_plane(c, got)
        pa = np.asarray(Image.open(io.BytesIO(jg.write(0))), np.float64)
        pb = np.asarray(Image.open(io.BytesIO(jr.write(0))), np.float64)
        mse = float(((pa - pb) ** 2).mean())
        psnrs.append(None if mse == 0 else 10 * np.log10(255.0 ** 2 / mse))
    finite = [x for x in psnrs if x is not None]
    return {"images": k, "coefficients": n, "changed_by_compose": changed, "differing_from_reference": differ,
            "differing_rate_of_changed": differ / max(1, changed), "max_abs_diff_in_quant_steps": mx,
            "decoded_psnr_db_vs_reference_min": round(min(finite), 2) if finite else "identical",
            "images_decoding_identically": sum(1 for x in psnrs if x is None),
            "against": "oracle/_ref = the unmodified reference built here, mj_compose on the same JPEGs",
            "path": "one launch of the timed kernels over the whole resident batch; the first images compared"}


ORIG_AFFINITY = None


def bind_to_gpu_numa_node(device: int) -> None:
    """Pin this rank's host threads (and so its first-touch page-locked buffers) to the CPUs NVML reports as
    local to the GPU: the e2e path streams host memory over PCIe, which a remote NUMA node would throttle."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        global ORIG_AFFINITY
        ORIG_AFFINITY = set(os.sched_getaffinity(0))
        allowed = cpus & ORIG_AFFINITY
        if allowed:
            os.sched_setaffinity(0, allowed)
            log(f"[gpu {device}] host threads bound to {len(allowed)} GPU-local CPUs")
    except Exception as e:  # affinity is an optimisation, never a requirement
        log(f"[gpu {device}] no NUMA binding ({type(e).__name__}: {e})")


def run_b200_arm(args, rank: int, local_rank: int, world: int):
    import torch

    import libmodjpeg_b200 as M
    from libmodjpeg_b200 import capi
    from libmodjpeg_b200.batch import shard_range

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        # keep stdout to the one JSON line: NCCL's version banner / debug output goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)

    bind_to_gpu_numa_node(local_rank)
    engine = M.Engine(local_rank)
    # torch is plumbing here: it owns the device memory and the stream the kernels are launched
    # on, so that torch.cuda.Event brackets exactly the engine's launches
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    engine.set_stream(stream.cuda_stream)

    # ---- inputs -------------------------------------------------------------------------
    t_setup = time.time()
    jpegs, logo = make_inputs(N_BASES)
    n_total = args.images_per_gpu * world
    lo, hi = shard_range(n_total, rank, world)
    n = hi - lo
    bases = []
    for jb in jpegs:
        j = M.Jpeg()
        assert j.read_jpeg_from_memory(jb) == 0
        bases.append((j.planes(), [j.qtable(c) for c in range(3)], j.info(), j.sampling()))
    info, samp = bases[0][2], bases[0][3]
    shapes = [p.shape[:2] for p in bases[0][0]]  # (rows, stride) per component
    plane_bytes = [r * s * 128 for r, s in shapes]
    comp_off = np.concatenate([[0], np.cumsum(plane_bytes)[:-1]]).astype(np.int64)
    image_bytes = int(sum(plane_bytes))

    # device slab: image i of this rank is base (lo + i) % N_BASES
    slab = torch.empty((n, image_bytes), dtype=torch.uint8, device=dev)
    base_flat = [np.concatenate([p.reshape(-1).view(np.uint8) for p in b[0]]) for b in bases]
    for k in range(N_BASES):
        idx = [i for i in range(n) if (lo + i) % N_BASES == k]
        if idx:
            slab[torch.tensor(idx, device=dev)] = torch.from_numpy(base_flat[k]).to(dev)
    ptrs = [[slab.data_ptr() + i * image_bytes + int(comp_off[c]) for c in range(3)] for i in range(n)]
    qt = np.stack([np.stack(bases[(lo + i) % N_BASES][1]) for i in range(n)])
    descs = capi.make_image_descs(ptrs, [s for _, s in shapes], [r for r, _ in shapes], qt)
    descs_dev = torch.from_numpy(descs.view(np.uint8).reshape(-1).copy()).to(dev)

    # ---- K1: compile the dropon once for this image geometry -------------------------------
    i3 = np.ascontiguousarray(logo[:, :, :3])
    a3 = np.ascontiguousarray(np.repeat(logo[:, :, 3:4], 3, 2))
    g = capi.geometry(info["width"], info["height"], info["max_h"] * 8, info["max_v"] * 8, logo.shape[1], logo.shape[0],
                      ALIGN_TOP_LEFT, 0, 0)
    layout = M.Layout.make(info["colorspace"], samp)
    cd = engine.dropon_compile(i3, a3, M.CS_RGB, layout, (g["blockoffset_x"], g["blockoffset_y"]),
                               (g["crop_x"], g["crop_y"], g["crop_w"], g["crop_h"]))
    counts = cd.class_counts()
    blocks_per_image = cd.blocks
    # algorithmic bytes of one launch: image traffic per block by class, plus the unique
    # compiled-dropon bytes once per launch (class word, D for non-T, W for G)
    img_bytes_alg = n * (counts["OPAQUE"] * 128 + counts["U"] * 256 + counts["G"] * 256)
    drop_bytes_alg = blocks_per_image * 4 + (counts["OPAQUE"] + counts["U"] + counts["G"]) * 128 + counts["G"] * 128
    alg_bytes = img_bytes_alg + drop_bytes_alg
    torch.cuda.synchronize(dev)
    log(f"[rank {rank}] setup {time.time() - t_setup:.1f}s: {n} images x {blocks_per_image} blocks, classes {counts}, "
        f"{n * image_bytes / 1e9:.2f} GB resident")

    def step():
        engine.compose_batch_device(descs_dev.data_ptr(), n, cd, g["block_x"], g["block_y"])

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident timing ---------------------------------------------------------------
    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    launches0 = engine.kernel_launches
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    t0 = time.time()
    ev[0].record(stream)
    for k in range(args.steps):
        step()
        ev[k + 1].record(stream)
    barrier()
    t1 = time.time()
    launches = engine.kernel_launches - launches0
    total_ms = ev[0].elapsed_time(ev[-1])
    per_launch_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms_max = float(tmax.item())

    # ---- the two K2 kernels on their own (same launches, one class group masked off) ------------
    def time_masked(mask: int) -> float:
        engine.set_class_mask(mask)
        for _ in range(3):
            step()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        torch.cuda.synchronize(dev)
        engine.set_class_mask(3)
        return e0.elapsed_time(e1) / args.steps

    ms_simple = time_masked(1) if counts["OPAQUE"] + counts["U"] else 0.0
    ms_generic = time_masked(2) if counts["G"] else 0.0
    # the other two ways the G class can run, timed in the same process right behind (reported under roofline.alternatives;
    # the headline numbers above are the library's default)
    alternatives = {}
    if counts["G"] and rank == 0 and engine.tensor_core_active(n, counts["G"]):
        default_mode = getattr(engine, "_tc_mode", 1)
        for name, mode in (("tensor_core_range_vouched", 2), ("fp32_kernel", 0)):
            if mode == default_mode:
                continue
            try:
                engine.set_tensor_core(mode)
                g_ms = time_masked(2)
                s_ms = time_masked(3)
                alternatives[name] = {"mode": mode, "g_kernel_ms": g_ms, "step_ms": s_ms}
            finally:
                engine.set_tensor_core(default_mode)
        for _ in range(3):
            step()
        torch.cuda.synchronize(dev)
    # clocks / throttle reasons sampled from just before the timed steps to the end of the per-kernel timings
    # (the same kernels back to back: the GPU is under the bench's load for the whole window)
    if rank == 0:
        # nvidia-smi needs a few hundred ms before its first line on some boxes: keep the same kernels running (untimed)
        # until a handful of samples exist, so that the clocks reported are always clocks under this load
        t_wait = time.time()
        while sum(1 for t, _ in sampler.lines if t >= t0 - 0.05) < 5 and time.time() - t_wait < 4.0 and sampler.proc is not None:
            for _ in range(10):
                step()
            torch.cuda.synchronize(dev)
    clocks = sampler.stop(t0, time.time()) if rank == 0 else None

    # ---- the other two kernels of the path, on the same resident batch (rank 0 reports) -----------
    other = {}
    if not args.no_other_kernels:
        engine.set_stream(stream.cuda_stream)

        def timed(fn, reps):
            fn()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(reps):
                fn()
            e1.record(stream)
            torch.cuda.synchronize(dev)
            return e0.elapsed_time(e1) / reps

        # K1: the 1920x1080 RGBA logo compiled for the 4:2:0 YCbCr layout, pixels already in HBM
        # (device time of k1_compile_kernel + the list/prepare kernels + the two small D2H count reads)
        px = torch.from_numpy(np.concatenate([i3.reshape(-1), a3.reshape(-1)])).to(dev)
        npx = i3.size

        def k1():
            c = engine.dropon_compile(None, None, M.CS_RGB, layout, (g["blockoffset_x"], g["blockoffset_y"]),
                                      (g["crop_x"], g["crop_y"], g["crop_w"], g["crop_h"]),
                                      device_pixels=(px.data_ptr(), px.data_ptr() + npx, logo.shape[1], logo.shape[0]))
            c.free()

        ms = timed(k1, 5)
        k1_bytes = 2 * npx + blocks_per_image * 64 * 2 * 2  # two 3-byte pixel buffers read, D and W int16 written
        other["k1_dropon_compile"] = {"ms": ms, "pixels": npx // 3, "algorithmic_bytes": k1_bytes,
                                      "achieved_gbs": k1_bytes / (ms * 1e-3) / 1e9, "mpixels_per_s": npx / 3 / (ms * 1e-3) / 1e6,
                                      "note": "whole mjx_dropon_compile call (K1 + list build + cudaMalloc/cudaFree + stream syncs); launch/sync-bound at this size"}
        # K3 on every image of the batch: DC-only pipeline (luminance + tint) and a rewrite pipeline (pixelate all components)
        yb = shapes[0][0] * shapes[0][1]
        cb = shapes[1][0] * shapes[1][1]
        ms = timed(lambda: engine.effects_batch_device(descs_dev.data_ptr(), n, 3, [(3, 0, 40), (3, 1, 30), (3, 2, -30)]), 5)
        dc_blocks = n * (yb + 2 * cb)
        other["k3_effects_dc_only"] = {"ms": ms, "ops": "luminance 40 + tint 30,-30 fused", "blocks": dc_blocks, "algorithmic_bytes": dc_blocks * 4,
                                       "sector_bytes": dc_blocks * 64, "achieved_gbs_sector": dc_blocks * 64 / (ms * 1e-3) / 1e9,
                                       "gblocks_per_s": dc_blocks / (ms * 1e-3) / 1e9,
                                       "note": "2 B read + 2 B written per block; DRAM moves a 32 B sector each way"}
        ms = timed(lambda: engine.effects_batch_device(descs_dev.data_ptr(), n, 3, [(2, 0, 0), (2, 1, 0), (2, 2, 0)]), 5)
        other["k3_effects_pixelate"] = {"ms": ms, "blocks": dc_blocks, "algorithmic_bytes": dc_blocks * 130,
                                        "achieved_gbs": dc_blocks * 130 / (ms * 1e-3) / 1e9, "gblocks_per_s": dc_blocks / (ms * 1e-3) / 1e9}

        # K4: Huffman coding of every image of the batch on the device (the entropy encoder behind mj_write_jpeg_to_memory,
        # reference src/image.c:194): coefficient planes resident in HBM -> entropy-coded segments in HBM
        try:
            scan = capi.standard_scan(info["width"], info["height"], samp)
            cap = 1 << 20
            seg = torch.empty((n, cap), dtype=torch.uint8, device=dev)
            sizes = torch.zeros(n, dtype=torch.int32, device=dev)
            ms = timed(lambda: engine.huffman_encode_batch_device(descs_dev.data_ptr(), n, scan, seg.data_ptr(), cap, sizes.data_ptr()), 3)
            sz = sizes.cpu().numpy().view(np.uint32)
            coded = int((sz != 0xFFFFFFFF).sum())
            out_bytes = int(sz[sz != 0xFFFFFFFF].astype(np.int64).sum())
            in_bytes = n * image_bytes
            # the same encoder on the host: libjpeg's jpeg_write_coefficients through the library's host path, all host threads
            hj = M.Jpeg()
            assert hj.read_jpeg_from_memory(jpegs[0]) == 0
            t_h = time.perf_counter()
            reps_h = 8
            for _ in range(reps_h):
                assert hj.write_jpeg_to_memory(0)[0] == 0
            host_ms = 1e3 * (time.perf_counter() - t_h) / reps_h
            other["k4_huffman_encode"] = {"ms": ms, "images": n, "coded": coded, "images_per_s": n / (ms * 1e-3), "bytes_in": in_bytes, "bytes_out": out_bytes,
                                          "algorithmic_bytes": in_bytes + out_bytes, "achieved_gbs": (in_bytes + out_bytes) / (ms * 1e-3) / 1e9,
                                          "host_libjpeg_ms_per_image_1_thread": host_ms,
                                          "note": "six launches (bit count, scan, emit, 0xFF count, scan, byte stuffing); the planes are read twice, "
                                                  "algorithmic bytes count them once; host figure: mj_write_jpeg_to_memory of one such image on one core "
                                                  "(markers + libjpeg's encode_mcu_huff), the encoder the reference uses"}
            del seg
        except Exception as e:  # noqa: BLE001
            other["k4_huffman_encode"] = {"unavailable": f"{type(e).__name__}: {str(e)[:160]}"}

        # K5: Huffman decoding of every image of the batch on the device (the entropy decoder behind mj_read_jpeg_from_memory,
        # reference src/image.c:94): entropy-coded segments resident in HBM -> coefficient planes in HBM (the batch's own slab:
        # what is decoded is what was there before the timed steps)
        try:
            parsed = [capi.scan_from_jpeg(jb) for jb in jpegs]
            segs = [jb[off:] for jb, (_, off, _) in zip(jpegs, parsed)]
            seg_off = np.concatenate([[0], np.cumsum([(len(sg) + 255) // 256 * 256 for sg in segs])]).astype(np.int64)
            blob = np.zeros(int(seg_off[-1]) + 256, np.uint8)
            for k, sg in enumerate(segs):
                blob[int(seg_off[k]):int(seg_off[k]) + len(sg)] = np.frombuffer(sg, np.uint8)
            blob_dev = torch.from_numpy(blob).to(dev)
            offs = np.array([seg_off[(lo + i) % N_BASES] for i in range(n)], np.uint64)
            lens = np.array([len(segs[(lo + i) % N_BASES]) for i in range(n)], np.uint32)
            st_dev = torch.zeros(n, dtype=torch.int32, device=dev)
            ms = timed(lambda: engine.huffman_decode_batch_device(blob_dev.data_ptr(), offs, lens, n, parsed[0][0], descs_dev.data_ptr(), st_dev.data_ptr()), 3)
            ok = int((st_dev.cpu().numpy() == 0).sum())
            same = bool(torch.equal(slab[0], torch.from_numpy(base_flat[lo % N_BASES]).to(dev)))
            hj = M.Jpeg()
            t_h = time.perf_counter()
            for _ in range(8):
                assert hj.read_jpeg_from_memory(jpegs[0]) == 0
            host_ms = 1e3 * (time.perf_counter() - t_h) / 8
            in_b, out_b = int(lens.astype(np.int64).sum()), n * image_bytes
            other["k5_huffman_decode"] = {"ms": ms, "images": n, "decoded": ok, "subsequences_per_image": int((lens[0] * 8 + 1023) // 1024), "first_image_equals_libjpeg": same, "images_per_s": n / (ms * 1e-3),
                                          "bytes_in": in_b, "bytes_out": out_b, "algorithmic_bytes": in_b + out_b,
                                          "achieved_gbs": (in_b + out_b) / (ms * 1e-3) / 1e9, "host_libjpeg_ms_per_image_1_thread": host_ms,
                                          "note": "one launch, one CTA per image: planes zeroed, un-stuffing, entry states of the 1024-bit subsequences settled "
                                                  "in rounds, write pass, DC prefix sums; host figure: mj_read_jpeg_from_memory of one such image on one core "
                                                  "(libjpeg's jpeg_read_coefficients), the decoder the reference uses"}
            del blob_dev
        except Exception as e:  # noqa: BLE001
            other["k5_huffman_decode"] = {"unavailable": f"{type(e).__name__}: {str(e)[:160]}"}

    # ---- parity of the timed path against the unmodified reference (rank 0, outside every timed region) ----
    parity = None
    if rank == 0 and not args.no_parity:
        try:
            parity = parity_block(engine, slab, base_flat, jpegs, logo, shapes, plane_bytes, comp_off, lo, min(8, n), step, dev)
        except Exception as e:  # noqa: BLE001 -- the checker is reported, never required for the GPU number
            parity = {"unavailable": f"{type(e).__name__}: {str(e)[:160]}"}

    # ---- like for like with the reference's mj_compose: dropon compile + blend PER IMAGE, device-resident --------
    per_image = None
    if rank == 0 and not args.no_other_kernels:
        engine.set_stream(stream.cuda_stream)
        px_d = torch.from_numpy(np.concatenate([i3.reshape(-1), a3.reshape(-1)])).to(dev)
        npx_d = i3.size
        k_img = min(n, 64)

        def one_image(i):
            c = engine.dropon_compile(None, None, M.CS_RGB, layout, (g["blockoffset_x"], g["blockoffset_y"]),
                                      (g["crop_x"], g["crop_y"], g["crop_w"], g["crop_h"]),
                                      device_pixels=(px_d.data_ptr(), px_d.data_ptr() + npx_d, logo.shape[1], logo.shape[0]))
            engine.compose_batch_device(descs_dev.data_ptr() + i * capi.IMAGE_DESC_DTYPE.itemsize, 1, c, g["block_x"], g["block_y"])
            c.free()

        one_image(0)
        torch.cuda.synchronize(dev)
        tp0 = time.perf_counter()
        for i in range(k_img):
            one_image(i)
        torch.cuda.synchronize(dev)
        tp = time.perf_counter() - tp0
        per_image = {"value": k_img * blocks_per_image / tp / 1e6, "unit": "Mblocks/s", "images": k_img, "ms_per_image": 1e3 * tp / k_img,
                     "what": "K1 (mjx_dropon_compile) + K2 on ONE device-resident image per call, nothing reused between images: the work "
                             "the reference's mj_compose does per call (compile + blend, src/compose.c:155-177), wall clock"}

    # ---- e2e: host planes through mjx_compose_batch_host ----------------------------------------
    n_e2e = min(n, args.e2e_images)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    pinned = engine.host_alloc(n_e2e * image_bytes)
    pview = pinned.reshape(n_e2e, image_bytes)
    for i in range(n_e2e):
        pview[i] = base_flat[(lo + i) % N_BASES]
    items = (capi.HostImage * n_e2e)()
    keep_q = []
    for i in range(n_e2e):
        it = capi.HostImage()
        for c in range(3):
            it.plane[c] = pinned.ctypes.data + i * image_bytes + int(comp_off[c])
            it.stride_blocks[c] = shapes[c][1]
            it.rows[c] = shapes[c][0]
            it.wreal[c] = shapes[c][1]
            it.hreal[c] = shapes[c][0]
            q = np.ascontiguousarray(bases[(lo + i) % N_BASES][1][c])
            keep_q.append(q)
            it.q[c] = q.ctypes.data
        items[i] = it
    engine.set_stream(None)

    def e2e_run(zero_copy: bool):
        engine.set_zero_copy(zero_copy)
        engine.compose_batch_host(items, n_e2e, cd, g["block_x"], g["block_y"])  # untimed: staging pools, first touch of the page-locked planes
        barrier()
        l0 = engine.kernel_launches
        te0 = time.perf_counter()
        for _ in range(e2e_steps):
            engine.compose_batch_host(items, n_e2e, cd, g["block_x"], g["block_y"])
        torch.cuda.synchronize(dev)
        te = time.perf_counter() - te0
        nl = engine.kernel_launches - l0
        tm = torch.tensor([te], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        return float(tm.item()), nl

    nsum = torch.tensor([n_e2e], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(nsum)
    t_staged, _ = e2e_run(False)
    t_e2e, e2e_launches = e2e_run(True)
    e2e_blocks = float(nsum.item()) * blocks_per_image * e2e_steps
    e2e_mbps = e2e_blocks / t_e2e / 1e6
    e2e_staged_mbps = e2e_blocks / t_staged / 1e6
    roi_bytes = sum(cd.dims(c)[0] * cd.dims(c)[1] * 128 for c in range(3))
    engine.host_free(pinned)

    # ---- what bounds the e2e tier when several ranks stream host memory at once: the host's own copy bandwidth --------
    host_copy = None
    try:
        nb_copy = 256 << 20
        src, dst = np.ones(nb_copy, np.uint8), np.empty(nb_copy, np.uint8)
        np.copyto(dst, src)
        barrier()
        th0 = time.perf_counter()
        for _ in range(4):
            np.copyto(dst, src)
        th = time.perf_counter() - th0
        tt = torch.tensor([th], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        host_copy = {"gbs_all_ranks": world * 4 * 2 * nb_copy / float(tt.item()) / 1e9, "ranks": world,
                     "what": "every rank copies 256 MB host -> host four times at the same moment (one thread each, bytes read + written): "
                             "the host DRAM the PCIe streams of all ranks share"}
        del src, dst
    except Exception as e:  # noqa: BLE001
        host_copy = {"unavailable": str(e)[:120]}

    # ---- tier (iii): JPEG bytes in -> JPEG bytes out through mj_compose_batch, every rank its share of host threads -------
    files = None
    if args.file_images > 0:
        if ORIG_AFFINITY:
            os.sched_setaffinity(0, ORIG_AFFINITY)  # entropy coding uses every host core, for both arms
        threads = max(1, host_threads() // world)
        d = M.Dropon()
        assert d.read_dropon_from_raw(logo, M.CS_RGBA, 255) == 0
        batch = [jpegs[i % len(jpegs)] for i in range(args.file_images)]
        capi.compose_batch((batch * 3)[:max(8 * threads, 512) + 8], d, ALIGN_TOP_LEFT, 0, 0, 0, nthreads=threads)  # warm both contexts' pools (two full windows and a bit)
        barrier()
        tm = {}
        rvb, status, outs = capi.compose_batch(batch, d, ALIGN_TOP_LEFT, 0, 0, 0, nthreads=threads, timing=tm)
        assert rvb == 0 and not any(status), (rvb, status[:8])
        tfm = torch.tensor([tm["call_s"]], dtype=torch.float64, device=dev)  # wall time of the C call (JPEG bytes in -> malloc()ed JPEG bytes out)
        if dist is not None:
            dist.all_reduce(tfm, op=dist.ReduceOp.MAX)
        tf = float(tfm.item())
        if rank == 0:
            files = {"images": args.file_images * world, "threads_per_rank": threads, "ranks": world, "images_per_s": args.file_images * world / tf,
                     "mblocks_per_s": args.file_images * world * blocks_per_image / tf / 1e6,
                     "path": ("mj_compose_batch per rank: markers read by libjpeg (thread pool) -> entropy-coded segments of a window to HBM -> K5 Huffman "
                              "decoding -> K1 once -> K2 -> K4 Huffman coding -> entropy-coded segments back, libjpeg's markers in front (thread pool); "
                              "JPEG bytes in, JPEG bytes out" if os.environ.get("MJX_GPU_HUFFMAN", "") != "0" and os.environ.get("MJX_GPU_DECODE", "") != "0" else
                              "mj_compose_batch per rank: libjpeg entropy decode (thread pool) -> K1 once -> whole planes of a window to HBM -> K2 -> "
                              "K4 Huffman coding on the device -> entropy-coded segments back, libjpeg's markers in front (thread pool); "
                              "JPEG bytes in, JPEG bytes out (MJX_GPU_DECODE=0)" if os.environ.get("MJX_GPU_HUFFMAN", "") != "0" else
                              "mj_compose_batch per rank: libjpeg entropy decode (thread pool) -> K1 once -> K2 one launch per window, zero-copy "
                              "on a page-locked slab -> libjpeg entropy encode (thread pool); JPEG bytes in, JPEG bytes out (MJX_GPU_HUFFMAN=0)"),
                     "bytes_in": sum(len(b) for b in batch) * world, "bytes_out": sum(len(o) for o in outs) * world}
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            try:
                nref = max(threads, min(args.file_images, 2 * threads))
                dtr = reference_file_pipeline(jpegs, logo, threads, nref)
                files["reference"] = {"images": nref, "threads": threads, "images_per_s": nref / dtr,
                                      "path": "unmodified reference: mj_read_jpeg_from_memory -> mj_compose -> mj_write_jpeg_to_memory per image"}
            except Exception as e:
                files["reference"] = {"unavailable": str(e)[:200]}

    # ---- CPU baseline beside it (rank 0, N = 1 only) ----------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            if ORIG_AFFINITY:
                os.sched_setaffinity(0, ORIG_AFFINITY)  # the reference arm gets every host core
            threads = host_threads()
            arm = CpuArm(jpegs[:4], logo, threads)
            arm.step(1)
            reps = 2
            dt = arm.step(reps)
            imgs = threads * reps
            cpu = {"value": imgs * arm.blocks_per_image / dt / 1e6, "unit": "Mblocks/s", "cores": threads, "kind": arm.kind,
                   "cpu_model": cpu_model(),
                   "sample": f"{imgs} x mj_compose (dropon compile + blend) of the same 1080p images over {threads} host threads",
                   "images_per_s": imgs / dt,
                   "detail": arm.baseline_report()}
        except Exception as e:  # the baseline is reported, never required for the GPU number
            cpu = {"value": None, "unit": "Mblocks/s", "cores": 0, "kind": "unavailable", "sample": str(e)[:200]}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
        launch_ms = statistics.mean(per_launch_ms)
        achieved = alg_bytes / (launch_ms * 1e-3) / 1e9
        # per kernel: image bytes by class + the compiled-dropon bytes that class needs, once per launch
        alg_generic = n * counts["G"] * 256 + counts["G"] * (256 + 4)
        alg_simple = n * (counts["OPAQUE"] * 128 + counts["U"] * 256) + (counts["OPAQUE"] + counts["U"]) * (128 + 4)
        # DRAM bytes per launch from the committed ncu capture -- only while it was taken from these very sources
        traffic = {}
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "k2_traffic.json")))
            if traffic.get("source_sha") != kernel_source_sha():
                traffic = {}
        except Exception:
            traffic = {}
        g_name = "k2_generic_op_kernel" if engine.tensor_core_active(n, counts["G"]) else "k2_generic_kernel"
        kernels = {
            g_name: {"class": "G", "launch_ms": ms_generic, "algorithmic_bytes": alg_generic,
                                  "achieved_gbs": alg_generic / (ms_generic * 1e-3) / 1e9 if ms_generic else None,
                                  "traffic": traffic.get(g_name),
                                  "path": ("tensor-core operator kernel (tcgen05, fp16 operands / fp32 accumulation; libmodjpeg_b200/csrc/k2_generic_op.cu) + "
                                           "its prepare / cache-check / redo launches; coefficient range checked in the kernel"
                                           if g_name == "k2_generic_op_kernel" else "fp32 kernel (libmodjpeg_b200/csrc/k2_compose.cu)")},
            "k2_simple_kernel": {"class": "OPAQUE+U", "launch_ms": ms_simple, "algorithmic_bytes": alg_simple,
                                 "achieved_gbs": alg_simple / (ms_simple * 1e-3) / 1e9 if ms_simple else None,
                                 "traffic": traffic.get("k2_simple_kernel"),
                                 "note": "write-only on this workload (cudaMemset reaches 3.9 TB/s on the same box, profiles/microbench/hbm_rw.txt)"},
        }
        dom = g_name if ms_generic >= ms_simple else "k2_simple_kernel"
        dom_gbs = kernels[dom]["achieved_gbs"] or 0.0
        blocks_all = n_total * blocks_per_image * args.steps
        line = {
            "metric": "composited_mblocks_per_s", "value": blocks_all / (total_ms_max * 1e-3) / 1e6, "unit": "Mblocks/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": total_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16+fp32", "data": "synthetic",
            "images_per_s": n_total * args.steps / (total_ms_max * 1e-3),
            "config": {"workload": WORKLOAD,
                       "images_per_gpu": args.images_per_gpu, "blocks_per_image": blocks_per_image, "class_mix": counts,
                       "resident_bytes_per_gpu": n * image_bytes,
                       "l2": "inputs (7.8 GB/GPU) larger than L2 (126 MB); no flush needed",
                       "parallelism": f"batch sharded by image over {world} GPU(s), no collective"},
            "roofline": {"bound": "hbm", "achieved": dom_gbs, "peak": peak, "unit": "GB/s", "frac": dom_gbs / peak,
                         "traffic": kernels[dom]["traffic"], "kernel": dom, "algorithmic_bytes_per_launch": kernels[dom]["algorithmic_bytes"],
                         "launch_ms": kernels[dom]["launch_ms"], "peak_source": peak_src,
                         "timing": "CUDA events on the launching stream around the kernel alone (other class group masked off), mean of the timed steps",
                         "step": {"achieved": achieved, "frac": achieved / peak, "algorithmic_bytes": alg_bytes, "ms": launch_ms,
                                  "what": "whole K2 step = the G-class kernel, then k2_simple_kernel (the operator kernel fills every SM's shared memory, so the two run one after the other; "
                                          "with the fp32 G kernel the simple kernel runs beside it on a side stream)"},
                         "kernels": kernels,
                         "alternatives": {k: dict(v, g_kernel_frac=alg_generic / (v["g_kernel_ms"] * 1e-3) / 1e9 / peak,
                                                  step_frac=alg_bytes / (v["step_ms"] * 1e-3) / 1e9 / peak,
                                                  value_mblocks_per_s=n * blocks_per_image / (v["step_ms"] * 1e-3) / 1e6) for k, v in alternatives.items()}},
            "parity": parity,
            "per_image_compile_and_blend": per_image,
            "other_kernels": other,
            "e2e_files": files,
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_mbps, "unit": "Mblocks/s",
                    # zero-copy: only touched blocks cross PCIe (G read + written, OPAQUE/U written; U also read)
                    "h2d_bytes_per_step": n_e2e * ((counts["G"] + counts["U"]) * 128 + 608),
                    "d2h_bytes_per_step": n_e2e * (counts["G"] + counts["U"] + counts["OPAQUE"]) * 128,
                    "images_per_s": float(nsum.item()) * e2e_steps / t_e2e,
                    "images_per_step_per_gpu": n_e2e, "steps": e2e_steps,
                    "path": "mjx_compose_batch_host on page-locked host planes: one K2 launch reads/writes the touched blocks over PCIe (zero-copy)",
                    "timing": "wall clock around mjx_compose_batch_host, max over ranks", "gpu_launches": e2e_launches,
                    "host_copy": host_copy,
                    "staged": {"value": e2e_staged_mbps, "unit": "Mblocks/s", "h2d_bytes_per_step": n_e2e * (roi_bytes + 608),
                               "d2h_bytes_per_step": n_e2e * roi_bytes,
                               "path": "same call with zero-copy off: region under the dropon copied H2D, blended, copied D2H (3-stream pipeline)"}},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--images-per-gpu", type=int, default=1250)
    ap.add_argument("--e2e-images", type=int, default=1250)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-kernels", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--file-images", type=int, default=1024, help="JPEGs pushed through mj_compose_batch (tier iii); 0 = skip")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    capture_stdout()
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    run_b200_arm(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
