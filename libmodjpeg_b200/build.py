"""Build the two in-tree shared libraries of the B200 engine.

* ``lib/libmjx.so``      -- CUDA kernels K1/K2/K3 + the kernel-level C-ABI (include/mjx.h),
  compiled by nvcc for sm_100a only (``-gencode arch=compute_100a,code=sm_100a -lineinfo``),
  static cudart so that the library has no CUDA runtime dependency of its own.
* ``lib/libmodjpeg.so``  -- the host boundary (include/libmodjpeg.h, the reference's public API)
  in C, linked against libmjx.so and the image's libjpeg-turbo (Pillow's pillow.libs).

nvcc cross-compiles without a GPU, so this runs on the CPU-only build container.  The built
files are git-ignored but travel to the GPU box with gpurun.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "lib")
OBJ = os.path.join(LIB, "obj")
INCLUDE = os.path.join(ROOT, "include")
JPEG_INC = os.path.join(ROOT, "third_party", "jpeg62")
PNG_INC = os.path.join(ROOT, "third_party", "png16")

CUDA_SOURCES = ["mjx_api.cu", "k1_dropon.cu", "k1_lists.cu", "k2_compose.cu", "k2_generic_op.cu", "k3_effects.cu", "k4_huffman.cu", "k5_huffman_decode.cu"]
HOST_SOURCES = ["mj_jpegio.c", "mj_image.c", "mj_dropon.c", "mj_compose.c", "mj_effect.c", "mj_device.c", "mj_batch.c", "mj_coalesce.c"]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def jpeg_runtime() -> str:
    """The libjpeg-turbo (ABI 62) shared object bundled with Pillow."""
    import PIL

    d = os.path.join(os.path.dirname(os.path.dirname(PIL.__file__)), "pillow.libs")
    cands = sorted(glob.glob(os.path.join(d, "libjpeg-*.so.62*")))
    if not cands:
        raise RuntimeError(f"no libjpeg .so.62 in {d}")
    return cands[0]


def png_runtime() -> str | None:
    """The libpng16 shared object bundled with Pillow (None: build without PNG dropon support)."""
    import PIL

    d = os.path.join(os.path.dirname(os.path.dirname(PIL.__file__)), "pillow.libs")
    cands = sorted(glob.glob(os.path.join(d, "libpng16-*.so.16*")))
    return cands[0] if cands else None


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _run(cmd: list[str], verbose: bool) -> None:
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("build failed")
    if verbose and (r.stdout or r.stderr):
        print(r.stdout + r.stderr)


def build(verbose: bool = False, force: bool = False, ptxas_verbose: bool = False) -> dict:
    os.makedirs(OBJ, exist_ok=True)
    headers = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h")) + \
        glob.glob(os.path.join(CSRC, "host", "*.h")) + glob.glob(os.path.join(JPEG_INC, "*.h")) + glob.glob(os.path.join(PNG_INC, "*.h")) + [__file__]
    nvcc = _nvcc()

    # ---- libmjx.so -------------------------------------------------------------------
    objs = []
    for src in CUDA_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _newer(o, [s] + headers):
            flags = list(NVCC_FLAGS)
            if ptxas_verbose:
                flags += ["-Xptxas", "-v"]
            _run([nvcc] + flags + ["-I", INCLUDE, "-I", CSRC, "-DMJX_BUILD", "-c", s, "-o", o], verbose or ptxas_verbose)
    mjx_so = os.path.join(LIB, "libmjx.so")
    if force or _newer(mjx_so, objs):
        _run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-o", mjx_so] + objs +
             ["-Xlinker", "--version-script=" + os.path.join(CSRC, "mjx.map")], verbose)

    # ---- libmodjpeg.so ---------------------------------------------------------------
    jpeg_so = jpeg_runtime()
    link = os.path.join(LIB, "libjpeg.so")
    if not os.path.islink(link) or os.readlink(link) != jpeg_so:
        if os.path.lexists(link):
            os.remove(link)
        os.symlink(jpeg_so, link)
    png_so = png_runtime()
    png_flags, png_link = [], []
    if png_so:
        plink = os.path.join(LIB, "libpng16.so")
        if not os.path.islink(plink) or os.readlink(plink) != png_so:
            if os.path.lexists(plink):
                os.remove(plink)
            os.symlink(png_so, plink)
        png_flags = ["-DWITH_LIBPNG", "-I", PNG_INC]
        png_link = ["-lpng16"]
    hobjs = []
    for src in HOST_SOURCES:
        s = os.path.join(CSRC, "host", src)
        o = os.path.join(OBJ, src.replace(".c", ".host.o"))
        hobjs.append(o)
        if force or _newer(o, [s] + headers):
            _run(["gcc", "-std=gnu11", "-O2", "-Wall", "-Wextra", "-Wno-unused-parameter", "-Wno-clobbered", "-fPIC",
                  "-I", INCLUDE, "-I", JPEG_INC, "-I", os.path.join(CSRC, "host")] + png_flags + ["-c", s, "-o", o], verbose)
    mj_so = os.path.join(LIB, "libmodjpeg.so")
    if force or _newer(mj_so, hobjs + [mjx_so]):
        _run(["gcc", "-shared", "-o", mj_so] + hobjs +
             ["-L", LIB, "-lmjx", "-ljpeg"] + png_link + ["-lpthread", "-lm",
              "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + os.path.dirname(jpeg_so),
              "-Wl,--version-script=" + os.path.join(CSRC, "host", "libmodjpeg.map")], verbose)
    return {"libmjx": mjx_so, "libmodjpeg": mj_so, "libjpeg": jpeg_so}


if __name__ == "__main__":
    out = build(verbose=True, force="--force" in sys.argv, ptxas_verbose="--ptxas" in sys.argv)
    print(out)
