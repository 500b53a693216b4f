"""ctypes bindings of the two product libraries.

* ``Engine``   -- libmjx.so, the kernel-level C-ABI (include/mjx.h): compiled dropons (K1),
  masked blend (K2) and effects (K3) on device-resident or host planes.
* ``ModJpeg``  -- libmodjpeg.so, the reference's public API (include/libmodjpeg.h) mirrored in
  Python with the reference's names (``read_jpeg_from_memory``, ``compose``, ``effect_tint`` ...).

The libraries are built in-tree by ``libmodjpeg_b200.build``; importing this module never
builds and never falls back to a CPU path: a missing library or a missing GPU raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIBDIR = os.path.join(PKG, "lib")
LIBMJX = os.path.join(LIBDIR, "libmjx.so")
LIBMODJPEG = os.path.join(LIBDIR, "libmodjpeg.so")

MAX_COMPONENTS = 4
OK, ERR_MEMORY, ERR_ARG, ERR_UNSUPPORTED, ERR_DEVICE = 0, 1, 2, 6, 10
CS_RGB, CS_RGBA, CS_GRAYSCALE, CS_GRAYSCALEA, CS_YCC, CS_YCCA = 1, 2, 3, 4, 5, 6
ALIGN_LEFT, ALIGN_RIGHT, ALIGN_TOP, ALIGN_BOTTOM, ALIGN_CENTER = 1, 2, 4, 8, 16
JCS_GRAYSCALE, JCS_RGB, JCS_YCbCr = 1, 2, 3
CLS_T, CLS_U, CLS_OPAQUE, CLS_G = 0, 1, 2, 3
FX_ZERO, FX_PIXELATE, FX_ADD_DC = 1, 2, 3
OPTION_NONE, OPTION_OPTIMIZE, OPTION_PROGRESSIVE, OPTION_ARITHMETRIC = 0, 1, 2, 4


class MjxError(RuntimeError):
    def __init__(self, code: int, what: str):
        super().__init__(f"{what}: error {code}")
        self.code = code


class Layout(C.Structure):
    _fields_ = [("colorspace", C.c_int), ("ncomp", C.c_int), ("h_samp", C.c_int * 4), ("v_samp", C.c_int * 4)]

    @classmethod
    def make(cls, colorspace: int, samp) -> "Layout":
        L = cls()
        L.colorspace = colorspace
        L.ncomp = len(samp)
        for i, (h, v) in enumerate(samp):
            L.h_samp[i] = h
            L.v_samp[i] = v
        return L


class Geometry(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("visible", "crop_x", "crop_y", "crop_w", "crop_h",
                                        "blockoffset_x", "blockoffset_y", "block_x", "block_y")]

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_}


class ImageDesc(C.Structure):
    """mjx_image_desc_t -- one image of a device-resident batch (608 bytes)."""
    _fields_ = [("plane", C.c_uint64 * 4), ("stride_blocks", C.c_int32 * 4), ("rows", C.c_int32 * 4),
                ("wreal", C.c_int32 * 4), ("hreal", C.c_int32 * 4), ("q", (C.c_uint16 * 64) * 4)]


class HostImage(C.Structure):
    _fields_ = [("plane", C.c_void_p * 4), ("stride_blocks", C.c_int32 * 4), ("rows", C.c_int32 * 4),
                ("wreal", C.c_int32 * 4), ("hreal", C.c_int32 * 4), ("q", C.c_void_p * 4)]


class EffectOp(C.Structure):
    _fields_ = [("op", C.c_int), ("comp", C.c_int), ("value", C.c_int)]


IMAGE_DESC_DTYPE = np.dtype([("plane", "<u8", 4), ("stride_blocks", "<i4", 4), ("rows", "<i4", 4),
                             ("wreal", "<i4", 4), ("hreal", "<i4", 4), ("q", "<u2", (4, 64))])
assert IMAGE_DESC_DTYPE.itemsize == C.sizeof(ImageDesc) == 608

_lib_mjx = None
_lib_mj = None


def load_mjx() -> C.CDLL:
    """Load libmjx.so.  Raises if it has not been built -- there is no fallback."""
    global _lib_mjx
    if _lib_mjx is not None:
        return _lib_mjx
    if not os.path.exists(LIBMJX):
        raise FileNotFoundError(f"{LIBMJX} is missing: run `python -m libmodjpeg_b200.build` (no CPU fallback exists)")
    L = C.CDLL(LIBMJX, mode=C.RTLD_GLOBAL)
    vp, ip = C.c_void_p, C.POINTER(C.c_int)
    L.mjx_device_count.restype = C.c_int
    L.mjx_ctx_create.argtypes = [C.POINTER(vp), C.c_int]
    L.mjx_ctx_destroy.argtypes = [vp]
    L.mjx_ctx_destroy.restype = None
    L.mjx_ctx_set_stream.argtypes = [vp, vp]
    L.mjx_ctx_use_own_stream.argtypes = [vp]
    L.mjx_ctx_set_strict.argtypes = [vp, C.c_int]
    L.mjx_ctx_set_tensor_core.argtypes = [vp, C.c_int]
    L.mjx_ctx_set_tensor_core_min_images.argtypes = [vp, C.c_int]
    L.mjx_ctx_set_operator_pieces.argtypes = [vp, C.c_int]
    L.mjx_ctx_stream.argtypes = [vp]
    L.mjx_ctx_stream.restype = vp
    L.mjx_ctx_sync.argtypes = [vp]
    L.mjx_ctx_last_error.argtypes = [vp]
    L.mjx_ctx_last_error.restype = C.c_char_p
    L.mjx_ctx_kernel_launches.argtypes = [vp]
    L.mjx_ctx_kernel_launches.restype = C.c_longlong
    L.mjx_device_alloc.argtypes = [vp, C.POINTER(vp), C.c_size_t]
    L.mjx_device_free.argtypes = [vp, vp]
    L.mjx_device_free.restype = None
    L.mjx_host_alloc.argtypes = [vp, C.POINTER(vp), C.c_size_t]
    L.mjx_host_free.argtypes = [vp, vp]
    L.mjx_host_free.restype = None
    L.mjx_copy_h2d.argtypes = [vp, vp, vp, C.c_size_t]
    L.mjx_copy_d2h.argtypes = [vp, vp, vp, C.c_size_t]
    L.mjx_geometry.argtypes = [C.c_int] * 6 + [C.c_uint, C.c_int, C.c_int, C.POINTER(Geometry)]
    L.mjx_geometry.restype = None
    L.mjx_dropon_compile.argtypes = [vp, C.POINTER(vp), vp, vp, C.c_int, C.c_int, C.c_int, C.POINTER(Layout)] + \
        [C.c_int] * 6 + [C.c_int]
    L.mjx_dropon_from_coefficients.argtypes = [vp, C.POINTER(vp), C.POINTER(Layout), ip, ip, C.POINTER(vp), C.POINTER(vp)]
    L.mjx_dropon_free.argtypes = [vp]
    L.mjx_dropon_free.restype = None
    L.mjx_dropon_ncomp.argtypes = [vp]
    L.mjx_dropon_dims.argtypes = [vp, C.c_int, ip, ip]
    L.mjx_dropon_blocks.argtypes = [vp]
    L.mjx_dropon_blocks.restype = C.c_longlong
    L.mjx_dropon_download.argtypes = [vp, vp, C.c_int, vp, vp, vp]
    L.mjx_dropon_class_counts.argtypes = [vp, vp, C.POINTER(C.c_longlong)]
    L.mjx_dropon_download_generic.argtypes = [vp, vp, vp, vp, vp]
    L.mjx_dropon_generic_slots.argtypes = [vp]
    L.mjx_ctx_set_zero_copy.argtypes = [vp, C.c_int]
    L.mjx_ctx_set_class_mask.argtypes = [vp, C.c_int]
    L.mjx_ctx_set_overlap.argtypes = [vp, C.c_int]
    L.mjx_selftest_reciprocal.argtypes = [vp, C.POINTER(C.c_longlong)]
    L.mjx_compose_batch_device.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.c_int]
    L.mjx_compose_batch_host.argtypes = [vp, C.POINTER(HostImage), C.c_int, vp, C.c_int, C.c_int]
    L.mjx_compose_rows_host.argtypes = [vp, C.c_int, vp, vp, vp]
    L.mjx_huffman_decode_batch_device.argtypes = [vp, vp, vp, vp, C.c_int, vp, vp, vp]
    L.mjx_huffman_encode_batch_device.argtypes = [vp, vp, C.c_int, vp, vp, C.c_size_t, vp]
    L.mjx_effects_batch_device.argtypes = [vp, vp, C.c_int, C.c_int, C.POINTER(EffectOp), C.c_int]
    L.mjx_effects_rows_host.argtypes = [vp, C.c_int, vp, ip, ip, vp, C.POINTER(EffectOp), C.c_int]
    _lib_mjx = L
    return L


def geometry(image_w, image_h, h_factor, v_factor, dropon_w, dropon_h, align, offset_x=0, offset_y=0) -> dict:
    """Placement arithmetic of mj_compose (reference: src/compose.c:42-172); host only, no GPU needed."""
    g = Geometry()
    load_mjx().mjx_geometry(image_w, image_h, h_factor, v_factor, dropon_w, dropon_h, align, offset_x, offset_y, C.byref(g))
    return g.as_dict()


class CompiledDropon:
    """A compiled dropon resident in HBM (mjx_dropon)."""

    def __init__(self, engine: "Engine", handle: int):
        self.engine = engine
        self.handle = C.c_void_p(handle)

    @property
    def ncomp(self) -> int:
        return self.engine.lib.mjx_dropon_ncomp(self.handle)

    def dims(self, c: int) -> tuple[int, int]:
        wb, hb = C.c_int(), C.c_int()
        self.engine._check(self.engine.lib.mjx_dropon_dims(self.handle, c, C.byref(wb), C.byref(hb)), "mjx_dropon_dims")
        return wb.value, hb.value

    @property
    def blocks(self) -> int:
        return self.engine.lib.mjx_dropon_blocks(self.handle)

    def class_counts(self) -> dict:
        a = (C.c_longlong * 4)()
        self.engine._check(self.engine.lib.mjx_dropon_class_counts(self.engine.ctx, self.handle, a), "mjx_dropon_class_counts")
        return {"T": a[0], "U": a[1], "OPAQUE": a[2], "G": a[3]}

    def download(self, c: int):
        wb, hb = self.dims(c)
        D = np.zeros((hb, wb, 64), np.int16)
        W = np.zeros((hb, wb, 64), np.int16)
        cls = np.zeros((hb, wb), np.uint8)
        self.engine._check(self.engine.lib.mjx_dropon_download(self.engine.ctx, self.handle, c, D.ctypes.data, W.ctypes.data,
                                                               cls.ctypes.data), "mjx_dropon_download")
        return D, W, cls

    def download_generic(self):
        """-> (entries uint32 [n], Ds float32 [n][64], A float32 [n][64]) of the generic-class list"""
        n = int(self.engine.lib.mjx_dropon_generic_slots(self.handle))
        lst = np.zeros(n, np.uint32)
        Ds = np.zeros((n, 64), np.float32)
        A = np.zeros((n, 64), np.float32)
        self.engine._check(self.engine.lib.mjx_dropon_download_generic(self.engine.ctx, self.handle, lst.ctypes.data, Ds.ctypes.data,
                                                                       A.ctypes.data), "mjx_dropon_download_generic")
        return lst, Ds, A

    def free(self) -> None:
        if self.handle:
            self.engine.lib.mjx_dropon_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Engine:
    """One mjx_ctx: a device, a stream and staging pools.  Not shareable between threads."""

    def __init__(self, device: int = 0):
        self.lib = load_mjx()
        ctx = C.c_void_p()
        rv = self.lib.mjx_ctx_create(C.byref(ctx), device)
        if rv != OK:
            raise MjxError(rv, f"mjx_ctx_create(device={device}): no usable CUDA device "
                               f"({self.lib.mjx_device_count()} visible); the engine has no CPU fallback")
        self.ctx = ctx
        self.device = device
        env = os.environ.get("MJX_K2_TC")
        self._tc_mode = 1 if env is None else max(0, min(2, int(env)))  # the ctx's initial mode (mjx_ctx_create)

    def close(self) -> None:
        if getattr(self, "ctx", None):
            self.lib.mjx_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rv: int, what: str) -> None:
        if rv != OK:
            msg = self.lib.mjx_ctx_last_error(self.ctx)
            raise MjxError(rv, f"{what} [{msg.decode() if msg else ''}]")

    # ---- context ----------------------------------------------------------------------
    def set_stream(self, cuda_stream: int | None) -> None:
        """Run on a caller-owned cudaStream_t (e.g. torch's); None goes back to the ctx's own stream."""
        if cuda_stream is None:
            self._check(self.lib.mjx_ctx_use_own_stream(self.ctx), "mjx_ctx_use_own_stream")
        else:
            self._check(self.lib.mjx_ctx_set_stream(self.ctx, C.c_void_p(cuda_stream)), "mjx_ctx_set_stream")

    def set_zero_copy(self, on: bool) -> None:
        """on (default): compose_batch_host runs K2 directly on page-locked host planes; off: always stage"""
        self._check(self.lib.mjx_ctx_set_zero_copy(self.ctx, 1 if on else 0), "mjx_ctx_set_zero_copy")

    def selftest_reciprocal(self) -> int:
        """exhaustive device check of K2's reciprocal tables; returns the number of mismatching (a, q) pairs"""
        n = C.c_longlong(-1)
        self._check(self.lib.mjx_selftest_reciprocal(self.ctx, C.byref(n)), "mjx_selftest_reciprocal")
        return n.value

    def set_class_mask(self, mask: int) -> None:
        """profiling aid: bit 0 = OPAQUE/U kernel, bit 1 = G kernel of the fast K2 path (default 3)"""
        self._check(self.lib.mjx_ctx_set_class_mask(self.ctx, mask), "mjx_ctx_set_class_mask")

    def set_overlap(self, on: bool) -> None:
        """large batches: OPAQUE/U kernel beside the G kernel (default) or one after the other"""
        self._check(self.lib.mjx_ctx_set_overlap(self.ctx, 1 if on else 0), "mjx_ctx_set_overlap")

    def set_tensor_core(self, mode: int) -> None:
        """G class of batches of >= 256 images: 1 tensor-core kernel with coefficient range check (default), 2 without
        check, 0 fp32 kernel"""
        self._check(self.lib.mjx_ctx_set_tensor_core(self.ctx, mode), "mjx_ctx_set_tensor_core")
        self._tc_mode = mode

    def set_tensor_core_min_images(self, n: int) -> None:
        """smallest batch that takes the tensor-core G kernel (>= 256; default 1025)"""
        self._check(self.lib.mjx_ctx_set_tensor_core_min_images(self.ctx, n), "mjx_ctx_set_tensor_core_min_images")
        self._tc_min = n

    def tensor_core_active(self, n_images: int, generic_blocks: int) -> bool:
        """does a batch of n_images take the tensor-core G kernel (mode set, batch large enough, class G present)?"""
        return self._tc_mode != 0 and n_images >= getattr(self, "_tc_min", 1025) and generic_blocks > 0

    def set_operator_pieces(self, pieces: int) -> None:
        """fp16 pieces per entry of the tensor-core kernel's operator (2 or 3); for dropons not yet used in a large batch"""
        self._check(self.lib.mjx_ctx_set_operator_pieces(self.ctx, pieces), "mjx_ctx_set_operator_pieces")

    def set_strict(self, strict: bool) -> None:
        """strict: one K2 kernel with the reference's int16 wrap-around (adversarial inputs); default fast kernels"""
        self._check(self.lib.mjx_ctx_set_strict(self.ctx, 1 if strict else 0), "mjx_ctx_set_strict")

    @property
    def stream(self) -> int:
        return self.lib.mjx_ctx_stream(self.ctx) or 0

    def sync(self) -> None:
        self._check(self.lib.mjx_ctx_sync(self.ctx), "mjx_ctx_sync")

    @property
    def kernel_launches(self) -> int:
        return self.lib.mjx_ctx_kernel_launches(self.ctx)

    def device_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._check(self.lib.mjx_device_alloc(self.ctx, C.byref(p), nbytes), "mjx_device_alloc")
        return p.value

    def device_free(self, ptr: int) -> None:
        self.lib.mjx_device_free(self.ctx, C.c_void_p(ptr))

    def host_alloc(self, nbytes: int, dtype=np.uint8) -> np.ndarray:
        """Page-locked host memory as a numpy array (freed with host_free)."""
        p = C.c_void_p()
        self._check(self.lib.mjx_host_alloc(self.ctx, C.byref(p), nbytes), "mjx_host_alloc")
        buf = (C.c_uint8 * nbytes).from_address(p.value)
        a = np.frombuffer(buf, dtype=np.uint8).view(dtype)
        return a

    def host_free(self, a: np.ndarray) -> None:
        self.lib.mjx_host_free(self.ctx, C.c_void_p(a.ctypes.data))

    def copy_h2d(self, dst_dev: int, src: np.ndarray) -> None:
        self._check(self.lib.mjx_copy_h2d(self.ctx, C.c_void_p(dst_dev), C.c_void_p(src.ctypes.data), src.nbytes), "mjx_copy_h2d")

    def copy_d2h(self, dst: np.ndarray, src_dev: int) -> None:
        self._check(self.lib.mjx_copy_d2h(self.ctx, C.c_void_p(dst.ctypes.data), C.c_void_p(src_dev), dst.nbytes), "mjx_copy_d2h")

    # ---- K1 -----------------------------------------------------------------------------
    def dropon_compile(self, image3, alpha3, dropon_cs: int, layout: Layout, blockoffset=(0, 0), crop=None,
                       device_pixels: tuple[int, int, int, int] | None = None) -> CompiledDropon:
        """mj_compile_dropon on the GPU.  image3/alpha3: uint8 [h][w][3] host arrays, or pass
        device_pixels=(image_ptr, alpha_ptr, width, height) for pixels already in HBM."""
        if device_pixels is not None:
            ip, ap, w, h = device_pixels
            on_dev = 1
        else:
            image3 = np.ascontiguousarray(image3, np.uint8)
            alpha3 = np.ascontiguousarray(alpha3, np.uint8)
            h, w = image3.shape[:2]
            ip, ap, on_dev = image3.ctypes.data, alpha3.ctypes.data, 0
        cx, cy, cw, ch = crop if crop is not None else (0, 0, w, h)
        out = C.c_void_p()
        rv = self.lib.mjx_dropon_compile(self.ctx, C.byref(out), C.c_void_p(ip), C.c_void_p(ap), w, h, dropon_cs,
                                         C.byref(layout), blockoffset[0], blockoffset[1], cx, cy, cw, ch, on_dev)
        self._check(rv, "mjx_dropon_compile")
        return CompiledDropon(self, out.value)

    def dropon_from_coefficients(self, layout: Layout, D: list[np.ndarray], W: list[np.ndarray]) -> CompiledDropon:
        n = layout.ncomp
        D = [np.ascontiguousarray(a, np.int16) for a in D]
        W = [np.ascontiguousarray(a, np.int16) for a in W]
        wb = (C.c_int * 4)(*[a.shape[1] for a in D] + [0] * (4 - n))
        hb = (C.c_int * 4)(*[a.shape[0] for a in D] + [0] * (4 - n))
        Dp = (C.c_void_p * 4)(*[a.ctypes.data for a in D] + [None] * (4 - n))
        Wp = (C.c_void_p * 4)(*[a.ctypes.data for a in W] + [None] * (4 - n))
        out = C.c_void_p()
        self._check(self.lib.mjx_dropon_from_coefficients(self.ctx, C.byref(out), C.byref(layout), wb, hb, Dp, Wp),
                    "mjx_dropon_from_coefficients")
        return CompiledDropon(self, out.value)

    # ---- K2 -----------------------------------------------------------------------------
    def compose_batch_device(self, items_dev: int, n: int, dropon: CompiledDropon, block_x: int, block_y: int) -> None:
        self._check(self.lib.mjx_compose_batch_device(self.ctx, C.c_void_p(items_dev), n, dropon.handle, block_x, block_y),
                    "mjx_compose_batch_device")

    def compose_batch_host(self, items, n: int, dropon: CompiledDropon, block_x: int, block_y: int) -> None:
        self._check(self.lib.mjx_compose_batch_host(self.ctx, items, n, dropon.handle, block_x, block_y),
                    "mjx_compose_batch_host")

    def compose_planes_host(self, planes: list[np.ndarray], qtables: list[np.ndarray], dropon: CompiledDropon,
                            block_x: int, block_y: int) -> None:
        """One image given as flat host planes [rows][cols][64] int16, blended in place."""
        item, keep = make_host_image(planes, qtables)
        arr = (HostImage * 1)(item)
        self.compose_batch_host(arr, 1, dropon, block_x, block_y)
        del keep

    # ---- K3 -----------------------------------------------------------------------------
    def effects_batch_device(self, items_dev: int, n: int, ncomp: int, ops: list[tuple[int, int, int]]) -> None:
        arr = (EffectOp * max(1, len(ops)))(*[EffectOp(*o) for o in ops])
        self._check(self.lib.mjx_effects_batch_device(self.ctx, C.c_void_p(items_dev), n, ncomp, arr, len(ops)),
                    "mjx_effects_batch_device")

    # ---- K4 -----------------------------------------------------------------------------
    def huffman_encode_batch_device(self, items_dev: int, n: int, scan: "Scan", out_dev: int, out_stride: int, sizes_dev: int) -> None:
        """entropy-coded segments of n device-resident images (asynchronous on the ctx stream)"""
        self._check(self.lib.mjx_huffman_encode_batch_device(self.ctx, C.c_void_p(items_dev), n, C.byref(scan), C.c_void_p(out_dev),
                                                             C.c_size_t(out_stride), C.c_void_p(sizes_dev)), "mjx_huffman_encode_batch_device")

    # ---- K5 -----------------------------------------------------------------------------
    def huffman_decode_batch_device(self, data_dev: int, offsets: np.ndarray, lengths: np.ndarray, n: int, scan: "Scan", items_dev: int,
                                    status_dev: int) -> None:
        """entropy-coded segments (device memory, offsets / lengths on the host) -> coefficient planes of n device-resident
        images (asynchronous on the ctx stream)"""
        offsets = np.ascontiguousarray(offsets, np.uint64)
        lengths = np.ascontiguousarray(lengths, np.uint32)
        assert offsets.size >= n and lengths.size >= n
        self._check(self.lib.mjx_huffman_decode_batch_device(self.ctx, C.c_void_p(data_dev), C.c_void_p(offsets.ctypes.data), C.c_void_p(lengths.ctypes.data), n,
                                                             C.byref(scan), C.c_void_p(items_dev), C.c_void_p(status_dev)), "mjx_huffman_decode_batch_device")


class HuffTable(C.Structure):
    _fields_ = [("bits", C.c_uint8 * 17), ("vals", C.c_uint8 * 256)]


class Scan(C.Structure):
    """mjx_scan_t"""
    _fields_ = [("ncomp", C.c_int32), ("h_samp", C.c_int32 * MAX_COMPONENTS), ("v_samp", C.c_int32 * MAX_COMPONENTS),
                ("dc_tbl", C.c_int32 * MAX_COMPONENTS), ("ac_tbl", C.c_int32 * MAX_COMPONENTS), ("mcus_per_row", C.c_int32),
                ("mcu_rows", C.c_int32), ("dc", HuffTable * 4), ("ac", HuffTable * 4)]


# The typical Huffman tables of ITU-T T.81 Annex K.3 (tables K.3 - K.6): what libjpeg installs for a compressor that is not
# asked to optimise (jpeg_set_defaults -> std_huff_tables), i.e. what the reference's mj_write_jpeg_to_memory writes.
STD_DC_LUMA = ([0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0], list(range(12)))
STD_DC_CHROMA = ([0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0], list(range(12)))
STD_AC_LUMA = ([0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d],
               [0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32,
                0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16,
                0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45,
                0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69,
                0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94,
                0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6,
                0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8,
                0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8,
                0xf9, 0xfa])
STD_AC_CHROMA = ([0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77],
                 [0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81,
                  0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34,
                  0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44,
                  0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68,
                  0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92,
                  0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4,
                  0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6,
                  0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8,
                  0xf9, 0xfa])


def _fill_table(t: HuffTable, spec) -> None:
    counts, vals = spec
    for k, cnt in enumerate(counts):
        t.bits[k + 1] = cnt
    for k, v in enumerate(vals):
        t.vals[k] = v


def standard_scan(width: int, height: int, samp: list[tuple[int, int]], chroma_from: int = 1) -> Scan:
    """the scan libjpeg writes for an image of this size and sampling with its default (Annex K) tables: component 0 uses
    tables 0 (luminance), components from `chroma_from` on tables 1 (jcparam.c jpeg_set_colorspace, YCbCr / grayscale)"""
    s = Scan()
    s.ncomp = len(samp)
    max_h, max_v = max(h for h, _ in samp), max(v for _, v in samp)
    for c, (h, v) in enumerate(samp):
        s.h_samp[c], s.v_samp[c] = h, v
        s.dc_tbl[c] = s.ac_tbl[c] = 1 if c >= chroma_from else 0
    if len(samp) == 1:  # not interleaved: the component's own block grid
        s.mcus_per_row, s.mcu_rows = (width + 7) // 8, (height + 7) // 8
    else:
        s.mcus_per_row, s.mcu_rows = -(-width // (8 * max_h)), -(-height // (8 * max_v))
    _fill_table(s.dc[0], STD_DC_LUMA)
    _fill_table(s.ac[0], STD_AC_LUMA)
    _fill_table(s.dc[1], STD_DC_CHROMA)
    _fill_table(s.ac[1], STD_AC_CHROMA)
    return s


def scan_from_jpeg(data: bytes) -> tuple[Scan, int, dict]:
    """Parse the markers of a baseline JPEG up to its first SOS: returns (the scan as K4 / K5 want it -- the FILE's Huffman
    tables, sampling factors, table selectors, MCU grid --, offset of the entropy-coded segment, frame info).  Raises
    ValueError for what the device decoder does not take: not SOF0/SOF1 8-bit, restart intervals, a scan that does not hold
    every component."""
    if data[:2] != b"\xff\xd8":
        raise ValueError("not a JPEG")
    i, frame, tables, dri = 2, None, {}, 0
    while True:
        if data[i] != 0xFF:
            raise ValueError("marker expected")
        while data[i + 1] == 0xFF:
            i += 1
        m = data[i + 1]
        n = (data[i + 2] << 8) | data[i + 3]
        seg = data[i + 4:i + 2 + n]
        if m in (0xC0, 0xC1):
            if seg[0] != 8:
                raise ValueError("not 8-bit")
            frame = {"height": (seg[1] << 8) | seg[2], "width": (seg[3] << 8) | seg[4],
                     "comps": [(seg[6 + 3 * k], seg[7 + 3 * k] >> 4, seg[7 + 3 * k] & 15, seg[8 + 3 * k]) for k in range(seg[5])]}
        elif 0xC2 <= m <= 0xCF and m not in (0xC4, 0xC8, 0xCC):
            raise ValueError("not a sequential Huffman frame")
        elif m == 0xC4:
            k = 0
            while k < len(seg):
                tc, th = seg[k] >> 4, seg[k] & 15
                counts = list(seg[k + 1:k + 17])
                nv = sum(counts)
                tables[(tc, th)] = (counts, list(seg[k + 17:k + 17 + nv]))
                k += 17 + nv
        elif m == 0xDD:
            dri = (seg[0] << 8) | seg[1]
        elif m == 0xDA:
            if frame is None or dri != 0 or seg[0] != len(frame["comps"]):
                raise ValueError("restart intervals or a scan without every component")
            s = Scan()
            s.ncomp = seg[0]
            ids = [c[0] for c in frame["comps"]]
            for k in range(seg[0]):
                cid, sel = seg[1 + 2 * k], seg[2 + 2 * k]
                if cid != ids[k]:
                    raise ValueError("scan components out of frame order")
                s.h_samp[k], s.v_samp[k] = frame["comps"][k][1], frame["comps"][k][2]
                s.dc_tbl[k], s.ac_tbl[k] = sel >> 4, sel & 15
            max_h, max_v = max(c[1] for c in frame["comps"]), max(c[2] for c in frame["comps"])
            if s.ncomp == 1:
                h0, v0 = frame["comps"][0][1], frame["comps"][0][2]
                s.mcus_per_row = -(-(frame["width"] * h0) // (8 * max_h))
                s.mcu_rows = -(-(frame["height"] * v0) // (8 * max_v))
            else:
                s.mcus_per_row, s.mcu_rows = -(-frame["width"] // (8 * max_h)), -(-frame["height"] // (8 * max_v))
            for (tc, th), spec in tables.items():
                if th < 4:
                    _fill_table(s.dc[th] if tc == 0 else s.ac[th], spec)
            return s, i + 2 + n, frame
        i += 2 + n


def make_host_image(planes: list[np.ndarray], qtables: list[np.ndarray], real_dims=None):
    """HostImage over flat numpy planes; returns (struct, keepalive)."""
    it = HostImage()
    keep = []
    for c, (p, q) in enumerate(zip(planes, qtables)):
        assert p.dtype == np.int16 and p.flags.c_contiguous and p.ndim == 3 and p.shape[2] == 64
        q = np.ascontiguousarray(q, np.uint16)
        keep += [p, q]
        it.plane[c] = p.ctypes.data
        it.stride_blocks[c] = p.shape[1]
        it.rows[c] = p.shape[0]
        it.wreal[c] = real_dims[c][0] if real_dims else p.shape[1]
        it.hreal[c] = real_dims[c][1] if real_dims else p.shape[0]
        it.q[c] = q.ctypes.data
    return it, keep


def make_image_descs(plane_ptrs, strides, rows, qtables, real_dims=None) -> np.ndarray:
    """Array of mjx_image_desc_t for a device-resident batch.
    plane_ptrs: [n][ncomp] device addresses; strides/rows: [ncomp]; qtables: [n][ncomp][64] or [ncomp][64]."""
    n = len(plane_ptrs)
    ncomp = len(strides)
    a = np.zeros(n, IMAGE_DESC_DTYPE)
    q = np.asarray(qtables, np.uint16)
    for c in range(ncomp):
        a["plane"][:, c] = [p[c] for p in plane_ptrs]
        a["stride_blocks"][:, c] = strides[c]
        a["rows"][:, c] = rows[c]
        a["wreal"][:, c] = real_dims[c][0] if real_dims else strides[c]
        a["hreal"][:, c] = real_dims[c][1] if real_dims else rows[c]
        a["q"][:, c, :] = q[:, c, :] if q.ndim == 3 else q[c]
    return a


# --------------------------------------------------------------------------------------
# the public API (include/libmodjpeg.h) mirrored in Python
# --------------------------------------------------------------------------------------


class _DroponStruct(C.Structure):
    _fields_ = [("image", C.POINTER(C.c_uint8)), ("alpha", C.POINTER(C.c_uint8)), ("width", C.c_int), ("height", C.c_int),
                ("colorspace", C.c_int), ("blend", C.c_int)]


def load_modjpeg() -> C.CDLL:
    global _lib_mj
    if _lib_mj is not None:
        return _lib_mj
    load_mjx()
    if not os.path.exists(LIBMODJPEG):
        raise FileNotFoundError(f"{LIBMODJPEG} is missing: run `python -m libmodjpeg_b200.build`")
    L = C.CDLL(LIBMODJPEG)
    vp = C.c_void_p
    L.mj_init_jpeg.argtypes = [vp]
    L.mj_init_jpeg.restype = None
    L.mj_free_jpeg.argtypes = [vp]
    L.mj_free_jpeg.restype = None
    L.mj_read_jpeg_from_memory.argtypes = [vp, C.c_char_p, C.c_size_t, C.c_size_t]
    L.mj_read_jpeg_from_file.argtypes = [vp, C.c_char_p, C.c_size_t]
    L.mj_write_jpeg_to_memory.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t), C.c_int]
    L.mjx_write_jpeg_to_memory_device.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t), C.c_int]
    L.mj_write_jpeg_to_file.argtypes = [vp, C.c_char_p, C.c_int]
    L.mj_init_dropon.argtypes = [vp]
    L.mj_init_dropon.restype = None
    L.mj_free_dropon.argtypes = [vp]
    L.mj_free_dropon.restype = None
    L.mj_read_dropon_from_raw.argtypes = [vp, C.c_char_p, C.c_uint, C.c_int, C.c_int, C.c_short]
    L.mj_read_dropon_from_memory.argtypes = [vp, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_short]
    L.mj_read_dropon_from_file.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_short]
    L.mj_compose.argtypes = [vp, vp, C.c_uint, C.c_int, C.c_int]
    L.mj_effect_grayscale.argtypes = [vp]
    L.mj_effect_pixelate.argtypes = [vp]
    L.mj_effect_tint.argtypes = [vp, C.c_int, C.c_int]
    L.mj_effect_luminance.argtypes = [vp, C.c_int]
    L.mjx_jpeg_image_info.argtypes = [vp, C.POINTER(C.c_int)]
    L.mjx_jpeg_component_info.argtypes = [vp, C.c_int, C.POINTER(C.c_int)]
    L.mjx_jpeg_qtable.argtypes = [vp, C.c_int, vp]
    L.mjx_jpeg_export_plane.argtypes = [vp, C.c_int, vp]
    L.mjx_jpeg_import_plane.argtypes = [vp, C.c_int, vp]
    L.mjx_jpeg_layout.argtypes = [vp, C.POINTER(Layout)]
    L.mj_batch_set_devices.argtypes = [C.c_int]
    L.mj_batch_set_devices.restype = None
    L.mj_coalesce_configure.argtypes = [C.c_int, C.c_int, C.c_int]
    L.mj_coalesce_configure.restype = None
    L.mj_coalesce_stats.argtypes = [C.POINTER(C.c_ulong), C.POINTER(C.c_ulong)]
    L.mj_coalesce_stats.restype = None
    L.mj_compose_batch.argtypes = [C.c_int, C.POINTER(Blob), C.POINTER(Blob), C.POINTER(C.c_int), vp, C.c_uint, C.c_int, C.c_int, C.c_int, C.c_int]
    L.mjx_host_ctx.argtypes = []
    L.mjx_host_ctx.restype = vp
    _lib_mj = L
    return L


_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]
_libc.free.restype = None


class Jpeg:
    """mj_jpeg_t: a decoded JPEG (libjpeg state + coefficient arrays)."""

    SIZE = 1024  # >= sizeof(mj_jpeg_t) == 696 with the ABI-62 libjpeg

    def __init__(self):
        self.lib = load_modjpeg()
        self.buf = C.create_string_buffer(self.SIZE)
        self.ptr = C.cast(self.buf, C.c_void_p)
        self.lib.mj_init_jpeg(self.ptr)

    # -- reading / writing (host libjpeg) --
    def read_jpeg_from_memory(self, data: bytes, max_pixel: int = 0) -> int:
        return self.lib.mj_read_jpeg_from_memory(self.ptr, data, len(data), max_pixel)

    def read_jpeg_from_file(self, path: str, max_pixel: int = 0) -> int:
        return self.lib.mj_read_jpeg_from_file(self.ptr, path.encode(), max_pixel)

    def write_jpeg_to_memory(self, options: int = 0) -> tuple[int, bytes]:
        mem, n = C.c_void_p(), C.c_size_t()
        rv = self.lib.mj_write_jpeg_to_memory(self.ptr, C.byref(mem), C.byref(n), options)
        if rv != OK:
            return rv, b""
        out = C.string_at(mem, n.value)
        _libc.free(mem)
        return rv, out

    def write_jpeg_to_memory_device(self, options: int = 0) -> tuple[int, bytes]:
        """the same file with the scan Huffman-coded on the GPU (K4); baseline only"""
        mem, n = C.c_void_p(), C.c_size_t()
        rv = self.lib.mjx_write_jpeg_to_memory_device(self.ptr, C.byref(mem), C.byref(n), options)
        if rv != OK:
            return rv, b""
        out = C.string_at(mem, n.value)
        _libc.free(mem)
        return rv, out

    def write_jpeg_to_file(self, path: str, options: int = 0) -> int:
        return self.lib.mj_write_jpeg_to_file(self.ptr, path.encode(), options)

    # -- the hot path --
    def compose(self, dropon: "Dropon", align: int, offset_x: int = 0, offset_y: int = 0) -> int:
        return self.lib.mj_compose(self.ptr, dropon.ptr if dropon is not None else None, align, offset_x, offset_y)

    def effect_grayscale(self) -> int:
        return self.lib.mj_effect_grayscale(self.ptr)

    def effect_pixelate(self) -> int:
        return self.lib.mj_effect_pixelate(self.ptr)

    def effect_tint(self, cb_value: int, cr_value: int) -> int:
        return self.lib.mj_effect_tint(self.ptr, cb_value, cr_value)

    def effect_luminance(self, value: int) -> int:
        return self.lib.mj_effect_luminance(self.ptr, value)

    # -- plane access (mjx_host.h) --
    def info(self) -> dict:
        a = (C.c_int * 8)()
        if self.lib.mjx_jpeg_image_info(self.ptr, a) != OK:
            raise RuntimeError("no image loaded")
        return dict(ncomp=a[0], colorspace=a[1], width=a[2], height=a[3], max_h=a[4], max_v=a[5])

    def comp_info(self, c: int) -> dict:
        a = (C.c_int * 8)()
        if self.lib.mjx_jpeg_component_info(self.ptr, c, a) != OK:
            raise RuntimeError("bad component")
        return dict(wreal=a[0], hreal=a[1], h=a[2], v=a[3], wvirt=a[4], hvirt=a[5])

    def sampling(self) -> list[tuple[int, int]]:
        return [(self.comp_info(c)["h"], self.comp_info(c)["v"]) for c in range(self.info()["ncomp"])]

    def layout(self) -> Layout:
        L = Layout()
        if self.lib.mjx_jpeg_layout(self.ptr, C.byref(L)) != OK:
            raise RuntimeError("no image loaded")
        return L

    def qtable(self, c: int) -> np.ndarray:
        q = np.zeros(64, np.uint16)
        assert self.lib.mjx_jpeg_qtable(self.ptr, c, q.ctypes.data) == OK
        return q

    def plane(self, c: int) -> np.ndarray:
        ci = self.comp_info(c)
        a = np.zeros((ci["hvirt"], ci["wvirt"], 64), np.int16)
        assert self.lib.mjx_jpeg_export_plane(self.ptr, c, a.ctypes.data) == OK
        return a

    def planes(self) -> list[np.ndarray]:
        return [self.plane(c) for c in range(self.info()["ncomp"])]

    def set_plane(self, c: int, a: np.ndarray) -> None:
        a = np.ascontiguousarray(a, np.int16)
        ci = self.comp_info(c)
        assert a.shape == (ci["hvirt"], ci["wvirt"], 64)
        assert self.lib.mjx_jpeg_import_plane(self.ptr, c, a.ctypes.data) == OK

    def free(self) -> None:
        if self.buf is not None:
            self.lib.mj_free_jpeg(self.ptr)
            self.buf = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Blob(C.Structure):
    """mj_blob_t (include/libmodjpeg.h)"""
    _fields_ = [("data", C.c_void_p), ("len", C.c_size_t)]


def batch_set_devices(devices: int) -> None:
    """GPUs mj_compose_batch spreads a batch over (0: $MJX_DEVICES, default 1)"""
    load_modjpeg().mj_batch_set_devices(devices)


def coalesce_configure(enable: bool, max_batch: int = 0, wait_us: int = -1) -> None:
    """request coalescer of mj_compose (libmodjpeg_b200/csrc/host/mj_coalesce.c): concurrent calls with one dropon share a launch"""
    load_modjpeg().mj_coalesce_configure(1 if enable else 0, max_batch, wait_us)


def coalesce_stats() -> tuple[int, int]:
    """(launches made, requests served) by the coalescer so far"""
    b, r = C.c_ulong(0), C.c_ulong(0)
    load_modjpeg().mj_coalesce_stats(C.byref(b), C.byref(r))
    return b.value, r.value


def compose_batch(jpegs: list[bytes], dropon: "Dropon", align: int, offset_x: int = 0, offset_y: int = 0, options: int = 0,
                  nthreads: int = 1, timing: dict | None = None) -> tuple[int, list[int], list[bytes | None]]:
    """mj_compose_batch: decode -> compose -> encode for many JPEGs and one dropon.  Returns (rv, status[], outputs[]);
    timing["call_s"] receives the wall time of the C call alone (without this wrapper's copies into Python bytes)."""
    import time

    L = load_modjpeg()
    n = len(jpegs)
    cin = (Blob * max(n, 1))(*[Blob(C.cast(C.c_char_p(j), C.c_void_p).value, len(j)) for j in jpegs])  # no copy: points into the bytes
    cout = (Blob * max(n, 1))()
    status = (C.c_int * max(n, 1))()
    t0 = time.perf_counter()
    rv = L.mj_compose_batch(n, cin, cout, status, dropon.ptr, align, offset_x, offset_y, options, nthreads)
    if timing is not None:
        timing["call_s"] = time.perf_counter() - t0
    outs: list[bytes | None] = []
    for i in range(n):
        if cout[i].data:
            outs.append(C.string_at(cout[i].data, cout[i].len))
            _libc.free(cout[i].data)
        else:
            outs.append(None)
    return rv, list(status[:n]), outs


class Dropon:
    """mj_dropon_t: an overlay and its alpha mask."""

    def __init__(self):
        self.lib = load_modjpeg()
        self.struct = _DroponStruct()
        self.ptr = C.cast(C.pointer(self.struct), C.c_void_p)
        self.lib.mj_init_dropon(self.ptr)

    def read_dropon_from_raw(self, raw: np.ndarray, colorspace: int, blend: int = 255) -> int:
        raw = np.ascontiguousarray(raw, np.uint8)
        h, w = raw.shape[:2]
        return self.lib.mj_read_dropon_from_raw(self.ptr, raw.tobytes(), colorspace, w, h, blend)

    def read_dropon_from_memory(self, data: bytes, mask: bytes | None = None, blend: int = 255) -> int:
        return self.lib.mj_read_dropon_from_memory(self.ptr, data, len(data), mask, len(mask) if mask else 0, blend)

    def read_dropon_from_file(self, path: str, mask_path: str | None = None, blend: int = 255) -> int:
        return self.lib.mj_read_dropon_from_file(self.ptr, path.encode(), mask_path.encode() if mask_path else None, blend)

    width = property(lambda self: self.struct.width)
    height = property(lambda self: self.struct.height)
    colorspace = property(lambda self: self.struct.colorspace)
    blend = property(lambda self: self.struct.blend)

    def image3(self) -> np.ndarray:
        n = 3 * self.width * self.height
        return np.ctypeslib.as_array(self.struct.image, (n,)).reshape(self.height, self.width, 3).copy()

    def alpha3(self) -> np.ndarray:
        n = 3 * self.width * self.height
        return np.ctypeslib.as_array(self.struct.alpha, (n,)).reshape(self.height, self.width, 3).copy()

    def free(self) -> None:
        if self.struct is not None:
            self.lib.mj_free_dropon(self.ptr)
            self.struct = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
