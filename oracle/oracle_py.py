"""ctypes bindings for the CPU oracle -- TEST INFRASTRUCTURE, never imported by the product.

Two checkers live here:

* ``OraclePort``  -- oracle/_build/libmj_oracle.so, the plain-C restatement (mj_oracle.c).
* ``Reference``   -- oracle/_ref/libmodjpeg_ref.so, the UNMODIFIED reference compiled from
  /root/reference/src by oracle/build_ref.sh (present when that build ran; it travels to the
  GPU box as a prebuilt file).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
REF_SO = os.path.join(HERE, "_ref", "libmodjpeg_ref.so")
PORT_SO = os.path.join(BUILD, "libmj_oracle.so")
HARNESS_SO = os.path.join(BUILD, "libmjharness.so")

# public constants (reference: src/libmodjpeg.h:38-69)
CS_RGB, CS_RGBA, CS_GRAY, CS_GRAYA, CS_YCC, CS_YCCA = 1, 2, 3, 4, 5, 6
ALIGN_LEFT, ALIGN_RIGHT, ALIGN_TOP, ALIGN_BOTTOM, ALIGN_CENTER = 1, 2, 4, 8, 16
JCS_GRAYSCALE, JCS_RGB, JCS_YCbCr = 1, 2, 3


def build(quiet: bool = True) -> None:
    """Compile the oracle port + harness (and oracle/_ref when /root/reference exists)."""
    out = subprocess.run(["make", "-C", HERE], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


def have_reference() -> bool:
    return os.path.exists(REF_SO)


_i16p = C.POINTER(C.c_int16)
_u16p = C.POINTER(C.c_uint16)
_u8p = C.POINTER(C.c_uint8)
_f32p = C.POINTER(C.c_float)


def _ptr(a: np.ndarray, t):
    return a.ctypes.data_as(t)


# --------------------------------------------------------------------------------------
# the plain-C restatement
# --------------------------------------------------------------------------------------


class _Geometry(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("visible", "crop_x", "crop_y", "crop_w", "crop_h",
                                        "blockoffset_x", "blockoffset_y", "block_x", "block_y")]


class _Layout(C.Structure):
    _fields_ = [("colorspace", C.c_int), ("ncomp", C.c_int), ("h", C.c_int * 4), ("v", C.c_int * 4)]


def make_layout(colorspace: int, samp: list[tuple[int, int]]) -> _Layout:
    L = _Layout()
    L.colorspace = colorspace
    L.ncomp = len(samp)
    for i, (h, v) in enumerate(samp):
        L.h[i] = h
        L.v[i] = v
    return L


class OraclePort:
    def __init__(self, path: str = PORT_SO):
        if not os.path.exists(path):
            build()
        self.lib = L = C.CDLL(path)
        L.mjo_geometry.argtypes = [C.c_int] * 6 + [C.c_uint, C.c_int, C.c_int, C.POINTER(_Geometry)]
        L.mjo_geometry.restype = None
        L.mjo_compiled_dims.argtypes = [C.POINTER(_Layout)] + [C.c_int] * 4 + [C.POINTER(C.c_int)] * 2
        L.mjo_compiled_dims.restype = None
        L.mjo_compile_dropon.argtypes = [_u8p, _u8p, C.c_int, C.c_int, C.c_int, C.POINTER(_Layout)] + \
            [C.c_int] * 6 + [C.POINTER(_i16p), C.POINTER(_i16p)]
        L.mjo_compile_dropon.restype = C.c_int
        L.mjo_alpha_weights.argtypes = [_i16p, _f32p]
        L.mjo_alpha_weights.restype = None
        L.mjo_convolve.argtypes = [_f32p, _f32p, C.c_float, C.c_int, C.c_int]
        L.mjo_convolve.restype = None
        L.mjo_compose_block.argtypes = [_i16p, _i16p, _i16p, _u16p]
        L.mjo_compose_block.restype = None
        L.mjo_compose_plane.argtypes = [_i16p, C.c_int, C.c_int, C.c_int, _i16p, _i16p, C.c_int, C.c_int, _u16p]
        L.mjo_compose_plane.restype = None
        L.mjo_effect_zero_plane.argtypes = [_i16p, C.c_int, C.c_int, C.c_int]
        L.mjo_effect_zero_plane.restype = None
        L.mjo_effect_pixelate_plane.argtypes = [_i16p, C.c_int, C.c_int, C.c_int]
        L.mjo_effect_pixelate_plane.restype = None
        L.mjo_effect_add_dc_plane.argtypes = [_i16p, C.c_int, C.c_int, C.c_int, C.c_uint16, C.c_int]
        L.mjo_effect_add_dc_plane.restype = None

    # A1
    def geometry(self, img_w, img_h, h_factor, v_factor, d_w, d_h, align, ox, oy) -> dict:
        g = _Geometry()
        self.lib.mjo_geometry(img_w, img_h, h_factor, v_factor, d_w, d_h, align, ox, oy, C.byref(g))
        return {n: getattr(g, n) for n, _ in _Geometry._fields_}

    def compiled_dims(self, layout: _Layout, boff_x, boff_y, crop_w, crop_h):
        wb = (C.c_int * 4)()
        hb = (C.c_int * 4)()
        self.lib.mjo_compiled_dims(C.byref(layout), boff_x, boff_y, crop_w, crop_h, wb, hb)
        return [(wb[c], hb[c]) for c in range(layout.ncomp)]

    # A2 + A3
    def compile_dropon(self, image3: np.ndarray, alpha3: np.ndarray, dropon_cs: int, layout: _Layout,
                       boff_x=0, boff_y=0, crop=None):
        """-> (rv, D[c], W[c]) with D/W int16 arrays [hb][wb][64]"""
        dh, dw = image3.shape[:2]
        image3 = np.ascontiguousarray(image3, dtype=np.uint8)
        alpha3 = np.ascontiguousarray(alpha3, dtype=np.uint8)
        cx, cy, cw, ch = crop if crop is not None else (0, 0, dw, dh)
        dims = self.compiled_dims(layout, boff_x, boff_y, cw, ch)
        D = [np.zeros((hb, wb, 64), np.int16) for wb, hb in dims]
        W = [np.zeros((hb, wb, 64), np.int16) for wb, hb in dims]
        n = layout.ncomp
        Dp = (_i16p * 4)(*[_ptr(a, _i16p) for a in D] + [None] * (4 - n))
        Wp = (_i16p * 4)(*[_ptr(a, _i16p) for a in W] + [None] * (4 - n))
        rv = self.lib.mjo_compile_dropon(_ptr(image3, _u8p), _ptr(alpha3, _u8p), dw, dh, dropon_cs,
                                         C.byref(layout), boff_x, boff_y, cx, cy, cw, ch, Dp, Wp)
        return rv, D, W

    def alpha_weights(self, Wblock: np.ndarray) -> np.ndarray:
        Wblock = np.ascontiguousarray(Wblock, np.int16)
        w = np.zeros(64, np.float32)
        self.lib.mjo_alpha_weights(_ptr(Wblock, _i16p), _ptr(w, _f32p))
        return w

    def convolve(self, x: np.ndarray, y: np.ndarray, w: float, k: int, l: int) -> None:
        assert x.dtype == np.float32 and y.dtype == np.float32
        self.lib.mjo_convolve(_ptr(x, _f32p), _ptr(y, _f32p), w, k, l)

    # A4
    def compose_plane(self, plane: np.ndarray, x0: int, y0: int, D: np.ndarray, W: np.ndarray, q: np.ndarray) -> None:
        """in place on plane [rows][cols][64] int16"""
        assert plane.dtype == np.int16 and plane.flags.c_contiguous
        D = np.ascontiguousarray(D, np.int16)
        W = np.ascontiguousarray(W, np.int16)
        q = np.ascontiguousarray(q, np.uint16)
        hb, wb = D.shape[:2]
        assert y0 + hb <= plane.shape[0] and x0 + wb <= plane.shape[1]
        self.lib.mjo_compose_plane(_ptr(plane, _i16p), plane.shape[1], x0, y0, _ptr(D, _i16p), _ptr(W, _i16p),
                                   wb, hb, _ptr(q, _u16p))

    # A7-A9
    def effect_zero(self, plane, wreal, hreal):
        self.lib.mjo_effect_zero_plane(_ptr(plane, _i16p), plane.shape[1], wreal, hreal)

    def effect_pixelate(self, plane, wreal, hreal):
        self.lib.mjo_effect_pixelate_plane(_ptr(plane, _i16p), plane.shape[1], wreal, hreal)

    def effect_add_dc(self, plane, wreal, hreal, q0, value):
        self.lib.mjo_effect_add_dc_plane(_ptr(plane, _i16p), plane.shape[1], wreal, hreal, int(q0), int(value))


# --------------------------------------------------------------------------------------
# a libmodjpeg-compatible shared library (the reference build; also usable on the product)
# --------------------------------------------------------------------------------------


class _Component(C.Structure):  # reference: src/libmodjpeg.h:88-97 (mj_component_t)
    _fields_ = [("width_in_blocks", C.c_int), ("height_in_blocks", C.c_int), ("h_samp_factor", C.c_int),
                ("v_samp_factor", C.c_int), ("nblocks", C.c_int), ("blocks", C.POINTER(_f32p))]


class _CompiledDropon(C.Structure):  # reference: src/libmodjpeg.h:120-127
    _fields_ = [("image_ncomponents", C.c_int), ("image_colorspace", C.c_int), ("image", C.POINTER(_Component)),
                ("alpha_ncomponents", C.c_int), ("alpha", C.POINTER(_Component))]


class _Sampling(C.Structure):  # reference: src/libmodjpeg.h:76-84
    _fields_ = [("max_h", C.c_int), ("max_v", C.c_int), ("h_factor", C.c_int), ("v_factor", C.c_int),
                ("samp", C.c_int * 8)]


class _Dropon(C.Structure):  # reference: src/libmodjpeg.h:109-118
    _fields_ = [("image", _u8p), ("alpha", _u8p), ("width", C.c_int), ("height", C.c_int),
                ("colorspace", C.c_int), ("blend", C.c_int)]


class MjLibrary:
    """Any shared library exporting the 16 public mj_* functions (reference or product)."""

    def __init__(self, path: str, harness: str = HARNESS_SO):
        if not os.path.exists(harness):
            build()
        self.path = path
        self.lib = L = C.CDLL(path)
        self.h = H = C.CDLL(harness)
        self.sizeof_jpeg = H.mjh_sizeof_jpeg()
        vp = C.c_void_p
        L.mj_init_jpeg.argtypes = [vp]
        L.mj_init_jpeg.restype = None
        L.mj_free_jpeg.argtypes = [vp]
        L.mj_free_jpeg.restype = None
        L.mj_read_jpeg_from_memory.argtypes = [vp, C.c_char_p, C.c_size_t, C.c_size_t]
        L.mj_write_jpeg_to_memory.argtypes = [vp, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_int]
        L.mj_init_dropon.argtypes = [vp]
        L.mj_init_dropon.restype = None
        L.mj_free_dropon.argtypes = [vp]
        L.mj_free_dropon.restype = None
        L.mj_read_dropon_from_raw.argtypes = [vp, C.c_char_p, C.c_uint, C.c_int, C.c_int, C.c_short]
        L.mj_read_dropon_from_memory.argtypes = [vp, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_short]
        L.mj_compose.argtypes = [vp, vp, C.c_uint, C.c_int, C.c_int]
        for n in ("mj_effect_grayscale", "mj_effect_pixelate"):
            getattr(L, n).argtypes = [vp]
        L.mj_effect_tint.argtypes = [vp, C.c_int, C.c_int]
        L.mj_effect_luminance.argtypes = [vp, C.c_int]
        for n in ("mjh_image_info", ):
            getattr(H, n).argtypes = [vp, C.POINTER(C.c_int)]
        H.mjh_comp_info.argtypes = [vp, C.c_int, C.POINTER(C.c_int)]
        H.mjh_qtable.argtypes = [vp, C.c_int, _u16p]
        H.mjh_export_plane.argtypes = [vp, C.c_int, _i16p]
        H.mjh_import_plane.argtypes = [vp, C.c_int, _i16p]
        self.libc = C.CDLL(None)
        self.libc.free.argtypes = [C.c_void_p]
        self.libc.free.restype = None

    def read_jpeg(self, data: bytes, max_pixel: int = 0) -> "Jpeg":
        j = Jpeg(self)
        rv = self.lib.mj_read_jpeg_from_memory(j.ptr, data, len(data), max_pixel)
        if rv != 0:
            raise RuntimeError(f"mj_read_jpeg_from_memory -> {rv}")
        return j

    def dropon_from_raw(self, raw: np.ndarray, colorspace: int, blend: int = 255) -> "Dropon":
        d = Dropon(self)
        raw = np.ascontiguousarray(raw, np.uint8)
        h, w = raw.shape[:2]
        rv = self.lib.mj_read_dropon_from_raw(d.ptr, raw.tobytes(), colorspace, w, h, blend)
        if rv != 0:
            raise RuntimeError(f"mj_read_dropon_from_raw -> {rv}")
        return d


class Jpeg:
    def __init__(self, lib: MjLibrary):
        self.lib = lib
        self.buf = C.create_string_buffer(max(1024, lib.sizeof_jpeg))
        self.ptr = C.cast(self.buf, C.c_void_p)
        lib.lib.mj_init_jpeg(self.ptr)

    def info(self) -> dict:
        a = (C.c_int * 8)()
        if self.lib.h.mjh_image_info(self.ptr, a) != 0:
            raise RuntimeError("no image")
        return dict(ncomp=a[0], colorspace=a[1], width=a[2], height=a[3], max_h=a[4], max_v=a[5])

    def comp_info(self, c: int) -> dict:
        a = (C.c_int * 8)()
        if self.lib.h.mjh_comp_info(self.ptr, c, a) != 0:
            raise RuntimeError("bad component")
        return dict(wreal=a[0], hreal=a[1], h=a[2], v=a[3], wvirt=a[4], hvirt=a[5])

    def sampling(self) -> list[tuple[int, int]]:
        return [(self.comp_info(c)["h"], self.comp_info(c)["v"]) for c in range(self.info()["ncomp"])]

    def qtable(self, c: int) -> np.ndarray:
        q = np.zeros(64, np.uint16)
        assert self.lib.h.mjh_qtable(self.ptr, c, _ptr(q, _u16p)) == 0
        return q

    def plane(self, c: int) -> np.ndarray:
        ci = self.comp_info(c)
        a = np.zeros((ci["hvirt"], ci["wvirt"], 64), np.int16)
        assert self.lib.h.mjh_export_plane(self.ptr, c, _ptr(a, _i16p)) == 0
        return a

    def planes(self) -> list[np.ndarray]:
        return [self.plane(c) for c in range(self.info()["ncomp"])]

    def set_plane(self, c: int, a: np.ndarray) -> None:
        a = np.ascontiguousarray(a, np.int16)
        ci = self.comp_info(c)
        assert a.shape == (ci["hvirt"], ci["wvirt"], 64)
        assert self.lib.h.mjh_import_plane(self.ptr, c, _ptr(a, _i16p)) == 0

    def compose(self, d: "Dropon", align: int, ox: int = 0, oy: int = 0) -> int:
        return self.lib.lib.mj_compose(self.ptr, d.ptr, align, ox, oy)

    def grayscale(self) -> int:
        return self.lib.lib.mj_effect_grayscale(self.ptr)

    def pixelate(self) -> int:
        return self.lib.lib.mj_effect_pixelate(self.ptr)

    def tint(self, cb: int, cr: int) -> int:
        return self.lib.lib.mj_effect_tint(self.ptr, cb, cr)

    def luminance(self, v: int) -> int:
        return self.lib.lib.mj_effect_luminance(self.ptr, v)

    def write(self, options: int = 0) -> bytes:
        mem = C.c_void_p()
        n = C.c_size_t()
        rv = self.lib.lib.mj_write_jpeg_to_memory(self.ptr, C.byref(mem), C.byref(n), options)
        if rv != 0:
            raise RuntimeError(f"mj_write_jpeg_to_memory -> {rv}")
        out = C.string_at(mem, n.value)
        self.lib.libc.free(mem)
        return out

    def free(self) -> None:
        if self.buf is not None:
            self.lib.lib.mj_free_jpeg(self.ptr)
            self.buf = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Dropon:
    def __init__(self, lib: MjLibrary):
        self.lib = lib
        self.struct = _Dropon()
        self.ptr = C.cast(C.pointer(self.struct), C.c_void_p)
        lib.lib.mj_init_dropon(self.ptr)

    @property
    def width(self):
        return self.struct.width

    @property
    def height(self):
        return self.struct.height

    @property
    def colorspace(self):
        return self.struct.colorspace

    @property
    def blend(self):
        return self.struct.blend

    def image3(self) -> np.ndarray:
        n = 3 * self.width * self.height
        return np.ctypeslib.as_array(self.struct.image, (n,)).reshape(self.height, self.width, 3).copy()

    def alpha3(self) -> np.ndarray:
        n = 3 * self.width * self.height
        return np.ctypeslib.as_array(self.struct.alpha, (n,)).reshape(self.height, self.width, 3).copy()

    def free(self) -> None:
        if self.struct is not None:
            self.lib.lib.mj_free_dropon(self.ptr)
            self.struct = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Reference(MjLibrary):
    """The unmodified reference build, plus its exported internals used to check the port."""

    def __init__(self):
        if not have_reference():
            build()
        if not have_reference():
            raise FileNotFoundError(REF_SO)
        super().__init__(REF_SO)
        L = self.lib
        L.mj_compile_dropon.argtypes = [C.POINTER(_CompiledDropon), C.c_void_p, C.c_int, C.POINTER(_Sampling)] + [C.c_int] * 6
        L.mj_free_compileddropon.argtypes = [C.POINTER(_CompiledDropon)]
        L.mj_free_compileddropon.restype = None
        L.mj_compose_with_mask.argtypes = [C.c_void_p, C.POINTER(_CompiledDropon), C.c_int, C.c_int]
        L.mj_convolve.argtypes = [_f32p, _f32p, C.c_float, C.c_int, C.c_int]
        L.mj_convolve.restype = None

    def compile_handle(self, d: Dropon, colorspace: int, samp: list[tuple[int, int]], boff_x=0, boff_y=0, crop=None):
        """mj_compile_dropon (reference: src/dropon.c:325) -> (rv, mj_compileddropon_t); release with free_compiled()"""
        s = _Sampling()
        s.max_h = max(h for h, _ in samp)
        s.max_v = max(v for _, v in samp)
        s.h_factor = s.max_h * 8
        s.v_factor = s.max_v * 8
        for i, (h, v) in enumerate(samp):
            s.samp[2 * i] = h
            s.samp[2 * i + 1] = v
        cx, cy, cw, ch = crop if crop is not None else (0, 0, d.width, d.height)
        cd = _CompiledDropon()
        rv = self.lib.mj_compile_dropon(C.byref(cd), d.ptr, colorspace, C.byref(s), boff_x, boff_y, cx, cy, cw, ch)
        return rv, cd

    def compose_with_mask(self, j: "Jpeg", cd, block_x: int, block_y: int) -> int:
        """mj_compose_with_mask (reference: src/compose.c:237): the blend alone, on an already compiled dropon"""
        return self.lib.mj_compose_with_mask(j.ptr, C.byref(cd), block_x, block_y)

    def free_compiled(self, cd) -> None:
        self.lib.mj_free_compileddropon(C.byref(cd))

    def compile_dropon(self, d: Dropon, colorspace: int, samp: list[tuple[int, int]], boff_x=0, boff_y=0, crop=None):
        """mj_compile_dropon (reference: src/dropon.c:325) -> (rv, image[c], alpha[c]) float32 [hb][wb][64]"""
        rv, cd = self.compile_handle(d, colorspace, samp, boff_x, boff_y, crop)
        if rv != 0:
            return rv, None, None

        def grab(comps, n):
            out = []
            for c in range(n):
                comp = comps[c]
                a = np.zeros((comp.height_in_blocks, comp.width_in_blocks, 64), np.float32)
                flat = a.reshape(-1, 64)
                for b in range(comp.nblocks):
                    flat[b] = np.ctypeslib.as_array(comp.blocks[b], (64,))
                out.append(a)
            return out

        img = grab(cd.image, cd.image_ncomponents)
        alp = grab(cd.alpha, cd.alpha_ncomponents)
        self.lib.mj_free_compileddropon(C.byref(cd))
        return rv, img, alp

    def convolve(self, x, y, w, k, l):
        self.lib.mj_convolve(_ptr(x, _f32p), _ptr(y, _f32p), w, k, l)
