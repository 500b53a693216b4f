"""CPU: the kernels' per-lane arithmetic (libmodjpeg_b200/csrc/mjx_math.cuh) compiled with g++ and
replayed block by block against the oracle -- catches arithmetic mistakes on the build box.
(The real parity tests are the -m gpu ones, which run the CUDA kernels through the C-ABI.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_build", "libmjx_emul.so")


@pytest.fixture(scope="module")
def emul(built):
    src = os.path.join(ROOT, "tests", "host_emul", "emul.cpp")
    hdr = os.path.join(ROOT, "libmodjpeg_b200", "csrc", "mjx_math.cuh")
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-I", os.path.dirname(hdr), src, "-o", SO])
    E = C.CDLL(SO)
    E.emul_tdiv_check.restype = C.c_longlong
    E.emul_requant_check.restype = C.c_longlong
    E.emul_uniform_pair_check.restype = C.c_longlong
    return E


def test_tdiv_reciprocal_is_exact(emul):
    # every int16 dividend against every 8-bit quantiser value, plus samples of the 16-bit range
    assert emul.emul_tdiv_check(1, 255) == 0
    for lo in (256, 1000, 4095, 20000, 65400):
        assert emul.emul_tdiv_check(lo, lo + 60) == 0


def test_requant_pair_division_is_exact(emul):
    # the fp32-pipe requantisation of the generic class: trunc(a / q) for every dequantised value the
    # fast path can see (|a| <= 2^17 covers int16 products plus any blend term), 8-bit and 16-bit tables
    for q in list(range(1, 256)) + [256, 1000, 4095, 20000, 65535]:
        assert emul.emul_requant_check(q, -(1 << 17), 1 << 17) == 0, q


def test_uniform_pair_equals_integer_formula(emul):
    # U / OPAQUE classes in the fp32 pipe == the integer formulation (which the oracle tests pin bit-exactly)
    for q in (1, 2, 3, 5, 8, 16, 17, 40, 99, 255):
        for wdc in (1, 8, 255, 1020, 1024, 2039, 2040):
            assert emul.emul_uniform_pair_check(q, wdc, 20000, q * 4099 + wdc) == 0, (q, wdc)


@pytest.mark.parametrize("v2", [0, 1])
def test_k2_arithmetic_vs_oracle(emul, built, port, v2):
    emul.emul_set_generic_v2(v2)
    from libmodjpeg_b200 import Jpeg
    from oracle import oracle_py as O

    i16p, u16p = C.POINTER(C.c_int16), C.POINTER(C.c_uint16)
    for subs, quality in [("420", 85), ("444", 95), ("422", 50)]:
        j = Jpeg()
        assert j.read_jpeg_from_memory(util.jpeg_bytes(256, 192, subs, quality, seed=7)) == 0
        info, samp = j.info(), j.sampling()
        L = O.make_layout(info["colorspace"], samp)
        for name, raw, cs, blend in [("logo", util.logo_rgba(256, 192, 64, 27), 2, 255), ("noise", util.noisy_rgba(256, 192, 5), 2, 255),
                                     ("uniform", util.noisy_rgba(256, 192, 6)[:, :, :3], 1, 128)]:
            i3, a3, scs, sblend = util.ingest_raw(raw, cs, blend)
            rv, D, W = port.compile_dropon(i3, a3, scs, L)
            assert rv == 0
            total = bad = 0
            for c in range(3):
                p0 = j.plane(c)
                a, b = p0.copy(), p0.copy()
                q = j.qtable(c)
                port.compose_plane(a, 0, 0, D[c], W[c], q)
                counts = (C.c_longlong * 4)()
                hb, wb = D[c].shape[:2]
                emul.emul_compose_plane(b.ctypes.data_as(i16p), b.shape[1], 0, 0, D[c].ctypes.data_as(i16p),
                                        W[c].ctypes.data_as(i16p), wb, hb, q.ctypes.data_as(u16p), counts)
                st = util.diff_stats(a, b)
                assert st["max_abs"] <= 1, (subs, name, c, st)
                total += st["n"]
                bad += st["differing"]
                if name == "uniform":
                    assert counts[1] == hb * wb and st["differing"] == 0  # class U: bit-exact
            assert bad <= max(2, total * 1e-4), (subs, name, bad, total)
