"""CPU restatement of libjpeg's baseline Huffman entropy encoder -- TEST INFRASTRUCTURE, never imported by the product.

What the reference's mj_write_jpeg_to_memory (reference: src/image.c:120-209) gets from jpeg_write_coefficients +
jpeg_finish_compress for a sequential, non-optimised, non-restart, single-scan file.  The algorithm lives in libjpeg
(a dependency of the reference, not in /root/reference): jctrans.c compress_output (MCU assembly, dummy blocks at the
right / bottom edge), jchuff.c encode_one_block / emit_bits / flush_bits, jchuff.c jpeg_make_c_derived_tbl (ITU-T T.81
Annex C), tables of Annex K.3.  Restated here from the published algorithm in plain Python loops (small images only).

PINNED: tests/test_huffman_oracle.py compares entropy_segment() byte for byte with the entropy-coded segment of files
written by the real libjpeg (through mj_write_jpeg_to_memory of the drop-in library's host path, which is libjpeg's
jpeg_write_coefficients) for 4:2:0 / 4:2:2 / 4:4:4 / grayscale images whose sizes are and are not multiples of the MCU.
"""
from __future__ import annotations

import numpy as np

# zigzag position -> natural index (T.81 figure A.6)
ZIGZAG = [0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28, 35, 42, 49,
          56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63]


def derive(counts: list[int], vals: list[int]) -> dict[int, tuple[int, int]]:
    """symbol -> (code, length) from a DHT-style table (counts[k] codes of length k + 1)"""
    sizes = []
    for length, cnt in enumerate(counts, start=1):
        sizes += [length] * cnt
    codes, code, si, p = [], 0, sizes[0] if sizes else 0, 0
    while p < len(sizes):
        while p < len(sizes) and sizes[p] == si:
            codes.append(code)
            code += 1
            p += 1
        code <<= 1
        si += 1
    return {v: (codes[i], sizes[i]) for i, v in enumerate(vals[:len(sizes)])}


class NotCodable(Exception):
    """a coefficient or symbol the tables cannot code (libjpeg: JERR_BAD_DCT_COEF / JERR_HUFF_MISSING_CODE)"""


def entropy_segment(planes: list[np.ndarray], real_dims: list[tuple[int, int]], samp: list[tuple[int, int]], width: int, height: int,
                    dc_tables: list[dict], ac_tables: list[dict], tbl_of_comp: list[int]) -> bytes:
    """planes[c]: int16 [rows][stride][64] natural order; real_dims[c] = (width_in_blocks, height_in_blocks)"""
    nc = len(planes)
    if nc == 1:
        hv = [(1, 1)]
        mcus_per_row, mcu_rows = real_dims[0]
    else:
        hv = samp
        max_h, max_v = max(h for h, _ in samp), max(v for _, v in samp)
        mcus_per_row, mcu_rows = -(-width // (8 * max_h)), -(-height // (8 * max_v))
    acc, nacc, out = 0, 0, bytearray()

    def put(code: int, size: int):
        nonlocal acc, nacc
        acc = (acc << size) | code
        nacc += size
        while nacc >= 8:
            b = (acc >> (nacc - 8)) & 0xFF
            out.append(b)
            if b == 0xFF:
                out.append(0)
            nacc -= 8
        acc &= (1 << nacc) - 1

    last_dc = [0] * nc
    for my in range(mcu_rows):
        for mx in range(mcus_per_row):
            for c in range(nc):
                h, v = hv[c]
                wb, hb = real_dims[c]
                dct, act = dc_tables[tbl_of_comp[c]], ac_tables[tbl_of_comp[c]]
                prev_in_mcu = None
                for yo in range(v):
                    for xo in range(h):
                        row, col = my * v + yo, mx * h + xo
                        if row < hb and col < wb:
                            blk = planes[c][row, col].astype(np.int64)
                        else:  # dummy block: no AC, the DC of the block before it in the MCU
                            blk = np.zeros(64, np.int64)
                            blk[0] = prev_in_mcu
                        prev_in_mcu = int(blk[0])
                        diff = int(blk[0]) - last_dc[c]
                        last_dc[c] = int(blk[0])
                        t, t2 = diff, diff
                        if t < 0:
                            t, t2 = -t, t2 - 1
                        nb = t.bit_length()
                        if nb > 11 or nb not in dct:
                            raise NotCodable("DC")
                        put(*dct[nb])
                        if nb:
                            put(t2 & ((1 << nb) - 1), nb)
                        r = 0
                        for k in range(1, 64):
                            val = int(blk[ZIGZAG[k]])
                            if val == 0:
                                r += 1
                                continue
                            while r > 15:
                                put(*act[0xF0])
                                r -= 16
                            t, t2 = val, val
                            if t < 0:
                                t, t2 = -t, t2 - 1
                            nb = t.bit_length()
                            if nb > 10 or ((r << 4) + nb) not in act:
                                raise NotCodable("AC")
                            put(*act[(r << 4) + nb])
                            put(t2 & ((1 << nb) - 1), nb)
                            r = 0
                        if r > 0:
                            put(*act[0])
    if nacc:  # flush_bits: fill the last byte with 1-bits
        put((1 << (8 - nacc)) - 1, 8 - nacc)
    return bytes(out)


def split_jpeg(data: bytes) -> tuple[bytes, bytes, bytes]:
    """(everything up to and including the SOS header, entropy-coded segment, trailer) of a single-scan JPEG file"""
    assert data[:2] == b"\xff\xd8"
    i = 2
    while True:
        assert data[i] == 0xFF, "marker expected"
        m = data[i + 1]
        n = (data[i + 2] << 8) | data[i + 3]
        i += 2 + n
        if m == 0xDA:
            break
    assert data[-2:] == b"\xff\xd9"
    return data[:i], data[i:-2], data[-2:]
