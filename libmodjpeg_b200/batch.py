"""Device-resident batches of JPEG coefficient planes and their sharding over GPUs.

The compositing path has no cross-image dependency (reference: src/compose.c:256-339 keeps no
state between blocks), so a batch is sharded by image: rank r of W owns a contiguous slice and
runs the same kernels on it; there is no collective on the data path.
"""
from __future__ import annotations

import numpy as np

from . import capi


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous slice [lo, hi) of n_items owned by `rank`; sizes differ by at most one."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class DeviceBatch:
    """n images with identical component geometry, resident in HBM in one slab.

    Layout: image-major; per image the component planes back to back, each plane int16
    [rows][stride_blocks][64] exactly as libjpeg holds it.  `descs_dev` is the device array of
    mjx_image_desc_t the kernels read.
    """

    def __init__(self, engine: capi.Engine, plane_shapes: list[tuple[int, int]], n: int, real_dims=None):
        self.engine = engine
        self.n = n
        self.plane_shapes = list(plane_shapes)  # (rows, stride_blocks) per component
        self.ncomp = len(plane_shapes)
        self.real_dims = real_dims or [(s, r) for r, s in plane_shapes]  # (wreal, hreal)
        self.plane_bytes = [r * s * 128 for r, s in plane_shapes]
        self.comp_offset = np.concatenate([[0], np.cumsum(self.plane_bytes)[:-1]]).astype(np.int64)
        self.image_bytes = int(sum(self.plane_bytes))
        self.blocks_per_image = int(sum(r * s for r, s in plane_shapes))
        self.slab = engine.device_alloc(max(1, self.image_bytes * n))
        self.descs_dev = engine.device_alloc(max(1, capi.IMAGE_DESC_DTYPE.itemsize * n))
        self.qtables = None

    def plane_ptr(self, i: int, c: int) -> int:
        return self.slab + i * self.image_bytes + int(self.comp_offset[c])

    def set_descs(self, qtables) -> None:
        """qtables: [ncomp][64] shared by all images, or [n][ncomp][64]."""
        ptrs = [[self.plane_ptr(i, c) for c in range(self.ncomp)] for i in range(self.n)]
        descs = capi.make_image_descs(ptrs, [s for _, s in self.plane_shapes], [r for r, _ in self.plane_shapes],
                                      qtables, self.real_dims)
        self.qtables = np.asarray(qtables, np.uint16)
        self.engine.copy_h2d(self.descs_dev, descs.view(np.uint8).reshape(-1))
        self.engine.sync()

    def upload_image(self, i: int, planes: list[np.ndarray]) -> None:
        for c, p in enumerate(planes):
            p = np.ascontiguousarray(p, np.int16)
            assert p.shape == (self.plane_shapes[c][0], self.plane_shapes[c][1], 64), (p.shape, self.plane_shapes[c])
            self.engine.copy_h2d(self.plane_ptr(i, c), p.reshape(-1).view(np.uint8))
        self.engine.sync()

    def download_image(self, i: int) -> list[np.ndarray]:
        out = []
        for c, (r, s) in enumerate(self.plane_shapes):
            a = np.zeros((r, s, 64), np.int16)
            self.engine.copy_d2h(a.reshape(-1).view(np.uint8), self.plane_ptr(i, c))
            out.append(a)
        self.engine.sync()
        return out

    def free(self) -> None:
        if self.slab:
            self.engine.device_free(self.slab)
            self.engine.device_free(self.descs_dev)
            self.slab = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
