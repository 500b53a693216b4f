"""libmodjpeg_b200 -- B200-native (sm_100a) engine for libmodjpeg's DCT-domain compositing
hot path: dropon compile (K1), masked block blend (K2), coefficient effects (K3), behind the
reference's public C API (include/libmodjpeg.h) and a kernel-level C-ABI (include/mjx.h).

Python here is plumbing only (ctypes bindings, batch sharding helpers); all arithmetic runs in
the CUDA kernels of ``csrc/``.  There is no CPU fallback.
"""
from .capi import (ALIGN_BOTTOM, ALIGN_CENTER, ALIGN_LEFT, ALIGN_RIGHT, ALIGN_TOP, CS_GRAYSCALE, CS_GRAYSCALEA, CS_RGB,
                   CS_RGBA, CS_YCC, CS_YCCA, CompiledDropon, Dropon, Engine, Jpeg, Layout, MjxError, geometry)

__all__ = ["Engine", "CompiledDropon", "Jpeg", "Dropon", "Layout", "MjxError", "geometry", "ALIGN_LEFT", "ALIGN_RIGHT",
           "ALIGN_TOP", "ALIGN_BOTTOM", "ALIGN_CENTER", "CS_RGB", "CS_RGBA", "CS_GRAYSCALE", "CS_GRAYSCALEA", "CS_YCC",
           "CS_YCCA"]
