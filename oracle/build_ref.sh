#!/usr/bin/env bash
# Build the UNMODIFIED reference (ioppermann/libmodjpeg, /root/reference/src/*.c) as
# oracle/_ref/libmodjpeg_ref.so, compiled from the sources where they lie.  No reference
# source is copied into this repository; only the built .so lands in oracle/_ref/
# (git-ignored, travels to the GPU box with gpurun).
#
# TEST INFRASTRUCTURE ONLY: loaded by tests/, __graft_entry__.smoke() and bench.py's
# cpu_baseline / --impl reference legs.  The product (libmodjpeg_b200) never loads it.
#
# The reference needs <jpeglib.h>; the image has a libjpeg-turbo 3.1.x runtime (Pillow's
# pillow.libs/libjpeg-*.so.62) but no headers, so third_party/jpeg62/ declares the ABI-62
# interface.  Flags are the reference's own (-O2, CMakeLists.txt:32) minus -Werror.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${MJ_REFERENCE_DIR:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/src" ]; then
    echo "build_ref: $REF/src not present (GPU box?) - keeping prebuilt $OUT" >&2
    exit 0
fi
JPEGSO="$(python3 - <<'PY'
import glob, os, PIL
d = os.path.join(os.path.dirname(os.path.dirname(PIL.__file__)), "pillow.libs")
c = sorted(glob.glob(os.path.join(d, "libjpeg-*.so.62*")))
print(c[0] if c else "")
PY
)"
[ -n "$JPEGSO" ] || { echo "build_ref: no libjpeg .so.62 found in pillow.libs" >&2; exit 1; }
mkdir -p "$OUT"
ln -sf "$JPEGSO" "$OUT/libjpeg.so"
# PNG dropons (WITH_LIBPNG, CMakeLists.txt:18-29): Pillow's libpng16 runtime + third_party/png16/png.h
PNGSO="$(ls "$(dirname "$JPEGSO")"/libpng16-*.so.16* 2>/dev/null | head -1 || true)"
PNGFLAGS=""
PNGLINK=""
if [ -n "$PNGSO" ]; then
    ln -sf "$PNGSO" "$OUT/libpng16.so"
    PNGFLAGS="-DWITH_LIBPNG -I$ROOT/third_party/png16"
    PNGLINK="-lpng16"
fi
gcc -O2 -Wall -Wextra -Wpointer-arith -Wno-uninitialized -Wno-unused-parameter \
    -Wno-deprecated-declarations -fPIC -shared $PNGFLAGS \
    -I"$ROOT/third_party/jpeg62" -I"$REF/src" \
    "$REF"/src/compose.c "$REF"/src/convolve.c "$REF"/src/dropon.c \
    "$REF"/src/effect.c "$REF"/src/image.c "$REF"/src/jpeg.c \
    -o "$OUT/libmodjpeg_ref.so" \
    -L"$OUT" -ljpeg $PNGLINK -lm -Wl,-rpath,"$(dirname "$JPEGSO")"
echo "build_ref: built $OUT/libmodjpeg_ref.so against $JPEGSO"
