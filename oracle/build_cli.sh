#!/usr/bin/env bash
# Build the reference's command line tool, UNMODIFIED (/root/reference/src/contrib/modjpeg.c, compiled where
# it lies against the reference's own header), twice:
#   oracle/_ref/modjpeg_ref    linked against oracle/_ref/libmodjpeg_ref.so   (the reference, CPU)
#   oracle/_ref/modjpeg_b200   linked against libmodjpeg_b200/lib/libmodjpeg.so (this repo's drop-in, B200)
# The second binary is the literal drop-in check of BASELINE config 1: same program, same header, other
# library.  TEST INFRASTRUCTURE ONLY; outputs stay in oracle/_ref (git-ignored, they travel to the GPU box).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${MJ_REFERENCE_DIR:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -f "$REF/src/contrib/modjpeg.c" ]; then
    echo "build_cli: $REF/src/contrib/modjpeg.c not present (GPU box?) - keeping prebuilt binaries in $OUT" >&2
    exit 0
fi
PILLOW_LIBS="$(python3 -c "import os,PIL;print(os.path.join(os.path.dirname(os.path.dirname(PIL.__file__)),'pillow.libs'))")"
mkdir -p "$OUT"
CFLAGS="-O2 -Wall -Wno-unused-parameter -I$ROOT/third_party/jpeg62 -I$REF/src"
if [ -f "$OUT/libmodjpeg_ref.so" ]; then
    gcc $CFLAGS "$REF/src/contrib/modjpeg.c" -o "$OUT/modjpeg_ref" \
        -L"$OUT" -l:libmodjpeg_ref.so -ljpeg -Wl,-rpath,'$ORIGIN' -Wl,-rpath,"$PILLOW_LIBS"
fi
B200="$ROOT/libmodjpeg_b200/lib"
if [ -f "$B200/libmodjpeg.so" ]; then
    gcc $CFLAGS "$REF/src/contrib/modjpeg.c" -o "$OUT/modjpeg_b200" \
        -L"$B200" -lmodjpeg -ljpeg -Wl,-rpath,'$ORIGIN/../../libmodjpeg_b200/lib' -Wl,-rpath,"$PILLOW_LIBS"
fi
echo "build_cli: built $(ls "$OUT"/modjpeg_* 2>/dev/null | tr '\n' ' ')"
