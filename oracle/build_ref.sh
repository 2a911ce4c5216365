#!/bin/sh
# TEST INFRASTRUCTURE ONLY.  Builds the UNMODIFIED reference (maikmerten/p64) from the sources where they
# lie under /root/reference into oracle/_ref/ (git-ignored; travels to the GPU box with the snapshot).
#   p64_ref      stock encoder/decoder (three-step search, me.c:352)
#   p64_ref_fs   same sources, but MotionEstimation() calls FastBME instead of StepBME: the two lines
#                me.c:351-352 have their comment markers toggled by sed into a temp file outside the repo
#   libp64ref.so me.c mem.c chendct.c transform.c io.c as a shared object for function-level checks
# Flags: -O is the reference's own (makefile:6); -fcommon/-fgnu89-inline/-w only let 1993 C link under gcc 13
# (SURVEY.md F5).  No reference source is copied into the repository.
set -e
REF=${P64_REFERENCE:-/root/reference}
OUT=$(dirname "$0")/_ref
[ -d "$REF" ] || { echo "reference tree $REF not present; keeping prebuilt $OUT"; exit 0; }
mkdir -p "$OUT"
CF="-O -fcommon -fgnu89-inline -w -I$REF"
SRCS="p64 codec huffman io chendct lexer marker mem stat stream transform y4m_input vidinput"
ALL=""; for s in $SRCS; do ALL="$ALL $REF/$s.c"; done
gcc $CF $ALL $REF/me.c -lm -o "$OUT/p64_ref"
TMP=$(mktemp -d)
sed -e 's|^\(\s*\)//FastBME(x,y,pmem,x,y,fmem);|\1FastBME(x,y,pmem,x,y,fmem);|' \
    -e 's|^\(\s*\)StepBME(x,y,pmem,x,y,fmem);|\1//StepBME(x,y,pmem,x,y,fmem);|' "$REF/me.c" > "$TMP/me_fs.c"
grep -q '^\s*FastBME(x,y,pmem,x,y,fmem);' "$TMP/me_fs.c"
gcc $CF $ALL "$TMP/me_fs.c" -lm -o "$OUT/p64_ref_fs"
gcc $CF -fPIC -shared $REF/me.c $REF/mem.c $REF/chendct.c $REF/transform.c $REF/io.c \
    $REF/vidinput.c $REF/y4m_input.c "$(dirname "$0")/ref_shim.c" -lm -o "$OUT/libp64ref.so"
cp "$REF/test.intra" "$OUT/test.intra"   # interpreter program fed on stdin for the intra-only config
rm -rf "$TMP"
echo "built: $(ls $OUT)"
