#!/bin/sh
# TEST INFRASTRUCTURE ONLY.  Builds the UNMODIFIED reference (maikmerten/p64) from the sources where they
# lie under /root/reference into oracle/_ref/ (git-ignored; travels to the GPU box with the snapshot).
#   p64_ref      stock encoder/decoder (three-step search, me.c:352)
#   p64_ref_fs   same sources, but MotionEstimation() calls FastBME instead of StepBME: the two lines
#                me.c:351-352 have their comment markers toggled by sed into a temp file outside the repo
#   libp64ref.so me.c mem.c chendct.c transform.c io.c (+ stat.c through ref_stat_shim.c) as a shared object for function-level checks
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
    $REF/vidinput.c $REF/y4m_input.c "$(dirname "$0")/ref_shim.c" "$(dirname "$0")/ref_stat_shim.c" -lm -o "$OUT/libp64ref.so"
# p64_gpu / p64_gpu_fs: the reference's own main() and stream writer with the body of p64EncodeFrame() replaced by one call
# into ../p64_b200/libp64b200.so per frame (examples/p64gpu_dropin.c).  p64.c goes through sed into the temp dir: the
# per-frame work is cut out and four calls are inserted; nothing of it is kept in the repository.
HERE=$(cd "$(dirname "$0")" && pwd)
LIBDIR=$HERE/../p64_b200
if [ -f "$LIBDIR/libp64b200.so" ]; then
  sed -e '/^void p64EncodeFrame()/,/^}/ s|^    GlobalMC();|    ;|' \
      -e '/^void p64EncodeFrame()/,/^}/ s|^  WritePictureHeader();|  ;|' \
      -e '/^void p64EncodeFrame()/,/^}/ s|^    p64EncodeGOB();|    ;|' \
      -e '/^void p64EncodeFrame()/,/^}/ s|^  for(CurrentGOB=0;|  p64gpu_frame_bits();\n  for(CurrentGOB=0;|' \
      -e '/^void p64EncodeSequence()/,/^}/ s|^  swopen(CImage->StreamFileName);|  swopen(CImage->StreamFileName);\n  p64gpu_init();|' \
      -e '/^void p64EncodeSequence()/,/^}/ s|^  GQuant=MQuant=InitialQuant;|  GQuant=MQuant=InitialQuant;\n  p64gpu_init_rate();|' \
      -e '/^void p64EncodeSequence()/,/^}/ s|^  WritePictureHeader();|  p64gpu_flush();\n  WritePictureHeader();|' \
      "$REF/p64.c" > "$TMP/p64_gpu.c"
  [ "$(grep -c 'p64gpu_' "$TMP/p64_gpu.c")" = 4 ]
  REST=""; for s in $SRCS; do [ "$s" = p64 ] || REST="$REST $REF/$s.c"; done
  for v in "" _fs; do
    DEF=""; [ -n "$v" ] && DEF="-DP64GPU_FULL"
    gcc $CF $DEF -I"$HERE/../include" "$TMP/p64_gpu.c" $REST $REF/me.c "$HERE/../examples/p64gpu_dropin.c" \
        -L"$LIBDIR" -lp64b200 -Wl,-rpath,'$ORIGIN/../../p64_b200' -lm -o "$OUT/p64_gpu$v"
  done
fi
# p64_gpu_hv / p64_gpu_hv_fs: the reference-shaped binding of INTEGRATION.md section 3 (examples/p64gpu_glue.c), executed: the
# reference keeps its own headers, VLC (WriteMBHeader, EncodeDC/AC, CBPEncodeAC), rate control (ExecuteQuantization, the
# overflow branch) and stream writer; GlobalMC, ReadCompressMDU and DecodeSaveMDU are replaced by the per-frame / per-GOB /
# per-macroblock calls into the library.  Again p64.c only passes through sed into the temp dir.
if [ -f "$LIBDIR/libp64b200.so" ]; then
  TMP2=$(mktemp -d)
  sed -e '1i extern int p64gpu_ovf;' \
      -e '/^void p64EncodeSequence()/,/^}/ s|^  swopen(CImage->StreamFileName);|  swopen(CImage->StreamFileName);\n  p64gpu_init();|' \
      -e '/^void p64EncodeFrame()/,/^}/ s|^  InstallFS(0,CFS);|  InstallFS(0,CFS);\n  p64gpu_frame_begin();|' \
      -e '/^void p64EncodeFrame()/,/^}/ s|^    GlobalMC();|    ;|' \
      -e '/^void p64EncodeFrame()/,/^}/ s|^  SwapFS(CFS,OFS);|  SwapFS(CFS,OFS);\n  p64gpu_frame_end();|' \
      -e '/^void p64EncodeGOB()/,/^}/ s|^  switch (ImageType)|  p64gpu_gob();\n  switch (ImageType)|' \
      -e '/^void p64EncodeGOB()/,/^}/ s|^      LastMType=MType;|      LastMType=MType; p64gpu_ovf=0;|' \
      -e '/^void p64EncodeGOB()/,/^}/ s|^[ \t]*MType=4;     /\* No coefficient transmission \*/|	  MType=4; p64gpu_ovf=1;|' \
      -e '/^static void p64EncodeMDU()/,/^}/ s|^  ReadCompressMDU();|  p64gpu_mb(p64gpu_ovf, inputbuf);|' \
      -e '/^static void p64EncodeMDU()/,/^}/ s|^  DecodeSaveMDU();|  ;|' \
      "$REF/p64.c" > "$TMP2/p64_gpu_hv.c"
  [ "$(grep -c 'p64gpu_' "$TMP2/p64_gpu_hv.c")" = 8 ]
  for v in "" _fs; do
    DEF=""; [ -n "$v" ] && DEF="-DP64GPU_FULL"
    gcc $CF $DEF -I"$HERE/../include" "$TMP2/p64_gpu_hv.c" $REST $REF/me.c "$HERE/../examples/p64gpu_glue.c" \
        -L"$LIBDIR" -lp64b200 -Wl,-rpath,'$ORIGIN/../../p64_b200' -lm -o "$OUT/p64_gpu_hv$v"
  done
  rm -rf "$TMP2"
fi
cp "$REF/test.intra" "$OUT/test.intra"
cp "$REF/short.p64" "$OUT/short.p64"      # the reference's own 1993 stream (SETUP:13-33): decoder known-answer test   # interpreter program fed on stdin for the intra-only config
rm -rf "$TMP"
echo "built: $(ls $OUT)"
