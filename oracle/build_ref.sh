#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE ONLY.
#
# Compiles the UNMODIFIED reference (msesia/knockoffgwas, snpknock2/) from the
# sources where they lie under /root/reference into oracle/_ref/ (git-ignored,
# but shipped to the GPU box).  Nothing is copied out of the reference tree; the
# reference's own Makefiles are not run.  Flags follow snpknock2/Makefile:15-20
# (-Ofast -std=c++11) and the vendored-library Makefiles (-O3).
#
# Products:
#   oracle/_ref/snpknock2          stock reference binary (the oracle of record)
#   oracle/_ref/snpknock2_probe    same objects, linked with GNU ld --wrap in front
#                                  of Knockoffs::generate() and weighted_choice()
#                                  (oracle/ref_shims/wrap_probe.cpp): std::chrono
#                                  timer around generate(), full-precision dump of
#                                  the hot path's inputs and haplotype-level Hk,
#                                  optional trace of every categorical draw.
#
# Workarounds (see DESIGN.md "Oracle"): the image has libbz2.so.1.0 but no
# bzlib.h -> oracle/ref_shims/bzlib.h declares the public API; the reference does
# not ship sqlite3.c and nothing on the snpknock2 path needs it -> not linked.
set -euo pipefail

REPO="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
REF="${KGW_REFERENCE_DIR:-/root/reference}/snpknock2"
OUT="$REPO/oracle/_ref"
OBJ="$OUT/obj"
JOBS="${JOBS:-$(nproc)}"

if [ ! -d "$REF/src" ]; then
  echo "build_ref.sh: reference not present at $REF (expected on the GPU box); keeping prebuilt files" >&2
  exit 0
fi

mkdir -p "$OBJ"/{src,alglib,bgen,zstd,boost}
INC="$REF/include"
SHIM="$REPO/oracle/ref_shims"

APPFLAGS="-Ofast -std=c++11 -w -I$INC/bgen/genfile/include -I$INC/zstd-1.1.0/src -I$INC/alglib-3.15/src -I$INC/boost-1.55.0"
BOOSTFLAGS="-O2 -std=c++11 -w -I$INC/boost-1.55.0 -I$SHIM"
BGENFLAGS="-O3 -std=c++11 -w -I$INC/bgen/genfile/include -I$INC/db/include -I$INC/zstd-1.1.0/src -I$INC/sqlite3 -I$INC/boost-1.55.0"
ZSTDFLAGS="-O3 -w -I$INC/zstd-1.1.0/src -I$INC/zstd-1.1.0/src/common -I$INC/zstd-1.1.0/src/legacy -DZSTD_LEGACY_SUPPORT=1"

CMDS="$OUT/compile_cmds.txt"
: > "$CMDS"
emit() { # emit <obj> <command...>
  local o="$1"; shift
  if [ ! -f "$o" ]; then echo "$* -o $o" >> "$CMDS"; fi
}

# application objects (everything the reference Makefile lists; bgen_to_haps.cpp has its own main)
for f in "$REF"/src/*.cpp; do
  b="$(basename "$f" .cpp)"
  [ "$b" = "bgen_to_haps" ] && continue
  emit "$OBJ/src/$b.o" g++ $APPFLAGS -c "$f"
done
for f in "$INC"/alglib-3.15/src/*.cpp; do
  emit "$OBJ/alglib/$(basename "$f" .cpp).o" g++ -O3 -w -c "$f"
done
for f in "$INC"/bgen/src/*.cpp; do
  emit "$OBJ/bgen/$(basename "$f" .cpp).o" g++ $BGENFLAGS -c "$f"
done
for d in common compress decompress dictBuilder legacy; do
  for f in "$INC"/zstd-1.1.0/src/$d/*.c; do
    emit "$OBJ/zstd/${d}_$(basename "$f" .c).o" gcc $ZSTDFLAGS -c "$f"
  done
done
for d in system/src filesystem/src thread/src/pthread timer/src chrono/src iostreams/src; do
  for f in "$INC"/boost-1.55.0/src/$d/*.cpp; do
    tag="$(echo "$d" | tr '/' '_')"
    emit "$OBJ/boost/${tag}_$(basename "$f" .cpp).o" g++ $BOOSTFLAGS -c "$f"
  done
done

if [ -s "$CMDS" ]; then
  echo "build_ref.sh: compiling $(wc -l < "$CMDS") objects with $JOBS jobs"
  # run the compile lines in parallel; a failure of any aborts
  xargs -P "$JOBS" -d '\n' -I{} bash -c '{}' < "$CMDS"
fi

rm -f "$OUT"/lib*.a
ar rcs "$OUT/libalglib.a" "$OBJ"/alglib/*.o
ar rcs "$OUT/libbgen.a"   "$OBJ"/bgen/*.o
ar rcs "$OUT/libzstd.a"   "$OBJ"/zstd/*.o
ar rcs "$OUT/libboost.a"  "$OBJ"/boost/*.o

LIBS="$OUT/libalglib.a $OUT/libbgen.a $OUT/libzstd.a $OUT/libboost.a -lpthread -l:libbz2.so.1.0 -lz"

g++ -Ofast -std=c++11 -o "$OUT/snpknock2" "$OBJ"/src/*.o $LIBS

# --wrap variant: same reference objects + one shim object, no source patch.
GEN_SYM="$(nm "$OBJ/src/knockoffs.o" | awk '$2=="T" && $3 ~ /^_ZN9Knockoffs8generateEi$/{print $3}')"
WC_SYM="$(nm "$OBJ/src/utils.o" | awk '$2=="T" && $3 ~ /^_Z15weighted_choice/{print $3}')"
g++ -O2 -std=c++11 -w $(echo "$APPFLAGS" | sed 's/-Ofast//') -I"$REF/src" \
    -DKGW_GEN_SYM="$GEN_SYM" -DKGW_WC_SYM="$WC_SYM" \
    -c "$SHIM/wrap_probe.cpp" -o "$OBJ/wrap_probe.o"
g++ -Ofast -std=c++11 -o "$OUT/snpknock2_probe" "$OBJ"/src/*.o "$OBJ/wrap_probe.o" \
    -Wl,--wrap="$GEN_SYM" -Wl,--wrap="$WC_SYM" $LIBS
# the binary bench.py times: the same reference objects with the std::chrono timer around Knockoffs::generate and
# nothing else in the way (weighted_choice is the reference's own symbol here)
g++ -O2 -std=c++11 -w $(echo "$APPFLAGS" | sed 's/-Ofast//') -I"$REF/src" \
    -DKGW_GEN_SYM="$GEN_SYM" -DKGW_TIMER_ONLY \
    -c "$SHIM/wrap_probe.cpp" -o "$OBJ/wrap_timer.o"
g++ -Ofast -std=c++11 -o "$OUT/snpknock2_timer" "$OBJ"/src/*.o "$OBJ/wrap_timer.o" -Wl,--wrap="$GEN_SYM" $LIBS
echo "build_ref.sh: done -> $OUT"
