#!/usr/bin/env bash
# integration/build_dropin.sh -- links the UNMODIFIED reference objects (built by oracle/build_ref.sh into
# oracle/_ref/obj) with integration/snpknock2_kgw_shim.cpp in front of Knockoffs::generate (GNU ld --wrap) and with
# knockoffgwas_b200/lib/libkgw.so: oracle/_ref/snpknock2_kgw is snpknock2 with the sampler on the B200.
# Needs /root/reference (headers) and a finished oracle/build_ref.sh; the binary travels to the GPU box with oracle/_ref/.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REPO="$(dirname "$HERE")"
REF="${KGW_REFERENCE_DIR:-/root/reference}/snpknock2"
OUT="$REPO/oracle/_ref"
OBJ="$OUT/obj"
LIBDIR="$REPO/knockoffgwas_b200/lib"
[ -d "$OBJ/src" ] || { echo "run oracle/build_ref.sh first" >&2; exit 1; }
[ -f "$LIBDIR/libkgw.so" ] || { echo "build libkgw.so first (python -m knockoffgwas_b200.build)" >&2; exit 1; }
GEN_SYM="$(nm "$OBJ/src/knockoffs.o" | awk '$2=="T" && $3 ~ /^_ZN9Knockoffs8generateEi$/{print $3}')"
RINC="$REF/include"
INC="-I$REF/src -I$RINC/bgen/genfile/include -I$RINC/zstd-1.1.0/src -I$RINC/alglib-3.15/src -I$RINC/boost-1.55.0 -I$REPO/include"
BGEN_SYM="$(nm "$OBJ/src/bgen_reader.o" | awk '$2=="T" && $3 ~ /^_ZN10BgenReader4readE.*Ebi$/{print $3}')"
WB_SYM="$(nm "$OBJ/src/plink_interface.o" | awk '$2=="T" && $3 ~ /^_ZN5plink12write_binaryE.*_b$/{print $3}')"
HAPS_SYM="$(nm "$OBJ/src/haps_reader.o" | awk '$2=="T" && $3 ~ /^_ZN10HapsReader4readE/{print $3}')"
AR_SYM="$(nm "$OBJ/src/kinship_computer.o" | awk '$2=="T" && $3 ~ /^_ZNK15KinshipComputer17assign_referencesE/{print $3}')"
KC_SYM="$(nm "$OBJ/src/kinship_computer.o" | awk '$2=="T" && $3 ~ /^_ZN15KinshipComputerC1ERKSt6vectorI8MetadataSaIS1_EEiiii$/{print $3}')"
[ -n "$KC_SYM" ] || { echo "could not find KinshipComputer's constructor" >&2; exit 1; }
[ -n "$GEN_SYM" ] && [ -n "$BGEN_SYM" ] && [ -n "$WB_SYM" ] && [ -n "$HAPS_SYM" ] && [ -n "$AR_SYM" ] || { echo "could not find the symbols to wrap" >&2; exit 1; }
g++ -O2 -std=c++11 -w $INC -DKGW_GEN_SYM="$GEN_SYM" -DKGW_BGEN_SYM="$BGEN_SYM" -DKGW_WB_SYM="$WB_SYM" -DKGW_HAPS_SYM="$HAPS_SYM" -DKGW_AR_SYM="$AR_SYM" -DKGW_KC_SYM="$KC_SYM" \
    -c "$HERE/snpknock2_kgw_shim.cpp" -o "$OBJ/kgw_shim.o"
LIBS="$OUT/libalglib.a $OUT/libbgen.a $OUT/libzstd.a $OUT/libboost.a -lpthread -l:libbz2.so.1.0 -lz"
g++ -Ofast -std=c++11 -o "$OUT/snpknock2_kgw" "$OBJ"/src/*.o "$OBJ/kgw_shim.o" -Wl,--wrap="$GEN_SYM" -Wl,--wrap="$BGEN_SYM" -Wl,--wrap="$WB_SYM" -Wl,--wrap="$HAPS_SYM" -Wl,--wrap="$AR_SYM" -Wl,--wrap="$KC_SYM" $LIBS \
    -L"$LIBDIR" -lkgw -Wl,-rpath,'$ORIGIN/../../knockoffgwas_b200/lib'
echo "build_dropin.sh: done -> $OUT/snpknock2_kgw"
