#!/usr/bin/env bash
# TEST INFRASTRUCTURE -- not part of the product.
#
# Compiles the UNMODIFIED reference (pdlfs/old-vpic) from where it lies under
# $VPIC_REF (default /root/reference) into oracle/_ref/ (git-ignored).  No
# reference source is copied: we mirror the tree with symlinks exactly the way
# the reference's own config/bootstrap:11-17 does (its headers use
# "../sibling/x.h" includes, so the include dir must sit inside src/), feed the
# 72 files of config/vpic_source_list to gcc/g++ with the flags of
# cray-haswell.conf, and link against the MPI shim in oracle/mpi_shim.
#
# Outputs:
#   _ref/libvpic_ref_sse.so     reference as shipped: V4/SSE pipelines + pthreads
#                               (the CPU baseline: "V4 pthreads path")
#   _ref/libvpic_ref_scalar.so  same sources, hot-path .cxx built WITHOUT USE_V4_*
#                               so every particle/voxel goes through the scalar C
#                               pipelines (IEEE sqrt/div) -- the bit-exact oracle
#   _ref/libvpic_ref_sse.a      for linking decks (oracle/decks/*.cxx)
#   _ref/<deck>.op              deck executables (reference main.cxx + deck)
# Both .so also contain oracle/ref_harness.c (flat helpers for ctypes).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${VPIC_REF:-/root/reference}"
OUT="$HERE/_ref"
JOBS="${JOBS:-8}"

if [ ! -d "$REF/src" ]; then
  echo "build_ref: $REF not present; keeping prebuilt $OUT (if any)"; exit 0
fi

mkdir -p "$OUT/tree/src/include" "$OUT/obj_sse" "$OUT/obj_scalar"
# 1. symlink mirror
for d in "$REF"/src/*/; do
  n="$(basename "$d")"
  [ "$n" = include ] && continue
  ln -sfn "$d" "$OUT/tree/src/$n"
done
for f in main.cxx deck_wrapper.cxx; do ln -sfn "$REF/src/$f" "$OUT/tree/src/$f"; done
for h in $(cat "$REF/config/vpic_header_list"); do
  ln -sfn "$REF/$h" "$OUT/tree/src/include/$(basename "$h")"
done

COMMON="-D_XOPEN_SOURCE=600 -O2 -fno-strict-aliasing -fomit-frame-pointer -mfpmath=sse -fPIC -w -I$OUT/tree/src/include -I$HERE/mpi_shim"
CFLAGS="-std=gnu99 $COMMON"
CXXFLAGS="-std=gnu++98 $COMMON"

compile_one() { # flavour relpath
  local flav="$1" rel="$2" obj
  obj="$OUT/obj_$flav/$(echo "$rel" | tr '/' '_').o"
  local src="$OUT/tree/$rel"
  if [ "$obj" -nt "$REF/$rel" ]; then return 0; fi
  case "$rel" in
    *.c)   gcc $CFLAGS -c "$src" -o "$obj" ;;
    *.cxx)
      if [ "$flav" = sse ]; then
        g++ $CXXFLAGS -DUSE_V4_SSE -c "$src" -o "$obj"
      else
        case "$rel" in
          */v4/*) g++ $CXXFLAGS -DUSE_V4_PORTABLE -c "$src" -o "$obj" ;;  # V4-only TUs need some V4
          *)      g++ $CXXFLAGS -c "$src" -o "$obj" ;;                    # scalar pipelines everywhere
        esac
      fi ;;
  esac
}
export -f compile_one
export OUT REF CFLAGS CXXFLAGS

SRCS="$(cat "$REF/config/vpic_source_list")"
for flav in sse scalar; do
  printf '%s\n' $SRCS | xargs -P "$JOBS" -I{} bash -c "compile_one $flav {}"
  gcc $CFLAGS -c "$HERE/mpi_shim/mpi_shim.c" -o "$OUT/obj_$flav/mpi_shim.o"
  HARNESS_DEF=""
  [ "$flav" = sse ] && HARNESS_DEF="-DUSE_V4_SSE"
  g++ $CXXFLAGS $HARNESS_DEF -c "$HERE/ref_harness.cxx" -o "$OUT/obj_$flav/ref_harness.o"
  # src/vpic/*.cxx need the deck's user_* callbacks: they go into the .a only
  g++ -shared -Wl,-Bsymbolic -o "$OUT/libvpic_ref_$flav.so" \
      $(ls "$OUT"/obj_$flav/*.o | grep -v 'src_vpic_') -lm -lpthread
done
rm -f "$OUT/libvpic_ref_sse.a"
ar rcs "$OUT/libvpic_ref_sse.a" $(ls "$OUT"/obj_sse/*.o | grep -v ref_harness)

# 2. deck executables, built the way buildscript.in:9 does
for deck in "$HERE"/decks/*.cxx; do
  [ -e "$deck" ] || continue
  name="$(basename "$deck" .cxx)"
  g++ $CXXFLAGS -DUSE_V4_SSE -DINPUT_DECK="$deck" \
      "$OUT/tree/src/main.cxx" "$OUT/tree/src/deck_wrapper.cxx" \
      "$OUT/libvpic_ref_sse.a" -lm -lpthread -o "$OUT/$name.op"
done
echo "build_ref: ok -> $OUT"
