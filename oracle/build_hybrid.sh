#!/usr/bin/env bash
# TEST INFRASTRUCTURE -- not part of the product.
#
# The drop-in link of INTEGRATION.md, done for real: the reference's own host objects (compiled by build_ref.sh)
# WITHOUT the translation units of the hot path, plus libvpic_b200.so, linked into deck executables exactly the way
# buildscript.in:9 links a deck.  Proves that the library provides every symbol the rest of the reference needs and
# gives tests/test_gpu_deck.py an UNMODIFIED reference host program to run on the GPU.
#   _ref/hybrid/libvpic_host.a              reference minus hot path (util.c with its two allocation symbols localized)
#   _ref/hybrid/turbulence.b200.op          decks/trecon-part/turbulence.cxx as shipped (link check; it wants 4 ranks)
#   _ref/hybrid/{thermal,sheet,absorb}_small.b200.op   oracle/decks/*.cxx on the library
#   _ref/{thermal,sheet,absorb}_small.op               the same decks on the reference alone (scalar flavour of the hot path)
#   _ref/turbulence.op                                 the shipped trecon-part deck on the reference alone
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${VPIC_REF:-/root/reference}"
OUT="$HERE/_ref"
LIBDIR="$HERE/../old_vpic_b200"
if [ ! -d "$REF/src" ]; then
  echo "build_hybrid: $REF not present; keeping prebuilt $OUT/hybrid (if any)"; exit 0
fi
[ -d "$OUT/obj_sse" ] || bash "$HERE/build_ref.sh"
[ -e "$LIBDIR/libvpic_b200.so" ] || { echo "build_hybrid: build libvpic_b200.so first"; exit 1; }
mkdir -p "$OUT/hybrid"
HOT='src_field_advance_standard_|src_sf_interface_|src_species_advance_standard_|ref_harness|src_util_util.c.o'
objcopy --localize-symbol=util_malloc_aligned --localize-symbol=util_free_aligned \
        "$OUT/obj_sse/src_util_util.c.o" "$OUT/hybrid/util_local.o"
rm -f "$OUT/hybrid/libvpic_host.a"
ar rcs "$OUT/hybrid/libvpic_host.a" $(ls "$OUT"/obj_sse/*.o | grep -Ev "$HOT") "$OUT/hybrid/util_local.o"
COMMON="-D_XOPEN_SOURCE=600 -O2 -fno-strict-aliasing -fomit-frame-pointer -mfpmath=sse -fPIC -w -I$OUT/tree/src/include -I$HERE/mpi_shim"
link_hybrid() { # deck.cxx name
  g++ -std=gnu++98 $COMMON -DUSE_V4_SSE -DINPUT_DECK="$1" "$OUT/tree/src/main.cxx" "$OUT/tree/src/deck_wrapper.cxx" \
      "$OUT/hybrid/libvpic_host.a" -L"$LIBDIR" -lvpic_b200 -Wl,-rpath,'$ORIGIN/../../../old_vpic_b200' -lm -lpthread -rdynamic \
      -o "$OUT/hybrid/$2.b200.op"
}
link_hybrid "$REF/decks/trecon-part/turbulence.cxx" turbulence
link_hybrid "$HERE/decks/thermal_small.cxx" thermal_small
link_hybrid "$HERE/decks/sheet_small.cxx" sheet_small
link_hybrid "$HERE/decks/absorb_small.cxx" absorb_small
link_hybrid "$HERE/decks/thermal_c1.cxx" thermal_c1
# the same deck on the reference alone; hot path in its scalar flavour (what the library is bit-compatible with)
rm -f "$OUT/hybrid/libvpic_ref_scalar.a" "$OUT/hybrid/libvpic_ref_sse.a"
ar rcs "$OUT/hybrid/libvpic_ref_scalar.a" $(ls "$OUT"/obj_scalar/*.o | grep -v ref_harness)
for deck in thermal_small sheet_small absorb_small pin_history; do
  g++ -std=gnu++98 $COMMON -DINPUT_DECK="$HERE/decks/$deck.cxx" "$OUT/tree/src/main.cxx" "$OUT/tree/src/deck_wrapper.cxx" \
      "$OUT/hybrid/libvpic_ref_scalar.a" -lm -lpthread -o "$OUT/$deck.op"
done
# BASELINE configs[0] as a deck, on the reference as shipped (V4/SSE + pthreads hot path): the CPU arm of bench.py --deck-e2e
ar rcs "$OUT/hybrid/libvpic_ref_sse.a" $(ls "$OUT"/obj_sse/*.o | grep -v ref_harness)
g++ -std=gnu++98 $COMMON -DUSE_V4_SSE -DINPUT_DECK="$HERE/decks/thermal_c1.cxx" "$OUT/tree/src/main.cxx" "$OUT/tree/src/deck_wrapper.cxx" \
    "$OUT/hybrid/libvpic_ref_sse.a" -lm -lpthread -o "$OUT/thermal_c1.op"
# the reference's own trecon-part deck, as shipped (16x16x1 cells, topology 2x2x1): golden energies for the 4-rank run
g++ -std=gnu++98 $COMMON -DINPUT_DECK="$REF/decks/trecon-part/turbulence.cxx" "$OUT/tree/src/main.cxx" "$OUT/tree/src/deck_wrapper.cxx" \
    "$OUT/hybrid/libvpic_ref_scalar.a" -lm -lpthread -o "$OUT/turbulence.op"
echo "build_hybrid: ok -> $OUT/hybrid"
