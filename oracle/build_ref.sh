#!/bin/bash
# Builds the UNMODIFIED reference library (wbuntine/libstb, /root/reference/lib/*.c) into
# oracle/_ref/ as shared objects, compiling the sources where they lie.  Nothing from the
# reference is copied into the repository: oracle/_ref/ is git-ignored and holds only .so files.
#
#   oracle/_ref/libstb_ref.so        default configuration of lib/Makefile:7 (-O5 -DNDEBUG -DH_THREADS;
#                                    PSAMPLE_ARS + LS_NOPOLYGAMMA defined => ARS samplers, digammaRN)
#   oracle/_ref/libstb_ref_slice.so  slice-sampler configuration: PSAMPLE_ARS and
#                                    LS_NOPOLYGAMMA undefined and polygamma.c added, so that
#                                    samplea/sampleb run SliceSimple and bmax/digammaInv exist.
#   oracle/_ref/libstb_ref_slice_m.so  the slice configuration with -DSAMPLEA_M: samplea2 as well.
#
# The two switches are "#define"s inside lib/psample.h:37 and lib/digamma.h:25, not -D flags, so the
# slice build compiles through a throw-away directory under /tmp that holds symlinks to the sources
# and sed-filtered copies of those two headers; the directory is removed afterwards.
#
# This is TEST INFRASTRUCTURE (the checker / the CPU baseline), never part of the product path.
set -euo pipefail
REF=${STB_REFERENCE_DIR:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
CC=${CC:-gcc}
if [ ! -d "$REF/lib" ]; then
  echo "build_ref.sh: $REF/lib not present (GPU box?) - keeping prebuilt files in $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
CFLAGS="-O3 -DNDEBUG -DH_THREADS -fPIC -w"
SRC="stable digamma arms sapprox sslice sampleb samplea yaps lgamma sympoly digammainv gslrandist"

# --- default (ARS) configuration, sources compiled in place ---
files=""
for s in $SRC; do files="$files $REF/lib/$s.c"; done
$CC $CFLAGS -shared -o "$OUT/libstb_ref.so" $files -lm -lpthread

# --- slice configuration ---
TMP=$(mktemp -d /tmp/stbref.XXXXXX)
trap 'rm -rf "$TMP"' EXIT
for f in "$REF"/lib/*.c "$REF"/lib/*.h; do ln -s "$f" "$TMP/$(basename "$f")"; done
rm "$TMP/psample.h" "$TMP/digamma.h"
sed 's|^#define PSAMPLE_ARS|// &|' "$REF/lib/psample.h" > "$TMP/psample.h"
sed 's|^#define LS_NOPOLYGAMMA|// &|' "$REF/lib/digamma.h" > "$TMP/digamma.h"
files=""
for s in $SRC polygamma; do files="$files $TMP/$s.c"; done
$CC $CFLAGS -shared -o "$OUT/libstb_ref_slice.so" $files -lm -lpthread
# the same with -DSAMPLEA_M (lib/psample.h:30): adds samplea2 / logminus (lib/samplea.c:227-341)
$CC $CFLAGS -DSAMPLEA_M -shared -o "$OUT/libstb_ref_slice_m.so" $files -lm -lpthread
# recorder for the gcache_value calls of the reference's aterms2 (tests/ref_samplea2_probe.py)
$CC -O2 -fPIC -shared -o "$OUT/shim_gcache.so" "$HERE/shim_gcache.c"
echo "built $OUT/libstb_ref.so $OUT/libstb_ref_slice.so $OUT/libstb_ref_slice_m.so"

# --- the reference's own test programs (test/list.c, test/demo.c, test/check.c), UNMODIFIED ---
#   oracle/_ref/ref_list      list.c linked with the reference library: generates tests/golden/list_*.txt
#   oracle/_ref/dropin_{list,demo,check}
#                             the same sources compiled against THIS repo's include/ and linked with
#                             libstb_b200.so: the drop-in proof (they need a GPU to run)
$CC -O2 -w -I"$REF/lib" "$REF/test/list.c" -o "$OUT/ref_list" "$OUT/libstb_ref.so" -Wl,-rpath,'$ORIGIN' -lm
#   oracle/_ref/ref_demo, ref_check   demo.c / check.c linked with the reference library (its default, ARS,
#                             configuration), and shim_time.so, the fixed clock that makes them (and the
#                             drop-in builds below) deterministic: tests/golden/programs/*.txt
for p in demo check; do
  $CC -O2 -w -I"$REF/lib" "$REF/test/$p.c" -o "$OUT/ref_$p" "$OUT/libstb_ref.so" -Wl,-rpath,'$ORIGIN' -lm
done
$CC -O2 -fPIC -shared -o "$OUT/shim_time.so" "$HERE/shim_time.c"
#   oracle/_ref/dropin_bench_{ref,b200}   oracle/bench_dropin.c (the default-flags caller pattern, timed) against either library
$CC -O2 -w -I"$REF/lib" "$HERE/bench_dropin.c" -o "$OUT/dropin_bench_ref" "$OUT/libstb_ref.so" -Wl,-rpath,'$ORIGIN' -lm
LIBDIR="$HERE/../libstb_b200/lib"
if [ -f "$LIBDIR/libstb_b200.so" ]; then
  for p in list demo check; do
    $CC -O2 -w -I"$HERE/../include" "$REF/test/$p.c" -o "$OUT/dropin_$p" -L"$LIBDIR" -lstb_b200 \
        -Wl,-rpath,'$ORIGIN/../../libstb_b200/lib' -lm
  done
  $CC -O2 -w -I"$HERE/../include" "$HERE/bench_dropin.c" -o "$OUT/dropin_bench_b200" -L"$LIBDIR" -lstb_b200 \
      -Wl,-rpath,'$ORIGIN/../../libstb_b200/lib' -lm
  echo "built $OUT/ref_list and $OUT/dropin_{list,demo,check} and $OUT/dropin_bench_{ref,b200}"
fi
