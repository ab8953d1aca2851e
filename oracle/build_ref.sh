#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE.
# Stages the reference's own prebuilt CATHY processor ELFs (no Fortran compiler
# exists in this image, so the reference cannot be compiled) plus
# the runtime libraries they need into oracle/_ref/ (git-ignored, travels with
# gpurun).  Nothing is copied into tracked files.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${CATHY_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
SCIPY_LIBS="$(python - <<'PY'
import os, scipy
print(os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs"))
PY
)"
mkdir -p "$OUT/lib" "$OUT/bin"
gf=$(ls "$SCIPY_LIBS"/libgfortran-*.so.5.0.0 | head -1)
cp -f "$gf" "$OUT/lib/libgfortran.so.5"
for q in "$SCIPY_LIBS"/libquadmath-*.so.0.0.0; do cp -f "$q" "$OUT/lib/"; done
gcc -O2 -shared -fPIC -o "$OUT/lib/liblapack.so.3" "$HERE/lapack_shim.c" -lm
ln -sf liblapack.so.3 "$OUT/lib/libblas.so.3"
if [ -d "$REF" ]; then
  # name -> reference ELF (CATHY.H limits differ per build)
  cp -f "$REF/examples/SSHydro/weill_exemple/cathy"                         "$OUT/bin/cathy_20x20x15"
  cp -f "$REF/examples/SSHydro/weil_exemple_outputs_plot/cathy"             "$OUT/bin/cathy_20x20x15_newton"
  cp -f "$REF/examplesTmp/SSHydro/ERA5_ETp_spatially_from_weill/cathy"      "$OUT/bin/cathy_100x50x15"
  cp -f "$REF/examples/SSHydro/soil_withzones/cathy"                        "$OUT/bin/cathy_20x20x15_zones"   # MAXZON=4
  cp -f "$REF/examples/SSHydro/weill_exemple/prepro/pycppp"                 "$OUT/bin/pycppp"
  chmod +x "$OUT/bin/"*
fi
echo "oracle/_ref staged: $(ls "$OUT/bin" | tr '\n' ' ')"
