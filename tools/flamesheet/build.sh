#!/bin/bash
# Build the reference's Exec/RegTests/FlameSheet (2-D, drm19, no MPI, TinyProfiler) with the B200 reactor adaptor linked in
# (VERDICT r01 item 7).  Needs /root/reference and a built libplmchem.so + oracle/_ref (`make -C oracle ref` compiles SUNDIALS
# from the reference's sources; its objects become the static libraries PelePhysics' Make.ThirdParty expects, so cmake and
# the reference's ThirdParty build are not needed).  Nothing is written into /root/reference: the case directory is copied
# to a scratch directory and built there with the reference's own GNUmake.  Output: oracle/_ref/flamesheet/ (git-ignored).
set -e
REPO=$(cd "$(dirname "$0")/../.." && pwd)
REF=${PLM_REFERENCE:-/root/reference}
W=${TMPDIR:-/tmp}/plm_flamesheet
rm -rf "$W"; mkdir -p "$W/case" "$W/sun/lib" "$W/sun/include"
cp -r "$REF"/Exec/RegTests/FlameSheet/* "$W/case/"
cp "$REPO/tools/flamesheet/GNUmakefile.b200" "$W/case/"
for d in arkode cvode nvector sundials sunlinsol sunmatrix sunnonlinsol sunmemory; do
  [ -d "$REF/Submodules/sundials/include/$d" ] && mkdir -p "$W/sun/include/$d" && ln -sf "$REF/Submodules/sundials/include/$d"/* "$W/sun/include/$d/"
done
cp "$REPO"/oracle/refcfg/sundials/*.h "$W/sun/include/sundials/"
O="$REPO/oracle/_ref/obj/sun"
ar rcs "$W/sun/lib/libsundials_cvode.a" $O/cvode/*.o $O/sundials/*.o $O/sunmatrix/*/*.o $O/sunlinsol/*/*.o $O/sunnonlinsol/*/*.o
ar rcs "$W/sun/lib/libsundials_arkode.a" $O/arkode/*.o
ar rcs "$W/sun/lib/libsundials_nvecserial.a" $O/nvector/serial/*.o
make -C "$W/case" -f GNUmakefile.b200 -j"$(nproc)" PLMCHEM_HOME="$REPO" SUNDIALS_LIB_DIR="$W/sun/lib" > "$W/build.log" 2>&1 || { tail -20 "$W/build.log"; exit 1; }
mkdir -p "$REPO/oracle/_ref/flamesheet"
cp "$W/case/PeleLM2d.gnu.TPROF.ex" "$REPO/oracle/_ref/flamesheet/"
strip "$REPO/oracle/_ref/flamesheet/PeleLM2d.gnu.TPROF.ex"
cp "$REF/Exec/RegTests/FlameSheet/inputs.2d-regt" "$REF/Exec/RegTests/FlameSheet/drm19_pmf.dat" "$REPO/oracle/_ref/flamesheet/"
echo "built $REPO/oracle/_ref/flamesheet/PeleLM2d.gnu.TPROF.ex"
