#!/bin/bash
# GPU pass: the store-group merges on top of the shipped one (c): a = acnrm in a register, b = convfail / jcur inside set_bdf's group
mkdir -p gpurun_out
for v in base vac vbc vabc; do
  lib=pelelm_b200/libplmchem_$v.so; [ "$v" = base ] && lib=pelelm_b200/libplmchem.so
  echo "== $v"; PLM_LIBPATH=$PWD/$lib timeout 120 python tools/one_react.py drm19 ReactorCvode 262144 warp 2 2>&1 | tail -1
  PLM_LIBPATH=$PWD/$lib timeout 120 python tools/one_react.py dodecane_lu ReactorCvode 32768 warp 1 2>&1 | tail -1
done 2>&1 | tee gpurun_out/r02y_groups_ab2.txt
