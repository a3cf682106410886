#!/bin/bash
# GPU pass: which of the merged scalar-store groups pays?  v0 = commit 9015b05 (counters by lane 0; tn / jcur / qwait merged);
# a = acnrm in a register; b = convfail / jcur inside set_bdf's group; c = cvCompleteStep + cvPrepareNextStep as one group;
# d = rescale, setup tail, order change as groups; base = all four
mkdir -p gpurun_out
for v in v0 va vb vc vd vab vcd base; do
  lib=pelelm_b200/libplmchem_$v.so; [ "$v" = base ] && lib=pelelm_b200/libplmchem.so
  echo "== $v"; PLM_LIBPATH=$PWD/$lib timeout 120 python tools/one_react.py drm19 ReactorCvode 262144 warp 2 2>&1 | tail -1
done 2>&1 | tee gpurun_out/r02v_groups_ab.txt
