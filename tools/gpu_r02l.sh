#!/bin/bash
# twelfth GPU pass: dodecane_lu with pooled Newton matrices (12 warps + 4 pool matrices per SM) against the 7-warp kernel
# (PLM_BIG_POOL=0) and two other pool shapes; parity of the dodecane cases; one ncu capture of the pooled kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "dodecane" > gpurun_out/r02l_pytest_dodecane.log 2>&1; echo "pytest dodecane exit $?"; tail -3 gpurun_out/r02l_pytest_dodecane.log
for v in base nopool p4w10 p3w14; do
  lib=pelelm_b200/libplmchem_$v.so; [ "$v" = base ] && lib=pelelm_b200/libplmchem.so
  echo "== $v"; PLM_LIBPATH=$PWD/$lib timeout 300 python tools/one_react.py dodecane_lu ReactorCvode 65536 warp 2 2>&1 | tail -2
done | tee gpurun_out/r02l_dodecane_ab.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:react_kernel -c 1 -o gpurun_out/r02l_dodecane_pool python tools/one_react.py dodecane_lu ReactorCvode 16384 warp > gpurun_out/r02l_ncu.log 2>&1; tail -1 gpurun_out/r02l_ncu.log
