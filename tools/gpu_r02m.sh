#!/bin/bash
# thirteenth GPU pass: dodecane_lu with larger pools / fewer warps (closing the pooled-matrix question), and fresh
# ncu --set full captures (source imported) of the two warp-per-cell integration kernels the headline runs on
mkdir -p gpurun_out
for v in base p5w10 p6w8; do
  lib=pelelm_b200/libplmchem_$v.so; [ "$v" = base ] && lib=pelelm_b200/libplmchem.so
  echo "== $v"; PLM_LIBPATH=$PWD/$lib timeout 200 python tools/one_react.py dodecane_lu ReactorCvode 65536 warp 1 2>&1 | tail -1
done | tee gpurun_out/r02m_dodecane_ab.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:react_kernel -c 1 -f -o gpurun_out/r02m_drm19_warp python tools/one_react.py drm19 ReactorCvode 32768 warp > gpurun_out/r02m_ncu1.log 2>&1; tail -1 gpurun_out/r02m_ncu1.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:react_kernel -c 1 -f -o gpurun_out/r02m_dodecane_warp python tools/one_react.py dodecane_lu ReactorCvode 8192 warp > gpurun_out/r02m_ncu2.log 2>&1; tail -1 gpurun_out/r02m_ncu2.log
ls -la gpurun_out/*.ncu-rep
