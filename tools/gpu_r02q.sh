#!/bin/bash
# seventeenth GPU pass: saved dodecane_lu Jacobian through L2-only (.cg, base) against streaming (.cs, libplmchem_cs.so) accesses:
# throughput and DRAM bytes of one launch; drm19 throughput / DRAM bytes of the shipped kernel after the box-table change
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct
{
for v in base cs; do
  lib=pelelm_b200/libplmchem_$v.so; [ "$v" = base ] && lib=pelelm_b200/libplmchem.so
  echo "== dodecane_lu $v"
  PLM_LIBPATH=$PWD/$lib timeout 200 python tools/one_react.py dodecane_lu ReactorCvode 65536 warp 1 2>&1 | tail -1
  PLM_LIBPATH=$PWD/$lib timeout 300 ncu --metrics $M --clock-control none -k regex:react_kernel -c 1 python tools/one_react.py dodecane_lu ReactorCvode 8192 warp 2>&1 | grep -E "dram__|lts__|cells:" | sed 's/  */ /g'
done
echo "== drm19 base"
timeout 200 python tools/one_react.py drm19 ReactorCvode 262144 warp 2 2>&1 | tail -2
timeout 300 ncu --metrics $M --clock-control none -k regex:react_kernel -c 1 python tools/one_react.py drm19 ReactorCvode 32768 warp 2>&1 | grep -E "dram__|lts__|cells:" | sed 's/  */ /g'
} 2>&1 | tee gpurun_out/r02q_jsave_ab.txt
