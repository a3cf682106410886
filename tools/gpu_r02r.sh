#!/bin/bash
# eighteenth GPU pass: where do the DRAM bytes of the `--set full` captures come from?  The same launch with three metric sets.
mkdir -p gpurun_out
{
echo "== dodecane_lu 8192 cells, --metrics dram only"
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:react_kernel -c 1 python tools/one_react.py dodecane_lu ReactorCvode 8192 warp 2>&1 | grep -E "dram__" | sed 's/  */ /g'
echo "== dodecane_lu 8192 cells, --section MemoryWorkloadAnalysis"
timeout 300 ncu --section MemoryWorkloadAnalysis --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:react_kernel -c 1 python tools/one_react.py dodecane_lu ReactorCvode 8192 warp 2>&1 | grep -E "dram__|DRAM|Mem Busy|Max Bandwidth" | sed 's/  */ /g'
echo "== dodecane_lu 8192 cells, --set full (no source import)"
timeout 600 ncu --set full --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:react_kernel -c 1 python tools/one_react.py dodecane_lu ReactorCvode 8192 warp 2>&1 | grep -E "dram__|cells:" | sed 's/  */ /g'
} 2>&1 | tee gpurun_out/r02r_dram_modes.txt
