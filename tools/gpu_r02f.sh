#!/bin/bash
# sixth GPU pass: warp-cooperative LU + batched Jacobians in the lane kernels; theta sweep; suite (aligned)
mkdir -p gpurun_out
for th in 4 16 48; do
  PLM_LANE_ALIGN=1 PLM_LANE_THETA=$th python tools/one_react.py drm19 ReactorCvode 262144 lane 1 2>&1 | tail -1 | sed "s/^/align theta $th: /"
  PLM_LANE_ALIGN=1 PLM_LANE_THETA=$th python tools/one_react.py LiDryer ReactorCvode 262144 lane 1 2>&1 | tail -1 | sed "s/^/align theta $th: /"
done
PLM_LANE_ALIGN=0 PLM_LANE_THETA=16 python tools/one_react.py drm19 ReactorCvode 262144 lane 1 2>&1 | tail -1 | sed "s/^/free theta 16: /"
PLM_LANE_ALIGN=1 PLM_LANE_THETA=16 python tools/one_react.py drm19 ReactorCvode 1048576 lane 1 2>&1 | tail -1 | sed "s/^/align theta 16, 2^20 cells: /"
PLM_LANE_ALIGN=1 MAX_ORDER=5 python tools/one_react.py drm19 ReactorCvode 262144 lane 1 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lane_react -c 1 -o gpurun_out/r02f_drm19_bdf_lane_coop env PLM_LANE_ALIGN=1 PLM_LANE_THETA=16 python tools/one_react.py drm19 ReactorCvode 262144 lane > gpurun_out/r02f_ncu1.log 2>&1; tail -1 gpurun_out/r02f_ncu1.log
PLM_LANE_ALIGN=1 PLM_LANE_AUTO=all timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02f_pytest.log 2>&1; echo "pytest exit $?"
grep -E "kernels\]|passed|failed|Error" gpurun_out/r02f_pytest.log | tail -12
