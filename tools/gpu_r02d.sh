#!/bin/bash
# fourth GPU pass: drm19 straight-line lane kernels at 2^18 cells, free-running vs phase-aligned rounds; multi-box launch; suite
mkdir -p gpurun_out
for al in 0 1; do
  PLM_LANE_ALIGN=$al python tools/one_react.py drm19 ReactorCvode 262144 lane 1 2>&1 | tail -1 | sed "s/^/align $al: /"
  PLM_LANE_ALIGN=$al python tools/one_react.py drm19 ReactorRK64 262144 lane 1 2>&1 | tail -1 | sed "s/^/align $al: /"
  PLM_LANE_ALIGN=$al python tools/one_react.py LiDryer ReactorCvode 262144 lane 1 2>&1 | tail -1 | sed "s/^/align $al: /"
  PLM_LANE_ALIGN=$al python tools/one_react.py LiDryer ReactorRK64 524288 lane 1 2>&1 | tail -1 | sed "s/^/align $al: /"
done
python tools/one_react.py drm19 ReactorCvode 262144 warp 1 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lane_react -c 1 -o gpurun_out/r02d_drm19_bdf_lane_aligned env PLM_LANE_ALIGN=1 python tools/one_react.py drm19 ReactorCvode 65536 lane > gpurun_out/r02d_ncu1.log 2>&1; tail -1 gpurun_out/r02d_ncu1.log
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02d_pytest.log 2>&1; echo "pytest exit $?"
grep -E "kernels\]|passed|failed|Error" gpurun_out/r02d_pytest.log | tail -20
