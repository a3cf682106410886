#!/bin/bash
# ninth GPU pass: FlameSheet through the adaptor; final captures of the LiDryer lane kernels; dodecane + drm19 warp capture for traffic
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dropin.py -m gpu -x -q -s -k flamesheet > gpurun_out/r02i_flamesheet.log 2>&1; echo "flamesheet exit $?"
grep -E "FlameSheet|passed|failed|Error|rror" gpurun_out/r02i_flamesheet.log | tail -8
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lane_react -c 1 -o gpurun_out/r02i_lidryer_bdf_lane python tools/one_react.py LiDryer ReactorCvode 262144 lane > gpurun_out/r02i_ncu1.log 2>&1; tail -1 gpurun_out/r02i_ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lane_react -c 1 -o gpurun_out/r02i_lidryer_rk64_lane python tools/one_react.py LiDryer ReactorRK64 524288 lane > gpurun_out/r02i_ncu2.log 2>&1; tail -1 gpurun_out/r02i_ncu2.log
