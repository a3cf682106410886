#!/bin/bash
# second GPU pass: lane kernels with the vectors in shared memory; A/B, ncu captures of the lane kernels
mkdir -p gpurun_out
AB_OUT=r02b_lane_ab.json timeout 900 python tools/lane_ab.py > gpurun_out/r02b_lane_ab.log 2>&1; echo "ab exit $?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r02b_lane_ab.json'))
for m,v in d.items():
    print(m,'rhs',v['rhs_per_s'],'jac',v['jac_per_s'])
    for k in ('ReactorRK64','ReactorCvode'):
        print('  ',k,v[k]['cells_per_s'],'fc_ident',v[k]['fc_identical_frac'],'err',v[k]['scaled_err_lane_vs_warp'])
P
for th in 16 96 200; do PLM_LANE_THETA=$th python tools/one_react.py drm19 ReactorCvode 65536 lane 1 2>&1 | tail -1 | sed "s/^/theta $th: /"; done
for th in 16 96 200; do PLM_LANE_THETA=$th python tools/one_react.py LiDryer ReactorCvode 65536 lane 1 2>&1 | tail -1 | sed "s/^/theta $th: /"; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lane_react -c 1 -o gpurun_out/r02b_drm19_bdf_lane python tools/one_react.py drm19 ReactorCvode 16384 lane > gpurun_out/r02b_ncu1.log 2>&1; tail -2 gpurun_out/r02b_ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lane_react -c 1 -o gpurun_out/r02b_lidryer_rk64_lane python tools/one_react.py LiDryer ReactorRK64 32768 lane > gpurun_out/r02b_ncu2.log 2>&1; tail -2 gpurun_out/r02b_ncu2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lane_rhs -c 1 -o gpurun_out/r02b_drm19_rhs_lane python tools/lane_ab.py > gpurun_out/r02b_ncu3.log 2>&1; tail -2 gpurun_out/r02b_ncu3.log
ls -la gpurun_out/*.ncu-rep
