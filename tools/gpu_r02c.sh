#!/bin/bash
# third GPU pass: drm19 through the table-driven lane kernels; A/B at 2^18 cells (13 cells per lane), LiDryer form A/B, ncu, parity suite
mkdir -p gpurun_out
AB_NCELL=262144 AB_OUT=r02c_lane_ab.json timeout 1200 python tools/lane_ab.py > gpurun_out/r02c_lane_ab.log 2>&1; echo "ab exit $?"
AB_NCELL=262144 AB_MECHS=LiDryer AB_OUT=r02c_lane_ab_alltab.json PLM_LIBPATH=$PWD/pelelm_b200/libplmchem_alltab.so timeout 600 python tools/lane_ab.py > gpurun_out/r02c_lane_ab_alltab.log 2>&1; echo "ab2 exit $?"
python - <<'P'
import json
for f in ('r02c_lane_ab.json','r02c_lane_ab_alltab.json'):
    d=json.load(open('gpurun_out/'+f))
    print(f)
    for m,v in d.items():
        print(m,'rhs',v['rhs_per_s'],'jac',v['jac_per_s'],'rhs diff',v['rhs_maxdiff_over_max'])
        for k in ('ReactorRK64','ReactorCvode'):
            print('  ',k,v[k]['cells_per_s'],'fc_ident',v[k]['fc_identical_frac'],'err',v[k]['scaled_err_lane_vs_warp'])
P
for th in 24 48 96 192; do PLM_LANE_THETA=$th python tools/one_react.py drm19 ReactorCvode 262144 lane 1 2>&1 | tail -1 | sed "s/^/theta $th: /"; done
MAX_ORDER=5 python tools/one_react.py drm19 ReactorCvode 262144 lane 1 2>&1 | tail -1
MAX_ORDER=5 python tools/one_react.py drm19 ReactorCvode 262144 warp 1 2>&1 | tail -1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lane_react -c 1 -o gpurun_out/r02c_drm19_bdf_lane python tools/one_react.py drm19 ReactorCvode 131072 lane > gpurun_out/r02c_ncu1.log 2>&1; tail -1 gpurun_out/r02c_ncu1.log
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02c_pytest.log 2>&1; echo "pytest exit $?"
grep -E "kernels\]|passed|failed|Error" gpurun_out/r02c_pytest.log | tail -20
