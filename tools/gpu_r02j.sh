#!/bin/bash
# tenth GPU pass: lane-0-store scalars in the warp-per-cell kernels -- throughput, parity, racecheck; FlameSheet again
mkdir -p gpurun_out
python tools/one_react.py drm19 ReactorCvode 262144 warp 2 2>&1 | tail -1
python tools/one_react.py dodecane_lu ReactorCvode 16384 warp 2 2>&1 | tail -1
MAX_ORDER=5 python tools/one_react.py drm19 ReactorCvode 262144 warp 1 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_dropin.py -m gpu -x -q -s -k flamesheet > gpurun_out/r02j_flamesheet.log 2>&1; echo "flamesheet exit $?"
grep -E "FlameSheet 3 steps|passed|failed" gpurun_out/r02j_flamesheet.log | tail -4; grep -n "no plotfile" -A40 gpurun_out/r02j_flamesheet.log | head -80
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02j_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r02j_pytest.log
S=/usr/local/cuda/bin/compute-sanitizer
timeout 1200 $S --tool racecheck --racecheck-report analysis --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02j_racecheck_smoke.log 2>&1; echo "racecheck smoke rc=$?"; tail -6 gpurun_out/r02j_racecheck_smoke.log
timeout 1200 $S --tool racecheck --racecheck-report analysis --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "golden and dodecane_lu and cv_tight" > gpurun_out/r02j_racecheck_dodecane.log 2>&1; echo "racecheck dodecane rc=$?"; tail -6 gpurun_out/r02j_racecheck_dodecane.log
