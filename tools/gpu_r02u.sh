#!/bin/bash
# GPU pass: scalar-store groups merged (set_bdf + cvNls head, cvCompleteStep + cvPrepareNextStep, rescale, setup tail),
# acnrm in a register -- throughput, the parity suite, racecheck on smoke()
mkdir -p gpurun_out
{
timeout 200 python tools/one_react.py drm19 ReactorCvode 262144 warp 2 2>&1 | tail -2
MAX_ORDER=5 timeout 200 python tools/one_react.py drm19 ReactorCvode 262144 warp 1 2>&1 | tail -1
timeout 200 python tools/one_react.py dodecane_lu ReactorCvode 65536 warp 1 2>&1 | tail -1
} | tee gpurun_out/r02u_timing.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02u_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r02u_pytest.log
S=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $S --tool racecheck --racecheck-report analysis --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02u_racecheck_smoke.log 2>&1; echo "racecheck smoke rc=$?"; tail -6 gpurun_out/r02u_racecheck_smoke.log
