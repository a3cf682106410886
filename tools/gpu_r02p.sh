#!/bin/bash
# sixteenth GPU pass: the warp-per-cell kernels behind plm_react_boxes in ONE launch (box table, as the lane kernels):
# parity suite, then A/B against one launch per box (PLM_BOXES_PER_LAUNCH=1) on the headline batch, TripleFlame, dodecane_lu
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02p_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r02p_pytest.log
run() { # tag env args
  r=$(env $2 timeout 600 python bench.py $3 --no-also --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1)
  echo "$1 $2: $(echo "$r" | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.read()); print("%.0f cells/s  %.1f ms/step  frac %.4f launches %d" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["gpu_launches"]))
except Exception as e: print("FAILED", e)')"
}
{
run drm19_2p22 PLM_X=0 "--steps 1 --warmup 1"
run drm19_2p22 PLM_BOXES_PER_LAUNCH=1 "--steps 1 --warmup 1"
run tripleflame PLM_X=0 "--workload tripleflame --steps 1 --warmup 1"
run tripleflame PLM_BOXES_PER_LAUNCH=1 "--workload tripleflame --steps 1 --warmup 1"
run dodecane_8x16k PLM_X=0 "--mech dodecane_lu --cells 131072 --box-cells 16384 --steps 1 --warmup 1"
run dodecane_8x16k PLM_BOXES_PER_LAUNCH=1 "--mech dodecane_lu --cells 131072 --box-cells 16384 --steps 1 --warmup 1"
} 2>&1 | tee gpurun_out/r02p_boxes_ab.txt
