#!/bin/bash
# A/B of the current library against libplmchem_old.so (+ the GPU suite with PYTEST=1)
mkdir -p gpurun_out
if [ "${PYTEST:-0}" = 1 ]; then echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/ab_pytest.log; fi
bash tools/gpu_variants.sh "base old base old" "${1:-dodecane_lu drm19 loose}"
