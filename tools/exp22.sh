PLM_LIBPATH=$PWD/pelelm_b200/libplmchem_redux.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -4
bash tools/gpu_variants.sh "base redux reduxmb3" "drm19 dodecane_lu loose"
