#!/bin/bash
# bench each prebuilt library variant (variants/libplmchem_<tag>.so) on the small drm19 workload
for f in variants/libplmchem_*.so; do
  cp $f pelelm_b200/libplmchem.so
  echo "== $f"
  timeout 300 python bench.py --cells 262144 --steps 2 --warmup 2 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['per_cell']['nfe'])"
done
