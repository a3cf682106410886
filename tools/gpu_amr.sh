#!/bin/bash
# configs[4] pass: the AMR parity test and the tripleflame workload of bench.py
mkdir -p gpurun_out
echo "== pytest amr"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "amr" 2>&1 | tail -15 | tee gpurun_out/amr_pytest.log
echo "== bench tripleflame drm19"
timeout 600 python bench.py --workload tripleflame --steps 2 --warmup 3 2>&1 | tail -1 > gpurun_out/amr_bench_drm19.json; cut -c1-1500 gpurun_out/amr_bench_drm19.json
echo "== bench tripleflame drm19 rtol 1e-6 order 4"
timeout 600 python bench.py --workload tripleflame --rtol 1e-6 --max-order 4 --steps 2 --warmup 3 2>&1 | tail -1 > gpurun_out/amr_bench_drm19_loose.json; cut -c1-600 gpurun_out/amr_bench_drm19_loose.json
