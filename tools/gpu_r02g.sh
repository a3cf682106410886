#!/bin/bash
# seventh GPU pass: final configuration -- parity suite, smoke, default bench (with the nested workloads), launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02g_pytest.log 2>&1; echo "pytest exit $?"
grep -E "kernels\]|passed|failed|Error" gpurun_out/r02g_pytest.log | tail -8
python __graft_entry__.py smoke 2>&1 | tail -3
timeout 1500 python bench.py > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; echo "bench exit $?"
tail -3 gpurun_out/r02g_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02g_bench.json').read().strip().splitlines()[-1])
print('main', d['value'], d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e']['value'], 'cpu', d.get('cpu_baseline',{}).get('value'))
for k,v in d.get('also',{}).items():
    print(k, v['value'], v['ms_per_step'], v['roofline']['frac'], 'e2e', (v.get('e2e') or {}).get('value'), 'cpu', (v.get('cpu_baseline') or {}).get('value'))
P
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02g_launches.csv python bench.py --steps 1 --warmup 1 --cells 1048576 --no-cpu-baseline --also-steps 1 > gpurun_out/r02g_bench_under_ncu.log 2>&1; echo "ncu exit $?"
