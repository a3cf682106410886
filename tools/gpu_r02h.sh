#!/bin/bash
# eighth GPU pass (2 GPUs): LiDryer RK64 bench after the NaN fast paths; the default bench under torchrun at N = 2
mkdir -p gpurun_out
timeout 300 python bench.py --mech LiDryer --integrator ReactorRK64 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02h_bench_lidryer_rk64.json 2> gpurun_out/r02h_bench_lidryer_rk64.err; echo "lidryer exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/r02h_bench_lidryer_rk64.json').read().strip().splitlines()[-1]); print('lidryer rk64', d['value'], d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e']['value'], d['config']['per_cell'])"
timeout 300 python bench.py --mech LiDryer --integrator ReactorCvode --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02h_bench_lidryer_cvode.json 2> gpurun_out/r02h_bench_lidryer_cvode.err; echo "lidryer cv exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/r02h_bench_lidryer_cvode.json').read().strip().splitlines()[-1]); print('lidryer cvode', d['value'], d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e']['value'])"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02h_bench_n2.json 2> gpurun_out/r02h_bench_n2.err; echo "n2 exit $?"
tail -3 gpurun_out/r02h_bench_n2.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02h_bench_n2.json').read().strip().splitlines()[-1])
print('main N=2', d['value'], d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e']['value'])
for k,v in d.get('also',{}).items():
    print(k, v['value'], v['ms_per_step'], v['roofline']['frac'], v['config'].get('boxes_moved'), v['config'].get('knapsack_efficiency'))
P
