#!/bin/bash
# round 2: 8-GPU bench line with the pipelined one-thread end-to-end leg beside the two-caller leg
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N --steps 10 --warmup 4 > gpurun_out/r02sw_bench_n${N}_raw.log 2> gpurun_out/r02sw_bench_n${N}.err
grep '^{' gpurun_out/r02sw_bench_n${N}_raw.log > gpurun_out/r02sw_bench_n${N}.json
python - $N <<'PY'
import json, sys
n = sys.argv[1]
d=json.loads(open(f'gpurun_out/r02sw_bench_n{n}.json').read())
e=d['e2e']
r=lambda v:[round(x,2) for x in v]
print('n', d['n_gpus'], 'value', round(d['value']), 'e2e', round(e['value']), 'three callers', round(e['three_callers']['value']), 'one', round(e['one_caller']['value']), d['clocks'])
print(' dev ms', r(d['ms_per_step_every_rank'])); print(' e2e ms', r(e['ms_per_step_every_rank'])); print(' 3 callers ms', r(e['three_callers']['ms_per_step_every_rank']))
print(' config5', d.get('config5_1m'))
PY
nproc; free -g | head -2
tail -3 gpurun_out/r02sw_bench_n${N}.err | cut -c1-300
