#!/bin/bash
# round 2: 8-GPU bench line with the pipelined one-thread end-to-end leg beside the two-caller leg
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N --steps 10 --warmup 4 --no-config5 > gpurun_out/r02sw_bench_n${N}_raw.log 2> gpurun_out/r02sw_bench_n${N}.err
grep '^{' gpurun_out/r02sw_bench_n${N}_raw.log > gpurun_out/r02sw_bench_n${N}.json
python - $N <<'PY'
import json, sys
n = sys.argv[1]
d=json.loads(open(f'gpurun_out/r02sw_bench_n{n}.json').read())
e=d['e2e']
print('n', d['n_gpus'], 'value', d['value'], 'e2e two', e['value'], 'one', e['one_caller']['value'], 'pipelined', e['one_caller_pipelined']['value'], d['clocks'])
PY
nproc; free -g | head -2
tail -3 gpurun_out/r02sw_bench_n${N}.err | cut -c1-300
