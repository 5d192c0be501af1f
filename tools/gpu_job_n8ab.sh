#!/bin/bash
# round 2: 8-GPU end-to-end leg, spinning against blocking host waits, with per-rank figures
mkdir -p gpurun_out
N=${1:-8}
lscpu | grep -E "Model name|Thread|Core|Socket|^CPU\(s\)" 
for mode in 0 1; do
ICP_B200_BLOCKING_SYNC=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$mode bench.py --gpus $N --steps 20 --warmup 4 --no-config5 --headline-only > gpurun_out/r02ab_bench_n${N}_block$mode.log 2> gpurun_out/r02ab_bench_n${N}_block$mode.err
grep '^{' gpurun_out/r02ab_bench_n${N}_block$mode.log > gpurun_out/r02ab_bench_n${N}_block$mode.json
python - $N $mode <<'PY'
import json, sys
n, mode = sys.argv[1], sys.argv[2]
d=json.loads(open(f'gpurun_out/r02ab_bench_n{n}_block{mode}.json').read())
e=d['e2e']
r=lambda v:[round(x,2) for x in v]
print('blocking', mode, 'value', round(d['value']), 'e2e two', round(e['value']), 'one', round(e['one_caller']['value']), 'pipelined', round(e['one_caller_pipelined']['value']))
print('  dev ms', r(d['ms_per_step_every_rank'])); print('  two ms', r(e['ms_per_step_every_rank'])); print('  pipe ms', r(e['one_caller_pipelined']['ms_per_step_every_rank']))
PY
done
