#!/bin/bash
# round 2: 8-GPU bench line (torchrun, one rank per GPU), as the driver launches it
mkdir -p gpurun_out
N=${1:-8}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 4 > gpurun_out/r02_bench_n${N}_raw.log 2> gpurun_out/r02_bench_n${N}.err
grep '^{' gpurun_out/r02_bench_n${N}_raw.log > gpurun_out/r02_bench_n${N}.json
python - $N <<'PY'
import json, sys
n = sys.argv[1]
d=json.loads(open(f'gpurun_out/r02_bench_n{n}.json').read())
print('n', d['n_gpus'], 'value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['one_caller']['value'], 'c5', d['config5_1m']['value'], d['clocks'])
PY
tail -3 gpurun_out/r02_bench_n${N}.err | cut -c1-300
