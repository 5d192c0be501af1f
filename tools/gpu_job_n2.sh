#!/bin/bash
# round 2: 2-GPU bench line (torchrun, one rank per GPU) with overlapping steps, two-caller e2e and config 5
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 4 > gpurun_out/r02_bench_n2_raw.log 2> gpurun_out/r02_bench_n2.err
grep '^{' gpurun_out/r02_bench_n2_raw.log > gpurun_out/r02_bench_n2.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n2.json').read())
print('n', d['n_gpus'], 'value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['one_caller']['value'], 'c5', d['config5_1m'])
PY
tail -3 gpurun_out/r02_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | grep '^{' | cut -c1-250
