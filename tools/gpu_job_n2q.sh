#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 10 --warmup 4 --no-config5 --headline-only > gpurun_out/r02q_bench_n2_raw.log 2> gpurun_out/r02q_bench_n2.err
grep '^{' gpurun_out/r02q_bench_n2_raw.log > gpurun_out/r02q_bench_n2.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02q_bench_n2.json').read()); e=d['e2e']
print('n', d['n_gpus'], 'value', round(d['value']), 'e2e', round(e['value']), e['bit_identical_to_device_path'], 'three callers', round(e['three_callers']['value']), 'one', round(e['one_caller']['value']), e['ms_per_step_every_rank'])
PY
tail -2 gpurun_out/r02q_bench_n2.err | cut -c1-200
