#!/bin/bash
for r in 0 5 2; do
timeout 600 python bench.py --steps 20 --warmup 4 --headline-only --no-cpu-baseline --as-rank $r > gpurun_out/r02t_bench_asrank$r.log 2> gpurun_out/r02t_bench_asrank$r.err
python - $r <<'PY'
import json, sys
r=sys.argv[1]
d=json.loads([l for l in open(f'gpurun_out/r02t_bench_asrank{r}.log') if l.startswith('{')][-1])
e=d['e2e']
print('as rank', r, 'value', round(d['value']), 'e2e three callers', round(e['value']), e['bit_identical_to_device_path'], 'one', round(e['one_caller']['value']), 'pipelined3', round(e['one_caller_pipelined']['value']), e['one_caller_pipelined']['bit_identical_to_device_path'])
PY
tail -2 gpurun_out/r02t_bench_asrank$r.err
done
