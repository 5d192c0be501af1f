#!/bin/bash
PYTHONPATH=. python tools/flagged_chunk.py
for r in 5 6; do
timeout 600 python bench.py --steps 20 --warmup 4 --headline-only --no-cpu-baseline --as-rank $r > gpurun_out/r02r_bench_asrank$r.log 2> gpurun_out/r02r_bench_asrank$r.err
python - $r <<'PY'
import json, sys
r=sys.argv[1]
d=json.loads([l for l in open(f'gpurun_out/r02r_bench_asrank{r}.log') if l.startswith('{')][-1])
e=d['e2e']
print('as rank', r, 'value', round(d['value']), 'serial', round(d['step_overlap']['steps_one_after_the_other']['value']), 'e2e two', round(e['value']), 'one', round(e['one_caller']['value']), 'pipelined', round(e['one_caller_pipelined']['value']))
PY
done
