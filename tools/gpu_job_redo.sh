#!/bin/bash
# round 2: guard re-runs chunk by chunk beside the later chunks (host-pointer path) — GPU suite, flagged buffers, bench headline
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r02r_pytest.log; tail -4 gpurun_out/r02r_pytest.log
PYTHONPATH=. python tools/redo_by_rank.py > gpurun_out/r02r_redo_by_rank.jsonl; grep -E '"rank": (2|5|6)' gpurun_out/r02r_redo_by_rank.jsonl
timeout 600 python bench.py --steps 20 --warmup 4 --headline-only --no-cpu-baseline > gpurun_out/r02r_bench.log 2> gpurun_out/r02r_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02r_bench.log') if l.startswith('{')][-1])
e=d['e2e']
print('value', round(d['value']), 'e2e two', round(e['value']), 'one', round(e['one_caller']['value']), 'pipelined', round(e['one_caller_pipelined']['value']), e['bit_identical_to_device_path'], e['one_caller_pipelined']['bit_identical_to_device_path'])
PY
tail -2 gpurun_out/r02r_bench.err
