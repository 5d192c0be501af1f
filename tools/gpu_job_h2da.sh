#!/bin/bash
# round 2: chunk inputs copied ahead on a copy stream (ICP_B200_H2D_AHEAD) against in-stream copies; GPU suite on the new path
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02h_pytest.log; tail -3 gpurun_out/r02h_pytest.log
for mode in 0 1 0 1; do
ICP_B200_H2D_AHEAD=$mode timeout 600 python bench.py --steps 20 --warmup 4 --headline-only --no-cpu-baseline > gpurun_out/r02h_bench_ahead$mode.log 2> gpurun_out/r02h_bench_ahead$mode.err
python - $mode <<'PY'
import json, sys
mode = sys.argv[1]
d=json.loads([l for l in open(f'gpurun_out/r02h_bench_ahead{mode}.log') if l.startswith('{')][-1])
e=d['e2e']
print('ahead', mode, 'value', round(d['value']), 'e2e two', round(e['value']), 'one', round(e['one_caller']['value']), 'pipelined', round(e['one_caller_pipelined']['value']), e['bit_identical_to_device_path'], e['one_caller_pipelined']['bit_identical_to_device_path'])
PY
done
