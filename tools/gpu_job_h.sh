#!/bin/bash
# round 2, job H: sanitizer on the late additions (small workloads), bench with the two-caller e2e leg
mkdir -p gpurun_out
timeout 600 python tools/sanitizer_small.py 2>&1 | tail -5
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitizer_small.py > gpurun_out/r02h_racecheck.log 2>&1
grep -E "RACECHECK SUMMARY|ok$|Error|error" gpurun_out/r02h_racecheck.log | tail -8
timeout 1500 compute-sanitizer --tool memcheck python tools/sanitizer_small.py > gpurun_out/r02h_memcheck.log 2>&1
grep -E "ERROR SUMMARY|ok$" gpurun_out/r02h_memcheck.log | tail -6
timeout 900 python bench.py --headline-only --no-cpu-baseline --steps 20 > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02h_bench.json') if l.startswith('{')][-1])
print('value',d['value'],'e2e',d['e2e']['value'],'two callers',d['e2e']['two_callers'])
PY
tail -2 gpurun_out/r02h_bench.err
