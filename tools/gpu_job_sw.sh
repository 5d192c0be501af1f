#!/bin/bash
# round 2, job SW: submit/wait form of the host-pointer call — its test, the whole GPU suite, the bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "submit_wait" 2>&1 | tail -15 > gpurun_out/r02sw_pytest_new.log; tail -15 gpurun_out/r02sw_pytest_new.log
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02sw_pytest.log; tail -3 gpurun_out/r02sw_pytest.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > gpurun_out/r02sw_clocks.csv &
SMI=$!
timeout 900 python bench.py > gpurun_out/r02sw_bench_n1.json 2> gpurun_out/r02sw_bench_n1.err
kill $SMI
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02sw_bench_n1.json') if l.startswith('{')][-1])
e=d['e2e']
print('value',d['value'],'e2e two callers',e['value'],'one',e['one_caller']['value'],'pipelined',e['one_caller_pipelined'])
print(' nn',d['nn_search']['value'],'c3',d['config3_shared_map']['value'],'c4',d['config4_3d']['value'])
print(' census',d['census'],'cpu',d['cpu_baseline']['value'])
print(' launches', d['gpu_launches'], d['clocks'])
PY
tail -3 gpurun_out/r02sw_bench_n1.err
