#!/bin/bash
# round 2, final records after the submit/wait form and the chunk-by-chunk guard re-runs:
# GPU tests, sanitizer on the new host paths, smoke, bench both arms, launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02g_pytest.log; tail -2 gpurun_out/r02g_pytest.log
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests -m gpu -x -q -k "submit_wait or knife_edge" > gpurun_out/r02g_racecheck.log 2>&1
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r02g_racecheck.log | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > gpurun_out/r02g_clocks.csv &
SMI=$!
timeout 900 python bench.py > gpurun_out/r02g_bench_n1.json 2> gpurun_out/r02g_bench_n1.err
kill $SMI
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02g_bench_ref.json 2> gpurun_out/r02g_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02g_ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --headline-only > gpurun_out/r02g_ncu_launches.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 4 --headline-only --no-cpu-baseline --as-rank 5 > gpurun_out/r02g_bench_asrank5.json 2> /dev/null
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02g_bench_n1.json') if l.startswith('{')][-1])
e=d['e2e']
print('value',round(d['value']),'e2e (one thread, submit/wait, 3 contexts)',round(e['value']),'three callers',round(e['three_callers']['value']),'one',round(e['one_caller']['value']),'ordered',round(d['ordered_sums_mode']['value']), 'roofline', d['roofline']['frac'], d['roofline']['fp64']['frac'])
print(' nn',d['nn_search']['value'],'c3',d['config3_shared_map']['value'],'c4',d['config4_3d']['value'])
print(' census',d['census'],'cpu',d['cpu_baseline']['value'], d['cpu_baseline'].get('single_pair_latency'))
print(' launches', d['gpu_launches'], d['clocks'])
d=json.loads([l for l in open('gpurun_out/r02g_bench_asrank5.json') if l.startswith('{')][-1]); e=d['e2e']
print('as rank 5: value',round(d['value']),'e2e',round(e['value']),'three callers',round(e['three_callers']['value']),'one',round(e['one_caller']['value']))
PY
cut -c1-300 gpurun_out/r02g_bench_ref.json
tail -2 gpurun_out/r02g_bench_n1.err
