#!/bin/bash
# round 2, final records: GPU tests, bench both arms, launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02f_pytest.log; tail -2 gpurun_out/r02f_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > gpurun_out/r02f_clocks.csv &
SMI=$!
timeout 900 python bench.py > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err
kill $SMI
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02f_bench_ref.json 2> gpurun_out/r02f_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --headline-only > gpurun_out/r02f_ncu_launches.log 2>&1
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02f_bench_n1.json') if l.startswith('{')][-1])
print('value',d['value'],'e2e',d['e2e']['value'],d['e2e']['one_caller']['value'],'ordered',d['ordered_sums_mode']['value'], 'roofline', d['roofline']['frac'], d['roofline']['fp64']['frac'])
print(' overlap', d['step_overlap']['steps_one_after_the_other'], ' full', d['fixed_point_exit']['every_iteration_computed'])
print(' nn',d['nn_search']['value'],d['nn_search']['jittered_map_points']['value'],'c3',d['config3_shared_map']['value'],'c4',d['config4_3d']['value'], d['config4_3d']['every_iteration_computed'])
print(' census',d['census'],'cpu',d['cpu_baseline']['value'], d['cpu_baseline'].get('single_pair_latency'))
print(' launches', d['gpu_launches'], d['clocks'])
PY
cat gpurun_out/r02f_bench_ref.json | cut -c1-300
tail -2 gpurun_out/r02f_bench_n1.err
