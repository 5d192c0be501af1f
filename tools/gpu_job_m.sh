#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02m_pytest.log; tail -2 gpurun_out/r02m_pytest.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > gpurun_out/r02m_clocks.csv &
SMI=$!
timeout 900 python bench.py > gpurun_out/r02m_bench_n1.json 2> gpurun_out/r02m_bench_n1.err
kill $SMI
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02m_bench_ref.json 2> gpurun_out/r02m_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02m_ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02m_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:icp_lean -s 1 -c 1 -o gpurun_out/r02m_lean python tools/profile_fused.py 888 2 > gpurun_out/r02m_ncu_lean.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nn_search -s 2 -c 2 -o gpurun_out/r02m_nn python tools/profile_nn.py > gpurun_out/r02m_ncu_nn.log 2>&1
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02m_bench_n1.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'ordered',d['ordered_sums_mode']['value'], 'roofline', d['roofline']['frac'], d['roofline']['fp64'])
print(' nn',d['nn_search']['value'],d['nn_search']['jittered_map_points']['value'],'c3',d['config3_shared_map']['value'],'c4',d['config4_3d']['value'])
print(' census',d['census'],'cpu',d['cpu_baseline'])
PY
cat gpurun_out/r02m_bench_ref.json | cut -c1-300
