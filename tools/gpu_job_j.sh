#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r02j_pytest.log; tail -2 gpurun_out/r02j_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02j_bench.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'ordered',d['ordered_sums_mode']['value'])
print(' nn',d['nn_search']['value'],'c3',d['config3_shared_map']['value'],'c4',d['config4_3d']['value'], d['config4_3d']['oracle_parity_sample'])
PY
timeout 600 python tools/census.py 65536 > gpurun_out/r02j_census.json 2> gpurun_out/r02j_census.err; cat gpurun_out/r02j_census.json
