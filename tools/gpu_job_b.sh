#!/bin/bash
# round 2, GPU job B: parity suite, bench line (new schedule of the lean kernel), reference arm, census, phase clocks
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02b_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r02b_bench_ref.json 2> gpurun_out/r02b_bench_ref.err
ICP_B200_LIB_DIR=$PWD/icp_rust_b200/lib_timing timeout 300 python tools/profile_fused.py 592 2 > gpurun_out/r02b_lean_timing.log 2>&1
timeout 900 python tools/census.py 131072 > gpurun_out/r02b_census.json 2> gpurun_out/r02b_census.err
tail -5 gpurun_out/r02b_pytest.log
cat gpurun_out/r02b_census.json
tail -3 gpurun_out/r02b_lean_timing.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02b_bench.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'ordered',d['ordered_sums_mode']['value'])
print('nn',d['nn_search']['value'],'c3',d['config3_shared_map'],'c4',d['config4_3d'])
print('census',d['census'],'guard',d['guard'])
PY
