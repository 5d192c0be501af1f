#!/bin/bash
# round 2, GPU job C: dedicated service warps (inst 3 = 384 x 4, SVC) + warp-cooperative NN kernel A/B
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r02c_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
ICP_B200_NN_WARP=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_bench_nnprivate.json 2> gpurun_out/r02c_bench_nnprivate.err
ICP_B200_LIB_DIR=$PWD/icp_rust_b200/lib_timing timeout 300 python tools/profile_fused.py 592 2 > gpurun_out/r02c_lean_timing.log 2>&1
tail -5 gpurun_out/r02c_pytest.log
tail -3 gpurun_out/r02c_lean_timing.log
python - <<'PY'
import json
for f in ('gpurun_out/r02c_bench.json','gpurun_out/r02c_bench_nnprivate.json'):
    try:
        d=json.load(open(f))
        print(f,'value',d['value'],'e2e',d['e2e']['value'],'ordered',d['ordered_sums_mode']['value'])
        print(' nn',d['nn_search']['value'],d['nn_search']['jittered_map_points']['value'],d['nn_search']['parity_sample_ok'],'c3',d['config3_shared_map']['value'],'c4',d['config4_3d']['value'])
    except Exception as e: print(f,'ERR',e)
PY
