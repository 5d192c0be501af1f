#!/bin/bash
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests -m gpu -x -q -k "test_small_instances or test_l_shape or test_status_codes or test_final_cost or test_multi_hypothesis" > gpurun_out/r02p_racecheck.log 2>&1
tail -3 gpurun_out/r02p_racecheck.log
python - <<'PY'
import sys, torch
sys.path.insert(0,'.')
import bench, icp_rust_b200 as icp
dev=torch.device('cuda',0); ctx=icp.Context(0); st=torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
r=bench.bench_nn_search(icp, ctx, torch, dev, st)
print('nn', r['value'], r['jittered_map_points']['value'], r['parity_sample_ok'])
PY
timeout 300 python tools/time_config4.py 592 8 | cut -c1-400
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
