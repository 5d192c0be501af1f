#!/bin/bash
# round 2, job T: stream kernel phase clock on config 4; NN kernel / config 3 with the shrinking disc
mkdir -p gpurun_out
ICP_B200_LIB_DIR=$PWD/icp_rust_b200/lib_timing timeout 300 python tools/time_config4.py 296 0 2>&1 | grep -v "^{" | tail -3
python - <<'PY'
import sys, torch
sys.path.insert(0,'.')
import bench, icp_rust_b200 as icp
dev=torch.device('cuda',0); ctx=icp.Context(0); st=torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
r=bench.bench_nn_search(icp, ctx, torch, dev, st)
print('nn', r['value'], r['jittered_map_points']['value'], r['parity_sample_ok'])
c3=bench.bench_config3(icp, ctx)
print('c3', c3['value'], c3['status_ok'], c3['redo_pairs'])
PY
timeout 900 python -m pytest tests -m gpu -q -x -k "search or config3 or shared" 2>&1 | tail -3
