#!/bin/bash
# round 2, job G: the group search kernel (4 / 8 lanes per query) against the private walk
mkdir -p gpurun_out
for g in 0 4 8; do
ICP_B200_NN_GROUP=$g python - <<'PY'
import os, sys, torch
sys.path.insert(0,'.')
import bench, icp_rust_b200 as icp
dev=torch.device('cuda',0); ctx=icp.Context(0); st=torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
r=bench.bench_nn_search(icp, ctx, torch, dev, st)
print('group', os.environ["ICP_B200_NN_GROUP"], 'nn', r['value'], r['jittered_map_points']['value'], r['parity_sample_ok'])
PY
done
ICP_B200_NN_GROUP=4 timeout 600 python -m pytest tests -m gpu -q -x -k "search" 2>&1 | tail -2
ICP_B200_NN_GROUP=8 timeout 600 python -m pytest tests -m gpu -q -x -k "search" 2>&1 | tail -2
