#!/bin/bash
# round 2, job Y: 2-D shared-target walk with two candidate loads in flight (ICP_NN_BATCH4), NN kernel at 2 / 4 blocks per SM
mkdir -p gpurun_out
for d in lib lib_exp lib_exp2; do
ICP_B200_LIB_DIR=$PWD/icp_rust_b200/$d python - <<'PY'
import os, sys, torch
sys.path.insert(0,'.')
import bench, icp_rust_b200 as icp
dev=torch.device('cuda',0); ctx=icp.Context(0); st=torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
r=bench.bench_nn_search(icp, ctx, torch, dev, st)
c3=bench.bench_config3(icp, ctx, scans_n=32768)
print(os.environ["ICP_B200_LIB_DIR"].split('/')[-1], 'nn', r['value'], r['jittered_map_points']['value'], r['parity_sample_ok'], 'c3', c3['value'], c3['status_ok'])
PY
done
