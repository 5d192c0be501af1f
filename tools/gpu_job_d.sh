#!/bin/bash
# round 2, GPU job D: ncu --set full of the lean kernel (592 pairs) and of the NN kernel (paired loads), NN timing
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:icp_lean -s 1 -c 1 -o gpurun_out/r02d_lean python tools/profile_fused.py 592 2 > gpurun_out/r02d_ncu_lean.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nn_search -s 2 -c 1 -o gpurun_out/r02d_nn python tools/profile_nn.py > gpurun_out/r02d_ncu_nn.log 2>&1
timeout 300 python - > gpurun_out/r02d_nn_timing.log 2>&1 <<'PY'
import sys, torch
sys.path.insert(0,'.')
import bench, icp_rust_b200 as icp
dev=torch.device('cuda',0); ctx=icp.Context(0); st=torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
r=bench.bench_nn_search(icp, ctx, torch, dev, st)
print('nn', r['value'], r['jittered_map_points']['value'], r['parity_sample_ok'])
PY
cat gpurun_out/r02d_nn_timing.log | tail -2
ls -la gpurun_out/*.ncu-rep
