#!/bin/bash
# round 2, job W: config 4 timing + ncu of the streaming kernel; config 3 on a 4-CTA/SM lean instance
mkdir -p gpurun_out
timeout 300 python tools/time_config4.py 1024 16 | cut -c1-500
python - <<'PY'
import os, sys
sys.path.insert(0,'.')
import bench, icp_rust_b200 as icp
for inst in (None, "14"):
    if inst: os.environ["ICP_B200_FORCE_INSTANCE"] = inst
    ctx = icp.Context(0)
    c3 = bench.bench_config3(icp, ctx, scans_n=32768)
    info = ctx.call_info()
    print('c3 instance', inst, c3['value'], c3['status_ok'], info['grid_ctas'], info['block_threads'], info['smem_bytes'])
    ctx.close()
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:icp_stream -s 2 -c 1 -o gpurun_out/r02v_stream python tools/time_config4.py 296 0 > gpurun_out/r02v_ncu_stream.log 2>&1
tail -2 gpurun_out/r02v_ncu_stream.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
timeout 600 python -m pytest tests -m gpu -q -x -k "config4 or 3d or streaming or fixed_point" 2>&1 | tail -2
