#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tools/h2d_contention.py > gpurun_out/r02_h2d_contention_n$N.log 2> gpurun_out/r02_h2d_contention_n$N.err
cat gpurun_out/r02_h2d_contention_n$N.log
nvidia-smi topo -m 2>&1 | head -30 >> gpurun_out/r02_h2d_contention_n$N.log
nvidia-smi topo -m 2>&1 | head -14 | cut -c1-200
tail -3 gpurun_out/r02_h2d_contention_n$N.err | cut -c1-300
