#!/bin/bash
# round 2, GPU job E: lean instance A/B (256x6x3 SVC vs 384x4x2 SVC), config-4 grid sweep, stream kernel at 2 CTAs/SM
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r02e_pytest.log
tail -3 gpurun_out/r02e_pytest.log
for inst in 2 3; do
  ICP_B200_FORCE_INSTANCE=$inst timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --headline-only > gpurun_out/r02e_bench_inst$inst.json 2> gpurun_out/r02e_bench_inst$inst.err
  python -c "
import json; d=json.load(open('gpurun_out/r02e_bench_inst$inst.json')); print('inst $inst', d['value'], d['e2e']['value'], d['kernel_shape'])"
done
for cpp in 8 32 128 512; do
  ICP_B200_STREAM_CELLS_PER_POINT=$cpp timeout 300 python tools/time_config4.py 296 8 >> gpurun_out/r02e_config4_sweep.jsonl 2>> gpurun_out/r02e_config4_sweep.err
done
for cpp in 32 128; do
  ICP_B200_LIB_DIR=$PWD/icp_rust_b200/lib_exp ICP_B200_STREAM_CELLS_PER_POINT=$cpp timeout 300 python tools/time_config4.py 296 8 >> gpurun_out/r02e_config4_sweep_2cta.jsonl 2>> gpurun_out/r02e_config4_sweep.err
done
cat gpurun_out/r02e_config4_sweep.jsonl gpurun_out/r02e_config4_sweep_2cta.jsonl | cut -c1-330
