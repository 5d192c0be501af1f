#!/bin/bash
# round 2, job Q: back-to-back device calls with ICP_B200_FLAG_INDEPENDENT (programmatic dependent launch)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "back_to_back" 2>&1 | tail -3
timeout 600 python bench.py --headline-only --no-cpu-baseline --steps 20 > gpurun_out/r02q_bench_overlap.json 2> gpurun_out/r02q_bench_overlap.err
timeout 600 python bench.py --headline-only --no-cpu-baseline --steps 20 --no-overlap > gpurun_out/r02q_bench_serial.json 2> gpurun_out/r02q_bench_serial.err
python - <<'PY'
import json
for f in ("gpurun_out/r02q_bench_overlap.json", "gpurun_out/r02q_bench_serial.json"):
    try:
        l = [x for x in open(f) if x.startswith("{")][-1]
        j = json.loads(l)
        print(f, "value", j["value"], "serial", j["step_overlap"]["steps_one_after_the_other"]["value"], "e2e", j["e2e"]["value"], "launch_ms", j["roofline"]["launch_ms"])
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/r02q_bench_overlap.err
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
