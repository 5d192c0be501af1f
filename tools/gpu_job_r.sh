#!/bin/bash
# round 2, job R: outer-loop fixed-point exit + overlapped device calls: tests, bench, config 4
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "fixed_point or back_to_back or config4" 2>&1 | tail -3
timeout 900 python bench.py --steps 20 > gpurun_out/r02r_bench.json 2> gpurun_out/r02r_bench.err
python - <<'PY'
import json
l = [x for x in open("gpurun_out/r02r_bench.json") if x.startswith("{")][-1]
j = json.loads(l)
print("value", j["value"], "serial", j["step_overlap"]["steps_one_after_the_other"], "e2e", j["e2e"]["value"])
print("full", j["fixed_point_exit"]["every_iteration_computed"])
print("fp64", j["roofline"]["fp64"])
print("c3", j["config3_shared_map"])
c4 = j["config4_3d"]; print("c4", c4.get("value"), c4.get("outer_iterations_computed_per_pair"), c4.get("every_iteration_computed"), c4.get("oracle_parity_sample"))
print("nn", j["nn_search"]["value"]); print("cpu", j["cpu_baseline"]["value"], j["cpu_baseline"].get("single_pair_latency"))
PY
tail -3 gpurun_out/r02r_bench.err
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
