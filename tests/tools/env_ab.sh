#!/bin/bash
# dev tool: A/B an environment knob on one box: tests/tools/env_ab.sh VAR v1 v2 ...
var=$1; shift
for m in "$@"; do
  export $var=$m
  python bench.py --steps 300 --warmup 20 --no-cpu-baseline --no-secondary > gpurun_out/ab_env_$m.json 2> gpurun_out/ab_env_$m.err
  python - <<PY
import json
d = json.load(open("gpurun_out/ab_env_$m.json"))
print("$var=$m", round(d["value"]), round(d["e2e"]["value"]), {k: round(x["ms_per_launch"] * 1e3, 1) for k, x in d["kernels"].items()}, round(d["roofline"]["frac"], 3))
PY
done
