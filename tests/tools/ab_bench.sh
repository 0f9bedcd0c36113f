#!/bin/bash
# A/B the default library against dev builds under build/alt/ on ONE box: tests/tools/ab_bench.sh main np p1 ...
for v in "$@"; do
  if [ "$v" = main ]; then unset RSF_LIBRARY; else export RSF_LIBRARY=$PWD/build/alt/librsfusion_$v.so; fi
  python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - <<PY
import json
d = json.load(open("gpurun_out/ab_$v.json"))
print("$v", round(d["value"]), round(d["e2e"]["value"]), {k: round(x["ms_per_launch"] * 1e3, 1) for k, x in d["kernels"].items()}, round(d["roofline"]["frac"], 3))
PY
done
