#!/bin/bash
# tests/tools/shard_matrix.sh "<N list>" "<shift list>" [extra args]: bench_sharded.py for each pair; one summary line each
for n in $1; do for sh in $2; do
  tag=n${n}_s${sh}
  if [ $n = 1 ]; then cmd="python bench_sharded.py"; else cmd="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n * 10 + sh)) bench_sharded.py"; fi
  timeout 400 $cmd --shift $sh $3 > gpurun_out/shard_$tag.json 2> gpurun_out/shard_$tag.err || echo "$tag FAILED rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/shard_$tag.json").read().strip().splitlines()[-1])
    v = d.get("verify") or {}
    print("$tag", round(d["value"]), "fps", round(d["ms_per_step"] * 1e3), "us/step", d["per_rank_us"]["rows"], "verify_ok", v.get("ok"), v.get("hit_mask_mismatch"), v.get("depth_over_half_vox"))
except Exception as e:
    print("$tag", "no result", e)
PY
done; done
