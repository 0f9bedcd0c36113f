#!/usr/bin/env python
"""offline_eval.py -- the reference's offline pipeline (main/offline_eval.cc:37-99, examples/scannet_evaluation/eval_one.cc)
on this engine: read a folder dataset, integrate every frame from the raw 16-bit images, optionally raycast along the way,
then dump what the evaluation scripts read back:

    <out>/tsdf.bin                       N x 5 float32 (x, y, z, tsdf, prob)   -- eval_one.cc:82-86
    <out>/mesh_vertices.bin, mesh_indices.bin, mesh_vertices_prob.bin           -- TSDFSystem::DownloadAllMesh

    python examples/offline_eval.py --dataset DIR --out DIR [--voxel 0.01 --truncation 0.06 --max-depth 4]
    python examples/offline_eval.py --synth 40 --dataset /tmp/ds --out /tmp/out      # write a synthetic dataset first

Needs a B200 (there is no CPU fallback)."""
import argparse
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dataset", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--voxel", type=float, default=0.01)
    ap.add_argument("--truncation", type=float, default=0.06)
    ap.add_argument("--max-depth", type=float, default=4.0)
    ap.add_argument("--render-every", type=int, default=0, help="raycast every n-th frame from its own pose (0 = never)")
    ap.add_argument("--synth", type=int, default=0, help="first write an n-frame synthetic dataset (config 2 shape) to --dataset")
    args = ap.parse_args()
    rsf = importlib.import_module("ra-slam_b200")
    fmt = importlib.import_module("ra-slam_b200.formats")
    if args.synth:
        synth = importlib.import_module("ra-slam_b200.synth")
        seq = synth.config2(n_frames=max(args.synth, 300))
        fmt.FolderDataset.write(args.dataset, seq.K, 5000.0, (seq.frame(i) for i in range(args.synth)), seq.poses[:args.synth])
    ds = fmt.FolderDataset(args.dataset)
    grid = rsf.TSDFGrid(args.voxel, args.truncation)
    t0 = time.perf_counter()
    rendered = 0
    for i in range(len(ds)):
        rgb, depth, ht, lt = ds.raw(i)
        grid.IntegrateU16(rgb, depth, ht, lt, ds.depth_factor, args.max_depth, ds.intrinsics, ds.pose(i))
        if args.render_every and i % args.render_every == 0:
            grid.RayCast(args.max_depth * 2, rsf.CameraParams(ds.intrinsics, depth.shape[0], depth.shape[1]), ds.pose(i))
            rendered += 1
    t1 = time.perf_counter()
    os.makedirs(args.out, exist_ok=True)
    rows = grid.GatherValidSemantic()
    fmt.write_semantic_bin(os.path.join(args.out, "tsdf.bin"), rows)
    v, t, p = grid.GatherValidMesh()
    fmt.write_mesh_bins(args.out, v, t, p)
    print(json.dumps({"frames": len(ds), "rendered": rendered, "integrate_loop_s": t1 - t0, "active_blocks": grid.NumActiveBlock(),
                      "tsdf_rows": int(len(rows)), "vertices": int(len(v)), "triangles": int(len(t)), "out": args.out}))


if __name__ == "__main__":
    main()
