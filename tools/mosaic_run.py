#!/usr/bin/env python
"""Mosaic strips across GPUs (one process per GPU, NCCL seed-halo exchange per octave); rank 0 checks the
union against the whole-image result when --verify is given.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tools/mosaic_run.py --size 4096 --octaves 4 --verify
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import sift_b200  # noqa: E402
from sift_b200 import _lib as L, fixtures, mosaic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=4096)
    ap.add_argument("--octaves", type=int, default=4)
    ap.add_argument("--margin", type=int, default=64)
    ap.add_argument("--verify", action="store_true")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--phases", action="store_true", help="time upload / octaves (exchange + blur) / scan+refine separately")
    ap.add_argument("--pinned", action="store_true", help="keep the strip's source rows in pinned host memory")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W = H = args.size
    prm = L.default_params(numberOfOctaves=args.octaves, minBlurLevel=1.6)
    layouts = mosaic.plan_strips(prm, W, H, world, args.margin)
    lay = layouts[rank]
    # every rank generates the source rows of its own strip (+ octave-0 halo) of the same seeded mosaic: the
    # generator is counter-based, nobody holds the whole 32768^2 image
    nblobs = max(64, W * H // 16384)
    t_gen = time.perf_counter()
    rows = fixtures.synthetic_u8_rows(W, H, lay.top[0] // 2, (lay.bottom[0] + 1) // 2, 4321, blobs=nblobs)
    t_gen = time.perf_counter() - t_gen
    if args.pinned:
        pin = torch.from_numpy(rows).pin_memory()
        rows = pin.numpy()
    eng = sift_b200.Engine(local)
    phases = {}
    if args.phases:                                      # the same steps as mosaic.detect_mosaic_distributed, timed one by one
        def tick(name, t0):
            torch.cuda.synchronize(); dist.barrier()
            phases[name] = round(time.perf_counter() - t0, 4)
        for _ in range(2):
            dist.barrier(); t0 = time.perf_counter()
            eng.strip_begin(prm, lay, rows); tick("upload", t0)
            for o in range(args.octaves):
                t0 = time.perf_counter()
                if o > 0:
                    mosaic.exchange_seed_halos(mosaic.seed_tensor(eng, lay, o), layouts, o, rank)
                    tick(f"exchange_{o}", t0); t0 = time.perf_counter()
                eng.strip_octave(o); tick(f"octave_{o}", t0)
            t0 = time.perf_counter()
            k, st = eng.strip_finish(); tick("scan_refine_download_order", t0)
            t0 = time.perf_counter()
            mosaic.resolve_escaped_distributed(eng, layouts, rank, k, st); tick("walk_handover", t0)
    times = []
    for _ in range(args.reps):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        kps, stats = mosaic.detect_mosaic_distributed(eng, rows, layouts, prm, rank)
        torch.cuda.synchronize(); dist.barrier()
        times.append(time.perf_counter() - t0)
    merged = mosaic.gather_keypoints(kps, rank, world)
    left = torch.tensor([stats["leftStrip"]], device="cuda")
    dist.all_reduce(left)
    if rank == 0:
        line = {"mosaic": f"{W}x{H}", "octaves": args.octaves, "n_gpus": world, "strips": world, "margin": args.margin,
                "keypoints": int(len(merged)), "walks_handed_over": int(left.item()), "seconds": min(times),
                "mpixel_per_s": W * H / 1e6 / min(times),
                "halo_rows": [int(lay.halo[o]) for o in range(args.octaves)], "strip_source_rows": int(rows.shape[0]),
                "generate_s": round(t_gen, 2), "reps_s": [round(t, 4) for t in times]}
        if phases:
            line["phases_s"] = phases
        if args.verify:
            img = fixtures.synthetic_u8(W, H, 4321, blobs=nblobs)
            whole, _ = eng.detect(img, prm)
            line["identical_to_whole_image"] = bool(whole.tobytes() == merged.tobytes())
            line["whole_keypoints"] = int(len(whole))
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    eng.close()


if __name__ == "__main__":
    main()
