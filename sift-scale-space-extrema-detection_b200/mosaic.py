"""Gigapixel mosaics: horizontal strips across GPUs with one seed-halo exchange per octave.

The reference blurs every level of an octave from that octave's base image only (background.js:185-190), so
the only dependency between strips is the base (seed) image's halo: per octave `max radius + margin` rows
from each neighbour.  Each strip is one `Engine` on one GPU:

    layout = plan_strips(...)[rank]
    engine.strip_begin(params, layout, source_rows)           # upload the strip's rows of the image (+ halo)
    for o in range(octaves):
        if o > 0: exchange_seed_halos(...)                    # NCCL send/recv of seed rows, device to device
        engine.strip_octave(o)                                # blur + DoG of octave o, owned seed rows of o+1
    keypoints = engine.strip_finish()                         # scan + refine of the owned rows, global coordinates

Per-pixel arithmetic does not depend on the decomposition (same taps, same order), so the union of the
strips' keypoints is bit-identical to the whole-image result (tests/test_mosaic.py).  A refinement walk may
jump to a row its strip does not hold (background.js:638-640 moves by Math.round(alpha), unbounded): the strip
hands the walk's state out (`strip_escaped`), the strip that owns the new row continues it (`strip_resume`),
until every walk has ended (`resolve_escaped_*`; at most maxIterations - 1 rounds).

`detect_mosaic_local` drives several engines from one process (tests; or all GPUs of a box from one host
thread); `detect_mosaic_distributed` is the one-process-per-GPU form on `torch.distributed` (NCCL over
NVLink on the GPU box, gloo for the CPU tests of the exchange logic).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


def strip_layout(params: L.Params, full_width: int, full_height: int, row0: int, row1: int, margin: int = 16) -> L.StripLayout:
    """Host-only arithmetic (works without a GPU): rows of the global octave grids one strip owns / holds."""
    out = L.StripLayout()
    rc = L.load().sift_strip_layout_compute(C.byref(params), full_width, full_height, row0, row1, margin, C.byref(out))
    if rc != L.SIFT_OK:
        raise L.SiftError(rc, (L.load().sift_last_error(None) or b"").decode())
    return out


def plan_strips(params: L.Params, full_width: int, full_height: int, world: int, margin: int = 16) -> list:
    """Cut the octave-0 grid (2 * full_height rows) into `world` strips of near-equal height whose boundaries
    are multiples of 2^(octaves-1), so that owned rows halve exactly from octave to octave and the decimation
    in[2a][2b] (matrix2d.js:129) lines up across strips."""
    h0 = 2 * full_height
    align = 1 << (params.numberOfOctaves - 1)
    units = h0 // align
    if units < world:
        raise ValueError(f"{h0} octave-0 rows cannot be cut into {world} strips aligned to {align}")
    bounds = [(units * r // world) * align for r in range(world)] + [h0]
    return [strip_layout(params, full_width, full_height, bounds[r], bounds[r + 1], margin) for r in range(world)]


def halo_transfers(layouts: list, octave: int) -> list:
    """Seed rows that must move before `octave` can be blurred: (src_rank, dst_rank, first_global_row, n_rows).
    A strip's halo rows are owned by its direct neighbours (strip_layout refuses thinner strips)."""
    out = []
    for r, lay in enumerate(layouts):
        top, own0, own1, bot = lay.top[octave], lay.own0[octave], lay.own1[octave], lay.bottom[octave]
        if own0 > top:
            assert r > 0 and layouts[r - 1].own0[octave] <= top and layouts[r - 1].own1[octave] == own0
            out.append((r - 1, r, top, own0 - top))
        if bot > own1:
            assert r + 1 < len(layouts) and layouts[r + 1].own1[octave] >= bot and layouts[r + 1].own0[octave] == own1
            out.append((r + 1, r, own1, bot - own1))
    return out


def exchange_seed_halos(seed, layouts: list, octave: int, rank: int, dist=None):
    """One-process-per-strip exchange.  `seed` is this rank's seed image of `octave` as a 2-D torch tensor
    (rows [top, bottom) x width, any device); halo rows are received in place, owned rows are sent."""
    import torch
    import torch.distributed as tdist
    dist = dist or tdist
    top = layouts[rank].top[octave]
    ops, keep = [], []
    for src, dst, row, n in halo_transfers(layouts, octave):
        if src == rank:
            buf = seed[row - top:row - top + n].contiguous()
            keep.append(buf)
            ops.append(dist.P2POp(dist.isend, buf, dst))
        elif dst == rank:
            ops.append(dist.P2POp(dist.irecv, seed[row - top:row - top + n], src))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    if seed.is_cuda:
        torch.cuda.synchronize(seed.device)
    return seed


class _DeviceRows:
    """__cuda_array_interface__ view of engine-owned device memory (so torch / NCCL work on it in place)."""

    def __init__(self, ptr: int, rows: int, cols: int):
        self.__cuda_array_interface__ = {"shape": (rows, cols), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def seed_tensor(engine, layout: L.StripLayout, octave: int):
    """The strip's fp64 seed image of `octave` as a torch tensor aliasing the engine's device memory."""
    import torch
    rows = layout.bottom[octave] - layout.top[octave]
    ptr = engine.strip_seed(octave)
    return torch.as_tensor(_DeviceRows(ptr, rows, layout.width[octave]), device=torch.device("cuda", engine.device))


def source_rows(image: np.ndarray, layout: L.StripLayout) -> np.ndarray:
    """Rows of the full source image a strip uploads: its owned rows plus the octave-0 halo."""
    return np.ascontiguousarray(image[layout.top[0] // 2:(layout.bottom[0] + 1) // 2])


def exchange_local(engines: list, octave: int):
    """sift_mosaic_exchange: the halo rows of every strip's seed of `octave`, device to device, inside the C ABI
    (cudaMemcpyPeerAsync, stream-ordered, the host does not wait) -- what a host without torch calls."""
    lib = L.load()
    arr = (C.c_void_p * len(engines))(*[e.handle for e in engines])
    rc = lib.sift_mosaic_exchange(arr, len(engines), octave)
    if rc != L.SIFT_OK:
        raise L.SiftError(rc, (lib.sift_last_error(engines[0].handle) or b"").decode())


def detect_mosaic_local(engines: list, image: np.ndarray, params: L.Params, margin: int = 16, layouts: list | None = None,
                        exchange: str = "c_abi"):
    """All strips from one process (engine i = strip i; the engines may sit on different GPUs or share one).
    exchange = "c_abi": halo rows move with sift_mosaic_exchange (peer copies inside the library); "torch": with
    torch tensor copies over the seed images (the form the distributed path uses).  `layouts`: explicit cuts
    (default: plan_strips).  Returns (keypoints in reference order, per-strip stats, layouts)."""
    h, w = image.shape[:2]
    layouts = layouts or plan_strips(params, w, h, len(engines), margin)
    for eng, lay in zip(engines, layouts):
        eng.strip_begin(params, lay, source_rows(image, lay))
    for o in range(params.numberOfOctaves):
        if o > 0 and exchange == "c_abi":
            exchange_local(engines, o)
        elif o > 0:
            import torch
            seeds = [seed_tensor(e, lay, o) for e, lay in zip(engines, layouts)]
            for src, dst, row, n in halo_transfers(layouts, o):
                s0 = row - layouts[src].top[o]
                d0 = row - layouts[dst].top[o]
                seeds[dst][d0:d0 + n].copy_(seeds[src][s0:s0 + n])
            torch.cuda.synchronize()
        for eng in engines:
            eng.strip_octave(o)
    parts, stats = [], []
    for eng in engines:
        k, st = eng.strip_finish()
        parts.append(k)
        stats.append(st)
    resolve_escaped_local(engines, layouts, parts, stats)
    return merge_keypoints(parts), stats, layouts


_OUTCOMES = ("keypoints", "rejLowContrast", "rejEdge", "rejLeftScale", "rejLeftRows", "rejLeftCols", "rejNoConvergence",
             "rejSingular")


def owner_of(layouts: list, octave: int, row: int) -> int:
    """The strip that owns global row `row` of `octave` (owned ranges tile the octave's rows)."""
    for r, lay in enumerate(layouts):
        if lay.own0[octave] <= row < lay.own1[octave]:
            return r
    raise ValueError(f"row {row} of octave {octave} is owned by no strip")


def _add_outcomes(total: dict, part: dict):
    for k in _OUTCOMES:
        total[k] += part[k]
    total["kernelLaunches"] += part["kernelLaunches"]


def resolve_escaped_local(engines: list, layouts: list, parts: list, stats: list, max_rounds: int = 8):
    """Continue, on the owning strips, the refinement walks that left a strip (all engines in this process).
    parts[i] / stats[i] are extended in place with the keypoints / outcomes decided by strip i."""
    pending = [e.strip_escaped() for e in engines]
    for _ in range(max_rounds):
        walks = np.concatenate(pending) if pending else np.zeros(0, dtype=L.WALK_DTYPE)
        if len(walks) == 0:
            return
        owners = np.array([owner_of(layouts, int(w["octave"]), int(w["y"])) for w in walks])
        pending = []
        for r, eng in enumerate(engines):
            mine = walks[owners == r]
            if len(mine) == 0:
                continue
            k, st = eng.strip_resume(mine)
            if len(k):
                parts[r] = np.concatenate([parts[r], k])
            _add_outcomes(stats[r], st)
            pending.append(eng.strip_escaped())
    raise RuntimeError("refinement walks still unresolved after %d rounds" % max_rounds)


def resolve_escaped_distributed(engine, layouts: list, rank: int, kps: np.ndarray, stats: dict, dist=None,
                                max_rounds: int = 8):
    """One process per strip: every round all ranks share their escaped walks (a few 40-byte records, padded
    tensors over the process group), each continues the ones whose row it owns.  Returns this rank's keypoints
    with the newly decided ones appended."""
    import torch
    import torch.distributed as tdist
    dist = dist or tdist
    world = dist.get_world_size()
    dev = torch.device("cuda", engine.device) if dist.get_backend() == "nccl" else torch.device("cpu")
    rec = L.WALK_DTYPE.itemsize
    mine_out = engine.strip_escaped()
    extra = []
    for _ in range(max_rounds):
        counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([len(mine_out)], dtype=torch.int64, device=dev))
        counts = [int(c.item()) for c in counts]
        if sum(counts) == 0:                                  # the common case: nobody has a walk to hand over
            return np.concatenate([kps] + extra) if extra else kps
        width = max(counts) * rec
        send = torch.zeros(width, dtype=torch.uint8)
        send[:len(mine_out) * rec] = torch.from_numpy(np.frombuffer(mine_out.tobytes(), dtype=np.uint8).copy())
        bufs = [torch.zeros(width, dtype=torch.uint8, device=dev) for _ in range(world)]
        dist.all_gather(bufs, send.to(dev))
        walks = np.concatenate([np.frombuffer(b.cpu().numpy().tobytes()[:n * rec], dtype=L.WALK_DTYPE)
                                for b, n in zip(bufs, counts)])
        owners = np.array([owner_of(layouts, int(w["octave"]), int(w["y"])) for w in walks])
        mine = walks[owners == rank]
        mine_out = np.zeros(0, dtype=L.WALK_DTYPE)
        if len(mine):
            k, st = engine.strip_resume(mine)
            if len(k):                                        # (appending copies the strip's whole record array)
                extra.append(k)
            _add_outcomes(stats, st)
            mine_out = engine.strip_escaped()
    raise RuntimeError("refinement walks still unresolved after %d rounds" % max_rounds)


def detect_mosaic_distributed(engine, image_rows: np.ndarray, layouts: list, params: L.Params, rank: int):
    """One process per GPU (torch.distributed already initialised, NCCL): `image_rows` are this rank's
    source_rows().  Returns this strip's keypoints (global coordinates) and stats; gather them with
    `gather_keypoints`."""
    lay = layouts[rank]
    engine.strip_begin(params, lay, image_rows)
    for o in range(params.numberOfOctaves):
        if o > 0:
            exchange_seed_halos(seed_tensor(engine, lay, o), layouts, o, rank)
        engine.strip_octave(o)
    kps, stats = engine.strip_finish()
    kps = resolve_escaped_distributed(engine, layouts, rank, kps, stats)
    return kps, stats


def merge_keypoints(parts: list) -> np.ndarray:
    """Strips own disjoint rows: the union, put back into the reference's order (octave, scale, y, x)."""
    allk = np.concatenate(parts) if parts else np.zeros(0, dtype=L.KEYPOINT_DTYPE)
    order = np.lexsort((allk["candX"], allk["candY"], allk["candScale"], allk["octave"]))
    return allk[order]


def gather_keypoints(local: np.ndarray, rank: int, world: int, dist=None) -> np.ndarray | None:
    """Rank 0 receives every strip's records (bytes over the process group) and merges them."""
    import torch.distributed as tdist
    dist = dist or tdist
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(local.tobytes(), gathered, dst=0)
    if rank != 0:
        return None
    return merge_keypoints([np.frombuffer(b, dtype=L.KEYPOINT_DTYPE) for b in gathered])
