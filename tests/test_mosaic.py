"""Mosaic strips + per-octave seed-halo exchange.

CPU (not gpu): the strip arithmetic (host-only C function) and the halo exchange itself on a world_size-2 gloo
process group.  GPU: tiling invariance -- the union of the strips' keypoints is bit-identical to the
whole-image result, and every strip level equals the same rows of the whole-image pyramid."""
import os
import socket
import sys

import numpy as np
import pytest

import sift_b200
from sift_b200 import _lib as L, fixtures, mosaic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _params(n_oct=3, min_blur=1.6):
    return L.default_params(numberOfOctaves=n_oct, minBlurLevel=min_blur)


def test_strip_layout_arithmetic():
    prm = _params(3)
    lays = mosaic.plan_strips(prm, 96, 320, 2, margin=8)       # octave 0: 192 x 640
    a, b = lays
    assert a.octaves == 3 and list(a.width[:3]) == [192, 96, 48] and list(a.height[:3]) == [640, 320, 160]
    assert (a.own0[0], a.own1[0], b.own0[0], b.own1[0]) == (0, 320, 320, 640)
    for o in range(3):
        assert a.own1[o] == b.own0[o] == 320 >> o                # owned rows halve exactly
        assert a.top[o] == 0 and b.bottom[o] == a.height[o]      # image borders: no halo
        # halo = max radius of the octave (15, 29, 58) + margin, even
        assert a.halo[o] == ([15, 29, 58][o] + 8 + 1) // 2 * 2
        assert a.bottom[o] == a.own1[o] + a.halo[o] and b.top[o] == b.own0[o] - b.halo[o]
        assert b.top[o] % 2 == 0
    # the transfers of octave 1: each strip receives its halo from the neighbour's owned rows
    tr = mosaic.halo_transfers(lays, 1)
    assert sorted(tr) == sorted([(1, 0, 160, a.halo[1]), (0, 1, 160 - b.halo[1], b.halo[1])])


def test_strip_layout_rejects_bad_cuts():
    prm = _params(4)
    with pytest.raises(sift_b200.SiftError):                     # boundary not a multiple of 2^(octaves-1)
        mosaic.strip_layout(prm, 64, 256, 0, 100)
    with pytest.raises(sift_b200.SiftError):                     # strip thinner than the octave-3 halo: gather instead
        mosaic.strip_layout(prm, 64, 256, 64, 128)
    one = mosaic.strip_layout(prm, 64, 256, 0, 512)              # a single strip = the whole image
    assert all(one.top[o] == 0 and one.bottom[o] == one.height[o] for o in range(4))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import sift_b200  # noqa: F401
    from sift_b200 import _lib as L2, mosaic as M
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        prm = L2.default_params(numberOfOctaves=3, minBlurLevel=1.6)
        lays = M.plan_strips(prm, 40, 400, world, margin=4)
        ok = True
        for o in (1, 2):
            lay = lays[rank]
            w = lay.width[o]
            top, bot, own0, own1 = lay.top[o], lay.bottom[o], lay.own0[o], lay.own1[o]
            seed = torch.full((bot - top, w), float("nan"), dtype=torch.float64)
            rows = torch.arange(top, bot, dtype=torch.float64)[:, None] * 1000 + torch.arange(w, dtype=torch.float64)[None, :]
            seed[own0 - top:own1 - top] = rows[own0 - top:own1 - top]          # only the owned rows are known
            M.exchange_seed_halos(seed, lays, o, rank)
            ok = ok and bool(torch.equal(seed, rows))                            # halos now hold the neighbours' rows
        kp = np.zeros(2, dtype=L2.KEYPOINT_DTYPE)
        kp["octave"] = rank
        kp["candY"] = [5 + rank, 1 + rank]
        merged = M.gather_keypoints(kp, rank, world)
        if rank == 0:
            ok = ok and len(merged) == 2 * world and list(merged["octave"]) == sorted(merged["octave"])
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_halo_exchange_on_gloo_world_size_2():
    """The N > 1 path on CPU: two processes, gloo, the same exchange_seed_halos / gather the GPUs run over NCCL."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results == {0: True, 1: True}


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("world,w,h,n_oct", [(2, 160, 384, 3), (3, 97, 600, 3), (2, 64, 1100, 4)])
def test_strips_are_bit_identical_to_the_whole_image(engine, world, w, h, n_oct):
    """Tiling invariance: same taps in the same order -> identical levels in the owned rows, identical keypoints."""
    u8 = fixtures.synthetic_u8(w, h, 77, blobs=max(8, w * h // 1024), sigma_lo=1.0, sigma_hi=6.0)
    prm = _params(n_oct)
    whole, st_whole = engine.detect(u8, prm)
    levels = {(o, s): engine.get_level(L.SIFT_LEVEL_DOG, o, s) for o in range(n_oct) for s in range(5)}
    engines = [sift_b200.Engine(0) for _ in range(world)]
    try:
        got, stats, lays = mosaic.detect_mosaic_local(engines, u8, prm, margin=16)
        assert sum(s["leftStrip"] for s in stats) == 0
        assert len(got) == len(whole) and len(whole) > 20
        assert got.tobytes() == whole.tobytes()
        assert sum(s["candidates"] for s in stats) == st_whole["candidates"]
        for r, (eng, lay) in enumerate(zip(engines, lays)):
            for o in range(n_oct):
                for s in range(5):
                    d = eng.get_level(L.SIFT_LEVEL_DOG, o, s)
                    a, b = lay.own0[o] - lay.top[o], lay.own1[o] - lay.top[o]
                    assert np.array_equal(d[a:b], levels[(o, s)][lay.own0[o]:lay.own1[o]]), (r, o, s)
    finally:
        for e in engines:
            e.close()


@pytest.mark.gpu
def test_engine_returns_to_whole_images_after_a_strip(engine):
    u8 = fixtures.synthetic_u8(96, 400, 5)
    prm = _params(3)
    before, _ = engine.detect(u8, prm)
    lay = mosaic.plan_strips(prm, 96, 400, 2)[1]
    engine.strip_begin(prm, lay, mosaic.source_rows(u8, lay))
    with pytest.raises(sift_b200.SiftError):
        engine.strip_octave(1)                                     # octaves run in order
    after, _ = engine.detect(u8, prm)
    assert before.tobytes() == after.tobytes()
