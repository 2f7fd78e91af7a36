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


def test_mosaic_module_surface():
    for name in ("strip_layout", "plan_strips", "halo_transfers", "exchange_seed_halos", "exchange_local", "seed_tensor", "source_rows",
                 "detect_mosaic_local", "detect_mosaic_distributed", "owner_of", "resolve_escaped_local",
                 "resolve_escaped_distributed", "merge_keypoints", "gather_keypoints"):
        assert callable(getattr(mosaic, name)), name


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
        # hand-over of refinement walks that left a strip: the protocol, with a stand-in for the GPU engine.  Every
        # walk bounces to the other strip until its third iteration, then becomes a keypoint where it stands.
        class FakeStrip:
            def __init__(self, first):
                self.out = first

            def strip_escaped(self):
                return self.out

            def strip_resume(self, walks):
                assert all(M.owner_of(lays, int(w["octave"]), int(w["y"])) == rank for w in walks)
                go = walks[walks["iteration"] < 3].copy()
                go["iteration"] += 1
                go["y"] = [lays[1 - rank].own0[int(o)] + 3 for o in go["octave"]]          # a row the other strip owns
                self.out = go
                end = walks[walks["iteration"] >= 3]
                k = np.zeros(len(end), dtype=L2.KEYPOINT_DTYPE)
                k["octave"], k["candY"], k["localY"], k["iterations"] = end["octave"], end["candY"], end["y"], end["iteration"]
                st = {n: 0 for n in M._OUTCOMES}
                st.update(keypoints=len(k), kernelLaunches=1)
                return k, st
        first = np.zeros(3, dtype=L2.WALK_DTYPE)
        first["octave"] = [0, 1, 2]
        first["iteration"] = [1, 2, 3]
        first["candY"] = 100 * (rank + 1) + np.arange(3)
        first["y"] = [lays[1 - rank].own0[o] + 1 for o in (0, 1, 2)]                      # jumped into the other strip
        stats = {n: 0 for n in M._OUTCOMES}
        stats["kernelLaunches"] = 0
        mine = M.resolve_escaped_distributed(FakeStrip(first), lays, rank, np.zeros(0, dtype=L2.KEYPOINT_DTYPE), stats)
        # walks that started at iteration 1 / 2 / 3 end after 3 / 2 / 1 hops: on this / the other / this... strip
        ok = ok and stats["keypoints"] == len(mine) == 3 and all(mine["iterations"] == 3)
        ok = ok and all(M.owner_of(lays, int(k["octave"]), int(k["localY"])) == rank for k in mine)
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
        assert len(got) == len(whole) and len(whole) > 20
        assert got.tobytes() == whole.tobytes()
        assert sum(s["candidates"] for s in stats) == st_whole["candidates"]
        for k in mosaic._OUTCOMES:                                   # every walk ended exactly once, the same way
            assert sum(s[k] for s in stats) == st_whole[k], k
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
def test_c_abi_exchange_equals_the_torch_exchange(engine):
    """sift_mosaic_exchange (peer copies inside the library, what a host without torch calls) moves the same rows as
    the torch copies: identical records, and both equal the whole image."""
    w, h, n_oct = 160, 1200, 3
    u8 = fixtures.synthetic_u8(w, h, 31, blobs=w * h // 512, sigma_lo=1.0, sigma_hi=6.0)
    prm = _params(n_oct)
    whole, _ = engine.detect(u8, prm)
    engines = [sift_b200.Engine(0) for _ in range(3)]
    try:
        a, _, lays = mosaic.detect_mosaic_local(engines, u8, prm, margin=16, exchange="c_abi")
        b, _, _ = mosaic.detect_mosaic_local(engines, u8, prm, margin=16, exchange="torch")
        assert a.tobytes() == b.tobytes() == whole.tobytes() and len(whole) > 100
        with pytest.raises(sift_b200.SiftError):                 # strips must be at the octave being exchanged
            mosaic.exchange_local(engines, 1)
    finally:
        for e in engines:
            e.close()


@pytest.mark.gpu
def test_refinement_walks_that_leave_a_strip_are_continued_by_the_owner(engine):
    """background.js:638-640 moves a sample by Math.round(alpha), unbounded: with a thin margin some walks jump to
    rows their strip does not hold.  They are handed to the owning strip and end exactly as in the whole image."""
    w, h, n_oct, world = 192, 1536, 3, 2
    u8 = fixtures.synthetic_u8(w, h, 5, blobs=w * h // 64, sigma_lo=0.6, sigma_hi=2.0)
    prm = _params(n_oct)
    whole, st_whole = engine.detect(u8, prm)
    # cut the image right behind the candidate of the longest vertical walk, so that the walk ends in the other strip
    H0, cut0 = 2 * h, None
    for kp in whole[np.argsort(-np.abs(whole["localY"] - whole["candY"]))][:8]:
        o = int(kp["octave"])
        ya, yb = sorted((int(kp["candY"]), int(kp["localY"])))
        align = max(1, (1 << (n_oct - 1)) >> o)                  # octave-0 cuts are multiples of 2^(octaves-1)
        c = (ya // align + 1) * align
        if yb - c >= 5 and 300 <= (c << o) <= H0 - 300:
            cut0 = c << o
            break
    assert cut0 is not None, "fixture has no refinement walk of 6+ rows: pick another seed"
    lays = [mosaic.strip_layout(prm, w, h, 0, cut0, 2), mosaic.strip_layout(prm, w, h, cut0, H0, 2)]
    engines = [sift_b200.Engine(0) for _ in range(world)]
    try:
        got, stats, lays = mosaic.detect_mosaic_local(engines, u8, prm, layouts=lays)
        escaped = sum(s["leftStrip"] for s in stats)
        assert escaped >= 1
        assert got.tobytes() == whole.tobytes() and len(whole) > 1000
        for k in mosaic._OUTCOMES:
            assert sum(s[k] for s in stats) == st_whole[k], k
        # the mechanism, directly: a fresh walk (iteration 0 on the candidate's pixel) resumed on the owner gives the
        # whole-image record; on a strip that does not hold the row it comes straight back, untouched
        moved = whole[(whole["iterations"] >= 1) & (whole["octave"] == 0)][:50]
        assert len(moved) >= 10
        walks = np.zeros(len(moved), dtype=L.WALK_DTYPE)
        walks["octave"] = moved["octave"]; walks["scaleLevel"] = moved["candScale"]; walks["x"] = moved["candX"]
        walks["y"] = moved["candY"]; walks["value"] = moved["dogValue"]; walks["candScale"] = moved["candScale"]
        walks["candX"] = moved["candX"]; walks["candY"] = moved["candY"]
        owners = np.array([mosaic.owner_of(lays, 0, int(y)) for y in walks["y"]])
        for r in sorted(set(owners.tolist())):
            k, st = engines[r].strip_resume(walks[owners == r])
            later = engines[r].strip_escaped()
            done = moved[owners == r]
            keep = ~np.isin(done["candY"] * 100000 + done["candX"], later["candY"] * 100000 + later["candX"])
            assert k.tobytes() == done[keep].tobytes() and st["keypoints"] == len(k)
        far = 1 - int(owners[0])
        sel = (owners == owners[0]) & (np.abs(walks["y"] - lays[far].own0[0]) > 100) & (np.abs(walks["y"] - lays[far].own1[0]) > 100)
        k, st = engines[far].strip_resume(walks[sel])
        back = engines[far].strip_escaped()
        assert len(k) == 0 and st["leftStrip"] == len(back) == int(sel.sum()) > 0
        assert back.tobytes() == np.sort(walks[sel], order=["octave", "candScale", "candY", "candX"]).tobytes()
        with pytest.raises(sift_b200.SiftError):
            bad = walks[:1].copy(); bad["y"] = 10 ** 6
            engines[0].strip_resume(bad)
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


def test_strip_generator_equals_the_whole_image_rows():
    """tools/mosaic_run.py: each rank generates only the source rows of its strip (counter-based PRNG)."""
    from sift_b200 import fixtures
    w, h = 211, 157
    whole = fixtures.synthetic_u8(w, h, 4321, blobs=40)
    for y0, y1 in ((0, 157), (0, 1), (30, 97), (150, 157), (-5, 20), (140, 400)):
        got = fixtures.synthetic_u8_rows(w, h, y0, y1, 4321, blobs=40)
        assert np.array_equal(got, whole[max(0, y0):min(h, y1)])
