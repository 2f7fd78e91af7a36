#!/usr/bin/env python
"""bench.py -- Mpixel/s of full SIFT detect (pyramid + DoG + extrema + refine).

Workload (BASELINE.json configs[1]): 1920x1080 synthetic frames, 4 octaves x 3 scales per octave
(6 blur levels), sigma0 = 1.6, assumed blur 0.5, the reference's thresholds, 2x-upsampled base octave
(always, as in the reference).  A step = one pass of the whole path over FRAMES frames per GPU.

  python bench.py --gpus N --steps K --warmup W            (own arm; torchrun for N > 1, weak scaling)
  python bench.py --impl reference --gpus N --steps K ...  (reference's CPU algorithm on the host cores)

Own arm, one JSON line on rank 0:
  value    device-resident throughput (inputs already in HBM), CUDA events on the engine's stream, max over ranks
  e2e      the same through sift_detect_batch() (C ABI, HOST buffers): H2D of the u8 frames, D2H and ordering of
           the keypoint records inside the timed region
  roofline dominant kernel class (octave-0 blur+DoG) against the measured HBM peak, plus the whole-path figure
  cpu_baseline  the float64 oracle (port of the reference's dense 2D algorithm) on one host core, bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "Mpixel/s full SIFT detect (pyramid+DoG+extrema+refine) at 1/2/4/8 B200"
W, H = 1920, 1080
N_OCT, SPO, MIN_BLUR, ASSUMED = 4, 3, 1.6, 0.5
FRAMES = 64                    # frames per GPU per step (distinct seeds)
# dram__bytes_read.sum + dram__bytes_write.sum per launch of each kernel, written by tools/ncu_traffic.py from the
# `ncu --set full` captures of this round (roofline.traffic is read from this file by kernel name, null if absent)
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "dram_traffic.json")
DOMINANT_KERNEL = "oct0_mma_kernel<1>"
# separable float64 blur of the reference radii: multiply-adds per input pixel (octave 0 polyphase-merged)
# and the measured fp64 peak of this B200 (DFMA 36.5, DMMA 37.0 TFLOP/s: tools/micro/fp64_pipes.cu) -- the binding unit of the path
FMA_PER_INPUT_PX = {4: 982.0, 6: 1048.0}
FP64_PEAK_TFMA = 18.5
LANES = int(os.environ.get("SIFT_B200_LANES", "0"))   # frames in flight per GPU (engine lanes); 0 = the engine picks by frame size
CPU_TILE = int(os.environ.get("SIFT_BENCH_CPU_TILE", "256"))   # cpu baseline / reference arm sample: CPU_TILE^2 crops of the same frames


def algorithmic_bytes_per_input_px(n_oct: int = N_OCT) -> dict:
    """SURVEY.md 8d, fp32 storage, every buffer touched the minimum number of times by the fused plan."""
    n = {0: 4.0}                                  # octave sizes in units of input pixels
    for o in range(1, n_oct):
        n[o] = n[o - 1] / 4.0
    blur0 = 4 * (1 + 6 * n[0] + 5 * n[0])         # read input, write 6 Gaussian + 5 DoG
    scan = {o: 4 * 5 * n[o] for o in range(n_oct)}
    blur = {0: blur0}
    for o in range(1, n_oct):
        blur[o] = 4 * (3 * n[o] + 10 * n[o])      # decimation read + seed write + seed read; 5 Gaussian + 5 DoG
    total = sum(blur.values()) + sum(scan.values())
    return {"blur": blur, "scan": scan, "total": total}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for i, nm in enumerate(names):
                if f[2 + i].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_frames(n: int, first_seed: int):
    from sift_b200 import fixtures
    return np.stack([fixtures.synthetic_u8(W, H, first_seed + i) for i in range(n)])


def ncu_traffic(kernel: str):
    """(bytes per launch, source) of `kernel` from profiles/dram_traffic.json, or (None, why)."""
    try:
        d = json.load(open(TRAFFIC_FILE))
        e = d["kernels"][kernel]
        return float(e["dram_bytes"]), f"{e['source']} ({d.get('how', 'ncu --set full')})"
    except Exception as ex:
        return None, f"no ncu capture on file for {kernel}: {ex}"


def bench_frames(eng, L, torch, frames_np, prm, n_frames, steps, warmup, cap, dist=None):
    """Device-resident and host-to-host throughput of `n_frames` frames per step drawn (cyclically) from frames_np
    [k, h, w] u8.  Returns a dict; times are max over ranks when dist is given."""
    k, h, w = frames_np.shape
    h_frames = torch.from_numpy(frames_np).pin_memory()
    d_frames = h_frames.cuda()
    ring = min(n_frames, 64)
    d_out = torch.zeros(ring, cap * L.KEYPOINT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    d_cnt = torch.zeros(n_frames, dtype=torch.int32, device="cuda")
    stream = torch.cuda.ExternalStream(eng.stream)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        for f in range(n_frames):
            eng.detect_device(d_frames[f % k].data_ptr(), L.SIFT_U8, w, h, 0, prm, d_out[f % ring].data_ptr(), cap,
                              d_cnt[f].data_ptr())

    for _ in range(warmup):
        step_device()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step_device()
    eng.flush()
    e1.record(stream)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    n_kp = int(d_cnt.clamp(min=0).sum().item())
    overflow = int((d_cnt < 0).sum().item())
    # host to host through sift_detect_batch: the k distinct frames repeated to n_frames
    reps = (n_frames + k - 1) // k
    h_out = torch.zeros(min(n_frames, k) * cap * L.KEYPOINT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    h_offs = torch.zeros(k + 1, dtype=torch.int32)

    def step_e2e():
        left = n_frames
        for _ in range(reps):
            m = min(left, k)
            eng.detect_batch_raw(h_frames.data_ptr(), L.SIFT_U8, w, h, 0, w * h, m, prm, h_out.data_ptr(), m * cap,
                                 h_offs.data_ptr())
            left -= m

    for _ in range(max(1, warmup - 1)):
        step_e2e()
    barrier()
    b0 = eng.transfer_bytes
    t0 = time.perf_counter()
    for _ in range(steps):
        step_e2e()
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], device="cuda")
    b1 = eng.transfer_bytes
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    del d_out, d_frames, h_out
    return {"ms_device": float(ms.item()) / steps, "s_e2e": float(t.item()) / steps, "keypoints_per_step": n_kp,
            "overflowed_frames": overflow, "h2d_bytes_per_step": (b1[0] - b0[0]) // steps,
            "d2h_bytes_per_step": (b1[1] - b0[1]) // steps}


def extra_configs(eng, L, torch, fixtures, rank, world, dist, peak):
    """BASELINE.json configs[0], [2], [3] (configs[1] is the headline line, configs[4] the mosaic entry): throughput in
    the headline's metric, as absolute Mpixel/s and as a fraction of the HBM roofline."""
    out = {}

    def entry(name, w, h, n_oct, per_rank, distinct, steps, scaling, note):
        prm = L.default_params(numberOfOctaves=n_oct, scalesPerOctave=SPO, minBlurLevel=MIN_BLUR, assumedBlur=ASSUMED)
        first = 1234 if scaling == "replicas" else 1234 + rank * per_rank          # cfg-3: contiguous blocks of the batch
        frames = np.stack([fixtures.synthetic_u8(w, h, first + i) for i in range(min(per_rank, distinct))])
        cap = max(1 << 13, (w * h) // 96)
        r = bench_frames(eng, L, torch, frames, prm, per_rank, steps, 2, cap, dist if scaling != "replicas" else None)
        ranks = 1 if scaling == "replicas" else world
        px = ranks * per_rank * w * h
        ab = algorithmic_bytes_per_input_px(n_oct)["total"]
        dev = px / 1e6 / (r["ms_device"] / 1e3)
        e2e = px / 1e6 / r["s_e2e"]
        out[name] = {"workload": f"{w}x{h}, {n_oct} octaves x s={SPO}, sigma0={MIN_BLUR}", "frames_per_gpu_per_step": per_rank,
                     "distinct_frames_per_gpu": int(frames.shape[0]), "n_gpus": ranks, "scaling": scaling,
                     "value": dev, "unit": "Mpixel/s", "e2e": e2e, "ms_per_frame_per_gpu": r["ms_device"] / per_rank,
                     "roofline_frac": ab * dev * 1e6 / ranks / 1e9 / peak, "algorithmic_bytes_per_input_px": ab,
                     "keypoints_per_step_rank0": r["keypoints_per_step"], "overflowed_frames": r["overflowed_frames"],
                     "h2d_bytes_per_step": r["h2d_bytes_per_step"], "d2h_bytes_per_step": r["d2h_bytes_per_step"], "note": note}

    entry("cfg1_512x512_4oct", 512, 512, 4, 64, 64, 4, "replicas", "BASELINE configs[0] shape; rank 0 only")
    per = max(1, 1024 // world)
    entry("cfg3_batch_1024x720p", 1280, 720, 4, per, 128, 2, "strong" if world > 1 else "strong (1 GPU: the whole batch)",
          f"BASELINE configs[2]: the fixed batch of 1024 frames, {per} per GPU in contiguous blocks, no collective; "
          "the rank's frames cycle through at most 128 distinct synthetic frames (same cost per frame)")
    entry("cfg4_3840x2160_6oct", 3840, 2160, 6, 8, 8, 3, "replicas", "BASELINE configs[3]; rank 0 only")
    return out


def mosaic_entry(eng, L, torch, fixtures, rank, world, local, dist):
    """BASELINE configs[4]: a square mosaic cut into `world` row strips, one per GPU, seed-halo exchange per octave over
    NCCL (the only path with a collective).  16384^2 at any N (checked against the whole-image result on rank 0);
    32768^2 at N = 8."""
    from sift_b200 import mosaic
    out = {}
    sizes = [16384] + ([32768] if world >= 8 else [])
    for size in sizes:
        prm = L.default_params(numberOfOctaves=N_OCT, scalesPerOctave=SPO, minBlurLevel=MIN_BLUR, assumedBlur=ASSUMED)
        layouts = mosaic.plan_strips(prm, size, size, world, 64)
        lay = layouts[rank]
        nblobs = max(64, size * size // 16384)
        t0 = time.perf_counter()
        rows = fixtures.synthetic_u8_rows(size, size, lay.top[0] // 2, (lay.bottom[0] + 1) // 2, 4321, blobs=nblobs)
        t_gen = time.perf_counter() - t0
        rows = torch.from_numpy(rows).pin_memory().numpy()
        phases = {}

        def tick(name, t0):
            torch.cuda.synchronize(); dist.barrier()
            phases[name] = phases.get(name, 0.0) + (time.perf_counter() - t0) * 1e3

        times = []
        for rep in range(3):
            dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            kps, stats = mosaic.detect_mosaic_distributed(eng, rows, layouts, prm, rank)
            torch.cuda.synchronize(); dist.barrier()
            times.append(time.perf_counter() - t0)
        # the same steps timed one by one (a barrier after each: the sum exceeds the pipelined time above)
        dist.barrier(); t0 = time.perf_counter()
        eng.strip_begin(prm, lay, rows); tick("upload_ms", t0)
        halo_bytes = 0
        for o in range(N_OCT):
            t0 = time.perf_counter()
            if o > 0:
                mosaic.exchange_seed_halos(mosaic.seed_tensor(eng, lay, o), layouts, o, rank)
                tick("halo_exchange_ms", t0); t0 = time.perf_counter()
                halo_bytes += sum(n * lay.width[o] * 8 for s_, d_, r_, n in mosaic.halo_transfers(layouts, o) if d_ == rank)
            eng.strip_octave(o); tick("blur_octaves_ms", t0)
        t0 = time.perf_counter()
        k2, st2 = eng.strip_finish(); tick("scan_refine_order_download_ms", t0)
        t0 = time.perf_counter()
        mosaic.resolve_escaped_distributed(eng, layouts, rank, k2, st2); tick("walk_handover_ms", t0)
        merged = mosaic.gather_keypoints(kps, rank, world)
        left = torch.tensor([stats["leftStrip"]], device="cuda")
        hb = torch.tensor([halo_bytes], device="cuda", dtype=torch.int64)
        dist.all_reduce(left); dist.all_reduce(hb)
        e = None
        if rank == 0:
            best = min(times)
            ab = algorithmic_bytes_per_input_px(N_OCT)["total"]
            e = {"mosaic": f"{size}x{size}", "n_gpus": world, "strips": world, "octaves": N_OCT, "seconds": best,
                 "value": size * size / 1e6 / best, "unit": "Mpixel/s", "reps_s": [round(t, 4) for t in times],
                 "roofline_frac_per_gpu": ab * size * size / best / world / 1e9 / measured_peak_gbs()[0],
                 "keypoints": int(len(merged)), "walks_handed_over": int(left.item()),
                 "halo_rows_per_octave": [int(lay.halo[o]) for o in range(N_OCT)], "halo_bytes_received_all_ranks": int(hb.item()),
                 "phases_ms": {k_: round(v, 3) for k_, v in phases.items()}, "generate_s_rank0": round(t_gen, 2),
                 "timed": "strip upload from pinned host rows + octaves with NCCL halo exchange + scan + refine + ordering + "
                          "download + walk hand-over, wall clock between two barriers, best of 3",
                 "collective": "torch.distributed batch_isend_irecv (NCCL P2P over NVLink) of fp64 seed rows, per octave"}
            if size == 16384:
                del rows
                img = fixtures.synthetic_u8(size, size, 4321, blobs=nblobs)
                whole, _ = eng.detect(img, prm)
                e["identical_to_whole_image"] = bool(whole.tobytes() == merged.tobytes())
                e["whole_image_keypoints"] = int(len(whole))
                del img, whole
            out[f"{size}x{size}"] = e
        dist.barrier()
    return out


# ------------------------------------------------------------------------------ CPU legs
def cpu_tiles(n: int):
    """n CPU_TILE^2 crops (float64 in [0,1]) of the benchmark's first frame."""
    from sift_b200 import fixtures
    frame = fixtures.to_float(fixtures.synthetic_u8(W, H, 1234))
    tiles = []
    for i in range(n):
        y = (i * 97) % (H - CPU_TILE)
        x = (i * 211) % (W - CPU_TILE)
        tiles.append(np.ascontiguousarray(frame[y:y + CPU_TILE, x:x + CPU_TILE]))
    return tiles


def cpu_detect_tile(tile):
    import oracle
    r = oracle.detect(tile, numberOfOctaves=N_OCT, scalesPerOctave=SPO, minBlurLevel=MIN_BLUR, assumedBlur=ASSUMED,
                      separable=False, keep_levels=False)
    return len(r.keypoints)


def cpu_baseline_leg() -> dict:
    """One host core, the float64 port of the reference's dense 2D path, bounded sample."""
    import oracle
    oracle.build()
    tiles = cpu_tiles(8)
    cpu_detect_tile(tiles[0])
    t0 = time.perf_counter()
    for t in tiles:
        cpu_detect_tile(t)
    dt = time.perf_counter() - t0
    return {"value": len(tiles) * CPU_TILE * CPU_TILE / 1e6 / dt, "unit": "Mpixel/s", "cores": 1, "kind": "port",
            "sample": f"{len(tiles)} crops {CPU_TILE}x{CPU_TILE} of frame 0, same params, dense 2D kernel, "
                      f"{dt:.1f} s of CPU work (Node.js absent: float64 C restatement of the reference JS)"}


def node_reference(tiles, cores):
    """The UNMODIFIED reference under Node.js (baseline/run_reference.mjs over the copy in baseline/_ref), one image
    per worker_thread.  Returns (seconds, info) or None when there is no `node` on this box / the copy is missing /
    the harness fails (the caller then falls back to the float64 port and says so)."""
    import shutil
    import tempfile
    node = shutil.which("node") or shutil.which("nodejs")
    ref = os.path.join(ROOT, "baseline", "_ref", "background.js")
    if not node or not os.path.exists(ref):
        return None
    try:
        u8 = np.stack([np.rint(t * 255.0).astype(np.uint8) for t in tiles])
        with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as f:
            f.write(u8.tobytes())
            path = f.name
        r = subprocess.run([node, "--max-old-space-size=8192", os.path.join(ROOT, "baseline", "run_reference.mjs"), path,
                            str(CPU_TILE), str(CPU_TILE), str(len(tiles)), str(N_OCT), str(SPO), str(MIN_BLUR), str(ASSUMED),
                            str(cores)], capture_output=True, text=True, timeout=1500)
        os.unlink(path)
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        if r.returncode != 0 or not line:
            return None
        d = json.loads(line[-1])
        return float(d["seconds"]), d
    except Exception:
        return None


def separable_port_whole_frames(n_frames: int = 2):
    """The separable float64 port (oracle_blur_image_separable: the same sums, row bands on all host cores) on WHOLE
    1920x1080 frames -- reported beside the dense-2D figure, which is what the reference's algorithm costs."""
    import oracle
    from sift_b200 import fixtures
    t0 = time.perf_counter()
    for i in range(n_frames):
        oracle.detect(fixtures.to_float(fixtures.synthetic_u8(W, H, 1234 + i)), numberOfOctaves=N_OCT, scalesPerOctave=SPO,
                      minBlurLevel=MIN_BLUR, assumedBlur=ASSUMED, separable=True, keep_levels=False)
    return n_frames * W * H / 1e6 / (time.perf_counter() - t0)


def run_reference(args):
    """--impl reference: the reference's own CPU implementation on all host cores, one image per thread (BASELINE.md
    section 3), same config / metric / unit: the unmodified JavaScript under Node.js when the box has a `node`
    (kind "reference"), else the float64 C port of its dense 2D algorithm (kind "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import concurrent.futures as cf
    import oracle
    oracle.build()
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    tiles = cpu_tiles(cores)
    js = node_reference(tiles, cores) if not os.environ.get("SIFT_BENCH_NO_NODE") else None
    if js is not None:
        dt, info = js
        value = len(tiles) * CPU_TILE * CPU_TILE / 1e6 / dt
        sample = (f"{len(tiles)} crops {CPU_TILE}x{CPU_TILE} of the 1920x1080 frame, one per worker_thread ({info.get('threads')} threads), "
                  f"unmodified reference JavaScript under Node {info.get('node')} (baseline/run_reference.mjs), one pass")
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Mpixel/s", "n_gpus": args.gpus,
                "steps": 1, "warmup": 0, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args.gpus),
                "cpu_baseline": {"value": value, "unit": "Mpixel/s", "cores": int(info.get("threads") or cores), "kind": "reference",
                                 "sample": sample},
                "e2e": {"value": value, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0, "keypoints_per_crop": info.get("keypoints")}
        print(json.dumps(line), flush=True)
        return
    ex = cf.ThreadPoolExecutor(max_workers=cores)   # ctypes releases the GIL inside the C call

    def step():
        list(ex.map(cpu_detect_tile, tiles))

    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = args.steps * cores * CPU_TILE * CPU_TILE / 1e6 / dt
    sample = (f"{cores} crops {CPU_TILE}x{CPU_TILE} per step (one per thread) of the 1920x1080 frame, "
              f"same params, dense 2D kernel; no `node` binary on this box, so this is the float64 C port of the reference")
    try:
        sep = separable_port_whole_frames(1 if CPU_TILE < 128 else 2)
    except Exception:
        sep = None
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Mpixel/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": "Mpixel/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "separable_port_whole_1080p_frames": {"value": sep, "unit": "Mpixel/s", "cores": cores,
                                                  "note": "same sums reorganised as two 1D passes (not the reference's "
                                                          "algorithm: reported beside it, not as it)"}}
    print(json.dumps(line), flush=True)


def workload_config(n_gpus: int) -> dict:
    return {"workload": f"{W}x{H} synthetic frames (blobs+checkerboard+noise), {N_OCT} octaves x s={SPO} "
                        f"(6 blur levels), sigma0={MIN_BLUR}, assumed blur {ASSUMED}, contrast 0.015 "
                        f"(reference constant), edge r=10, 2x-upsampled base octave",
            "frames_per_gpu_per_step": FRAMES, "sharding": f"images split by rank x{n_gpus}, no collective",
            "frames_in_flight_per_gpu": LANES or "auto by frame size: 8 below 6 Mpixel (1920x1080, 1280x720, 512x512), 6 at 3840x2160",
            "l2": "no flush: one frame's pyramid (735 MB algorithmic) exceeds the 126 MB L2 and frames rotate"}


# ------------------------------------------------------------------------------ own arm
def run_own(args):
    import torch
    import sift_b200
    from sift_b200 import _lib as L

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist = None

    eng = sift_b200.Engine(local)
    prm = L.default_params(numberOfOctaves=N_OCT, scalesPerOctave=SPO, minBlurLevel=MIN_BLUR, assumedBlur=ASSUMED)
    frames_np = make_frames(FRAMES, 1234 + rank * FRAMES)
    h_frames = torch.from_numpy(frames_np).pin_memory()
    d_frames = h_frames.cuda()
    cap = 1 << 15
    d_out = torch.zeros(FRAMES, cap * L.KEYPOINT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    d_cnt = torch.zeros(FRAMES, dtype=torch.int32, device="cuda")
    h_out = torch.zeros(FRAMES * cap * L.KEYPOINT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    h_offs = torch.zeros(FRAMES + 1, dtype=torch.int32)
    stream = torch.cuda.ExternalStream(eng.stream)
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        for f in range(FRAMES):
            eng.detect_device(d_frames[f].data_ptr(), L.SIFT_U8, W, H, 0, prm, d_out[f].data_ptr(), cap,
                              d_cnt[f].data_ptr())

    def step_e2e():
        # the reference-facing C-ABI call: HOST frames in, ordered HOST keypoint records out (copies inside)
        eng.detect_batch_raw(h_frames.data_ptr(), L.SIFT_U8, W, H, 0, W * H, FRAMES, prm, h_out.data_ptr(),
                             FRAMES * cap, h_offs.data_ptr())
        return int(h_offs[FRAMES])

    # ---- device-resident timing
    sampler = ClockSampler(local)
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    if rank == 0:
        sampler.start()
        # nvidia-smi needs ~0.3 s to print its first line: keep the GPU on the same load until it does, so
        # that the samples cover the timed region even when K steps take only a few milliseconds
        t_spin = time.perf_counter()
        while not sampler.lines and time.perf_counter() - t_spin < 3.0:
            step_device()
            eng.synchronize()
    barrier()
    l0 = eng.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    eng.flush()                      # the public stream waits for every frame in flight (host not blocked)
    e1.record(stream)
    barrier()
    launches = eng.kernel_launches - l0
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if rank == 0:                                   # a few more loaded samples, then stop
        t_spin = time.perf_counter()
        while len(sampler.lines) < 4 and time.perf_counter() - t_spin < 1.0:
            step_device()
            eng.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    px_step_all = world * FRAMES * W * H
    value = args.steps * px_step_all / 1e6 / (ms_total / 1e3)
    n_kp = int(d_cnt.sum().item())

    # ---- end to end through the C ABI with host buffers
    for _ in range(max(args.warmup, 3)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], device="cuda")
    if dist is not None:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = args.steps * px_step_all / 1e6 / float(t_e2e.item())
    # bytes counted by the engine where it issues the copies (sift_transfer_bytes), one more step
    tb0 = eng.transfer_bytes
    step_e2e()
    tb1 = eng.transfer_bytes
    h2d_step, d2h_step = tb1[0] - tb0[0], tb1[1] - tb0[1]

    # ---- per-kernel-class timing (separate instrumented pass, CUDA events on the engine's stream)
    eng.set_lanes(1)                 # one frame at a time: per-kernel times without cross-frame overlap
    eng.set_profiling(True)
    prof_steps = 2
    for _ in range(prof_steps):
        step_device()
    prof = eng.get_profile()
    eng.set_profiling(False)
    eng.set_lanes(LANES)

    # ---- the other named shapes, and (N > 1) the mosaic: extra keys of the same line
    from sift_b200 import fixtures
    del d_out, d_frames
    torch.cuda.empty_cache()
    configs, mosaic_res = {}, {}
    if not args.headline_only:
        try:
            configs = extra_configs(eng, L, torch, fixtures, rank, world, dist, measured_peak_gbs()[0])
        except Exception as ex:                       # never lose the headline over an extra
            configs = {"error": repr(ex)}
        if world > 1:
            try:
                mosaic_res = mosaic_entry(eng, L, torch, fixtures, rank, world, local, dist)
            except Exception as ex:
                mosaic_res = {"error": repr(ex)}

    if rank == 0:
        ab = algorithmic_bytes_per_input_px()
        peak, peak_src = measured_peak_gbs()
        n_px = W * H
        kinds = {}
        for kind, (kms, cnt) in prof.items():
            kinds[kind] = {"ms_per_frame": kms / (prof_steps * FRAMES), "launch_groups": cnt}
        tot_ms = sum(k["ms_per_frame"] for k in kinds.values()) or 1.0
        dom = "blur_octave0"
        traffic, traffic_src = ncu_traffic(DOMINANT_KERNEL)
        dom_ms = kinds[dom]["ms_per_frame"]
        dom_bytes = ab["blur"][0] * n_px
        achieved = dom_bytes / (dom_ms / 1e3) / 1e9 if dom_ms > 0 else 0.0
        whole_gbs = ab["total"] * (value * 1e6 / world) / 1e9
        for k in kinds.values():
            k["share"] = k["ms_per_frame"] / tot_ms
        line = {
            "metric": METRIC, "value": value, "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(world),
            "e2e": {"value": e2e_value, "unit": "Mpixel/s", "h2d_bytes_per_step": h2d_step,
                    "d2h_bytes_per_step": d2h_step, "bytes_counted_by": "sift_transfer_bytes() around one step",
                    "note": "sift_detect_batch(): pinned host u8 frames in, ordered keypoint records out; "
                            "upload / compute / download+ordering of consecutive frames overlap"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "octave-0 upsample+blur+DoG (" + dom + ")",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": dom_bytes, "ms_per_launch": dom_ms,
                         "whole_path": {"algorithmic_bytes_per_input_px": ab["total"],
                                        "achieved": whole_gbs, "frac": whole_gbs / peak},
                         "fp64_pipe": {"note": "the blur is bound by the fp64 pipe, not by HBM (DESIGN.md section 3): useful "
                                               "multiply-adds of the separable float64 blur against the measured fp64 peak",
                                       "useful_fma_per_input_px": FMA_PER_INPUT_PX[N_OCT],
                                       "achieved_tfma_per_s": FMA_PER_INPUT_PX[N_OCT] * (value * 1e6 / world) / 1e12,
                                       "peak_tfma_per_s": FP64_PEAK_TFMA,
                                       "frac": FMA_PER_INPUT_PX[N_OCT] * (value * 1e6 / world) / 1e12 / FP64_PEAK_TFMA},
                         "kernels": kinds},
            "keypoints_per_step": n_kp,
            "clocks": clocks,
        }
        line["configs"] = configs
        if mosaic_res:
            line["mosaic"] = mosaic_res
        if world == 1:
            try:
                line["cpu_baseline"] = cpu_baseline_leg()
            except Exception as ex:   # the baseline is a reported figure; never lose the GPU line over it
                line["cpu_baseline"] = {"value": None, "unit": "Mpixel/s", "cores": 1, "kind": "port",
                                        "sample": f"failed: {ex}"}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--headline-only", action="store_true", help="skip the extra configs / mosaic entries")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
