"""Per-kernel-class device times (CUDA events inside the engine, one frame at a time) + frames-in-flight throughput.
usage: python tools/kernel_times.py [W H OCTAVES [FRAMES]]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import sift_b200
from sift_b200 import _lib as L, fixtures

W, H, NO = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1920, 1080, 4)
FR = int(sys.argv[4]) if len(sys.argv) > 4 else 16
eng = sift_b200.Engine(0)
prm = L.default_params(numberOfOctaves=NO, minBlurLevel=1.6)
frames = torch.from_numpy(np.stack([fixtures.synthetic_u8(W, H, 1234 + i) for i in range(FR)])).cuda()
cap = max(1 << 15, W * H // 32)
out = torch.zeros(FR, cap * L.KEYPOINT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
cnt = torch.zeros(FR, dtype=torch.int32, device="cuda")

def step():
    for f in range(FR):
        eng.detect_device(frames[f].data_ptr(), L.SIFT_U8, W, H, 0, prm, out[f].data_ptr(), cap, cnt[f].data_ptr())

for lanes in (1, 3, 4, 5, 6, 8, 0):
    eng.set_lanes(lanes)
    for _ in range(3):
        step()
    eng.synchronize()
    t0 = time.perf_counter()
    n = 6
    for _ in range(n):
        step()
    eng.synchronize()
    dt = (time.perf_counter() - t0) / (n * FR)
    print(f"{W}x{H}/{NO}oct lanes={lanes}: {dt*1e3:.4f} ms/frame  {W*H/1e6/dt:.0f} Mpx/s  kp/frame {int(cnt.sum())//FR}")
eng.set_lanes(1)
eng.set_profiling(True)
for _ in range(3):
    step()
prof = eng.get_profile()
eng.set_profiling(False)
tot = 0
for k, (ms, c) in prof.items():
    print(f"  {k:20s} {ms/(3*FR):.4f} ms/frame")
    tot += ms / (3 * FR)
print(f"  {'sum':20s} {tot:.4f}")
