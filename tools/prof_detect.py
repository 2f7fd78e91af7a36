"""Tiny driver for ncu: a few device-resident detections of one 1920x1080 frame (BASELINE configs[1])."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sift_b200
from sift_b200 import _lib as L, fixtures

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
n = int(sys.argv[3]) if len(sys.argv) > 3 else 3
eng = sift_b200.Engine(0)
prm = L.default_params(numberOfOctaves=4, minBlurLevel=1.6)
d = torch.from_numpy(fixtures.synthetic_u8(W, H, 1234)).cuda()
cap = 1 << 15
out = torch.zeros(cap * 80, dtype=torch.uint8, device="cuda")
cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
for _ in range(n):
    eng.detect_device(d.data_ptr(), L.SIFT_U8, W, H, 0, prm, out.data_ptr(), cap, cnt.data_ptr())
eng.synchronize()
print("keypoints", int(cnt.item()))
