"""BASELINE configs[2]: a batch of 1280x720 frames through sift_detect_batch (host buffers in, ordered records out).
usage: python tools/batch_720p.py [n_frames=1024] [distinct=64]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sift_b200
from sift_b200 import _lib as L, fixtures

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
distinct = int(sys.argv[2]) if len(sys.argv) > 2 else 64
W, H = 1280, 720
base = np.stack([fixtures.synthetic_u8(W, H, 1234 + i) for i in range(distinct)])
frames = torch.from_numpy(np.concatenate([base] * (n // distinct))).pin_memory()
n = frames.shape[0]
eng = sift_b200.Engine(0)
prm = L.default_params(numberOfOctaves=4, minBlurLevel=1.6)
cap = n * 8192
out = torch.zeros(cap * 80, dtype=torch.uint8).pin_memory()
offs = torch.zeros(n + 1, dtype=torch.int32)
for rep in range(2):
    t0 = time.perf_counter()
    st = eng.detect_batch_raw(frames.data_ptr(), L.SIFT_U8, W, H, 0, W * H, n, prm, out.data_ptr(), cap, offs.data_ptr())
    dt = time.perf_counter() - t0
k = np.frombuffer(out.numpy(), dtype=L.KEYPOINT_DTYPE)[:int(offs[n])]
o = offs.numpy()
same = all(k[o[i]:o[i + 1]].tobytes() == k[o[i + distinct]:o[i + distinct + 1]].tobytes() for i in range(0, n - distinct, 7))
print(json.dumps({"config": f"batch of {n} {W}x{H} frames ({distinct} distinct), 4 octaves, sift_detect_batch (host in / ordered host out)",
                  "seconds": dt, "mpixel_per_s": n * W * H / 1e6 / dt, "frames_per_s": n / dt, "keypoints": int(offs[n]),
                  "repeated_frames_identical": bool(same), "device_ms": st.msDevice}))
