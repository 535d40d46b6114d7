"""Throughput of the GPU input pipeline: 1902 ACDC-sized synthetic slices resident on the GPU, batches of 64 at 256x256.
Prints images/s of (host parameter draws + table upload + wsl_augment_batch) and of the kernel alone."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from wsl4mis_b200.dataloaders import GpuLoader, SliceStore  # noqa: E402

rs = np.random.RandomState(0)
ims, lbs = [], []
for i in range(1902):
    h, w = int(rs.choice([208, 216, 224, 232, 256])), int(rs.choice([154, 174, 208, 216, 256]))
    ims.append(rs.rand(h, w).astype(np.float32))
    lb = np.full((h, w), 4, np.uint8)
    m = rs.rand(h, w) < 0.03
    lb[m] = rs.randint(0, 4, size=int(m.sum()))
    lbs.append(lb)
st = SliceStore.from_arrays(ims, lbs)
ld = GpuLoader(st, 64, (256, 256), shuffle=True, drop_last=True)
for _ in ld:      # warm-up epoch
    pass
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 0
for _ in range(3):
    for b in ld:
        n += b["image"].shape[0]
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"loader: {n / dt:.0f} images/s end to end ({n} images, {dt * 1e3:.1f} ms)")
idx = list(range(64))
params = [ld.transform.draw(True) for _ in idx]
st.augment(idx, params, (256, 256))
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(50):
    st.augment(idx, params, (256, 256))
e.record()
torch.cuda.synchronize()
print(f"augment(64 x 256 x 256) incl. table upload: {s.elapsed_time(e) / 50 * 1e3:.1f} us per batch")
