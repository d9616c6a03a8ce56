"""Loader arithmetic: GPU kernels (aewn/loader.py) vs the host path they replace, at the cfg2 batch shape (8 windows of
16384 + receptive field + MFCC wings = 18470 samples).  The host side is the oracle's numpy / scipy restatement of
librosa's MFCC (librosa itself is not in this image), one item at a time like data.Collate (data.py:223-231).

    python profiles/loader_latency.py > profiles/rN_loader_latency.txt      (on the GPU box)
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ae-wavenet_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from aewn import loader  # noqa: E402
import loader_oracle as lo  # noqa: E402

B, n = 8, 18470
rs = np.random.RandomState(0)
wav = rs.randint(0, 256, (B, n)).astype(np.uint8)
pw, jit = loader.ProcessWav(), loader.Jitter(0.12)
dev = torch.from_numpy(wav).cuda()


def gpu_us(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


F = pw.n_frames(n)[1]
print(f"batch {B} x {n} samples -> mel (B, 39, {F})")
print(f"GPU  ProcessWav.batch          {gpu_us(lambda: pw.batch(dev)):9.1f} us per batch (3 launches + table lookups)")
print(f"GPU  Jitter.batch              {gpu_us(lambda: jit.batch(B, F)):9.1f} us per batch")
x = torch.rand(B, n, device='cuda') * 2 - 1
print(f"GPU  mu_encode_torch           {gpu_us(lambda: loader.mu_encode_torch(x, 256)):9.1f} us per batch")
t0 = time.perf_counter()
for b in range(B):
    lo.process_wav(wav[b])
t1 = time.perf_counter()
print(f"host numpy/scipy MFCC + deltas {1e6 * (t1 - t0):9.1f} us per batch (one item at a time, 1 thread)")
t0 = time.perf_counter()
for b in range(B):
    lo.jitter_from_uniforms(np.random.random_sample(F - 2), F, 0.12)
t1 = time.perf_counter()
print(f"host jitter (python loop)      {1e6 * (t1 - t0):9.1f} us per batch")
pinned = torch.from_numpy(wav).pin_memory()
print(f"H2D copy of the uint8 windows  {gpu_us(lambda: dev.copy_(pinned, non_blocking=True)):9.1f} us per batch ({wav.nbytes} bytes, pinned)")
