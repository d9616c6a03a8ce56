"""Latency of the VQ-EMA bottleneck step at the cfg3 shapes (batch 16, n_in 768 -> d 32, N 65, K 4096): the kernel path
(exact-fp32 1x1 projection + ONE fused distance / argmin / gather / histogram / EMA-statistics kernel + EMA update) against
the reference's op sequence run eagerly on the same GPU (oracle port of vqema_bn.py:125-195: two (B, K, d, N) temporaries
of 545 MB each, ~15 launches).  SURVEY.md 8d: the step is launch-latency bound, so microseconds are the metric.

    python profiles/vq_latency.py > profiles/rN_vq_latency.txt        (on the GPU box)
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ae-wavenet_b200"))
from aewn import vqema_bn  # noqa: E402
from oracle import torch_oracle as orc  # noqa: E402

torch.manual_seed(2507)
B, n_in, d, N, K = 16, 768, 32, 65, 4096
bn = vqema_bn.VQEMA(n_in, d, 0.25, 0.99, K, True).cuda()
z = torch.randn(B, n_in, N, device="cuda")


def timed(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps


def ours():
    with torch.no_grad():
        bn(z)


lin_w, emb = bn.linear.weight.detach(), bn.emb.detach()
numer, denom = bn.ema_numer.clone(), bn.ema_denom.clone()


def eager():
    with torch.no_grad():
        orc.vqema_forward(lin_w, emb, z, numer, denom, 0.99)


def ours_assign_only():
    with torch.no_grad():
        vqema_bn._VQAssignFn.apply(bn.ze.detach(), emb, 1, bn.ind_hist, bn.z_sum, bn.n_sum, True)


t_ours, t_eager = timed(ours), timed(eager, reps=10)
t_k = timed(ours_assign_only)
r = orc.vqema_forward(lin_w, emb, z, numer, denom, 0.99)
same = float((r["min_ind"] == bn.min_ind).float().mean())
print(f"VQEMA.forward (train mode, diagnostics included), kernel path : {t_ours:8.1f} us per call")
print(f"  of which the fused vq_fwd kernel call alone                  : {t_k:8.1f} us")
print(f"reference op sequence, eager PyTorch on this GPU               : {t_eager:8.1f} us per call   ({t_eager / t_ours:.1f}x)")
print(f"codes equal to the eager fp32 result on this input             : {100 * same:.2f} %")
