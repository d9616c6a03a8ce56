"""Phase clock of the fused layer kernel: runs one cfg2-shaped layer (batch 8, T0 18430, arch.basic widths) with the
debug stamp buffer attached and prints, for cluster 0's first four tiles, what the MMA issuer and epilogue warp 0 did
per job (cycles): wait for the TMEM region / operands, issue (MMA) or drain (epilogue).

    python profiles/gf_phase_clock.py [layer_dilation]  > profiles/rN_gf_phase_clock.txt      (on the GPU box)
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ae-wavenet_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from aewn import ops, _lib as L  # noqa: E402
import ctypes as C  # noqa: E402
from test_gpu_fused_layer import make_params  # noqa: E402

dil = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda")
gen = torch.Generator().manual_seed(0)
R, D, S, Cc, B, T0 = 368, 256, 256, 138, 8, 18430
dils = [dil, 8]
params = make_params(R, D, S, Cc, dils, True, gen, dev)
plan = ops.StackPlan(B, R, D, S, Cc, ops.StackGeom(dils, T0), params, dev, relu_last=False)
plan.sig[0][:, :, :T0] = torch.randn(B, R, T0, generator=gen).to(dev)
plan.cond[:, :Cc, :T0] = torch.randn(B, Cc, T0, generator=gen).to(dev)
for _ in range(3):
    plan.forward(save=True)
clk = torch.zeros(2 * 4 * 8 * 6, dtype=torch.int64, device=dev)
kind, d, tag = plan.fwd_train[0]
d.dbg_clock = clk.data_ptr()
d.max_ctas = int(os.environ.get("GF_MAX_CTAS", "0"))      # experiment: fewer CTAs -> is the limit per SM or chip-wide?
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
L.check(L.lib().aewn_grcc_fwd(C.byref(d), ops._stream()), "aewn_grcc_fwd")
e1.record()
torch.cuda.synchronize()
print(f"layer dil={dil}: {e0.elapsed_time(e1) * 1e3:.1f} us")
c = clk.cpu().view(2, 4, 8, 6)
t0 = int(c[:, :, :, 0][c[:, :, :, 0] > 0].min())
names = ["G1.0", "G1.1", "SKP", "RES1", "RES0", "-", "-", "-"]      # job order of arch.basic (R = 368)
for role, rn in enumerate(("MMA issuer", "epilogue warp 0")):
    print(rn, "(cycles from first stamp: job seen | ready | done ; wait, work)")
    for tile in range(4):
        for job in range(8):
            a, b, e, x1, x2, x3 = [int(v) for v in c[role, tile, job]]
            if a == 0:
                continue
            if job == 6:
                if role == 0:
                    print(f"  whole kernel, cluster 0: {b} cycles in {e} ns over {x1} tiles = {b / max(e, 1):.3f} GHz, "
                          f"{b / max(x1, 1):.0f} cycles per tile")
                continue
            if job == 7:
                print(f"  tile {tile} top   {a - t0:8d}")
                continue
            aux = (f"ring-wait {x1:6d}  z-wait {x3:6d}" if role == 0 else
                   (f"loads issued +{x1 - a:6d}  first chunk in registers +{x2 - a:6d}" if x1 > 0 else ""))
            print(f"  tile {tile} {names[job]:5s} {a - t0:8d} {b - t0:8d} {e - t0:8d}   wait {b - a:7d}  work {e - b:7d}   {aux}")
