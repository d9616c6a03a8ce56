"""Find what invalidates the CUDA-graph capture of the cfg2 step (debug helper)."""
import os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ae-wavenet_b200"))
import torch
import aewn
from aewn import geometry as vc, ops
from aewn.dist import FlatGradSync
from aewn.train import GraphedStep
import bench

fa = os.environ.get("FUSED_ACC", "1") == "1"
W = 2048
wn, geo = bench.build_decoder(W, aewn.WaveNet, vc)
wn = wn.cuda().train()
loss_fn = aewn.RecLoss()
sync = FlatGradSync(wn.parameters(), fused_accumulate=fa)
opt = torch.optim.Adam(wn.parameters(), lr=2e-5, fused=True, capturable=True)
wav, lc, spk, jit = [t.cuda() for t in bench.synth_batch(2, geo["wav_len"], geo["lc_len"], 64, 40, 1)]
t0w, t1w = geo["trim_dec_out"]

def step(wav, lc, spk, jit):
    sync.zero_grad()
    quant = wn(wav, lc, spk, jit)
    loss = loss_fn(quant[..., :-1], wav[:, t0w:t1w][..., 1:])
    loss.backward()
    sync.sync()
    opt.step()
    return loss

for _ in range(3):
    step(wav, lc, spk, jit)
torch.cuda.synchronize()
try:
    g = GraphedStep(step, [wav, lc, spk, jit], warmup=1)
    print("capture OK, loss", float(g(wav, lc, spk, jit)))
except Exception:
    traceback.print_exc()
