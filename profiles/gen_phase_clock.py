"""Phase breakdown of one generated sample inside the persistent sampler kernel (csrc/gen.cu).

The kernel's optional `dbg_clock` hook makes CTA 0 / thread 0 write clock64() stamps for the 9th step of a launch:
step prologue, then per layer: gate chunks, CTA barrier, z finalize, z exchange, mix block, x exchange.
    python profiles/gen_phase_clock.py > profiles/r1g_gen_phase_clock.txt      (on the GPU box)
"""
import os, sys, torch, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'ae-wavenet_b200'))
import bench_generate as bg
from aewn import generate
wn = bg.build()
rf1 = wn.base_global_rf
n_ts = rf1 + 200
g = torch.Generator().manual_seed(1)
cond = torch.randn(wn.n_cond, n_ts, generator=g).cuda()
codes = torch.randint(0, 256, (n_ts + 8,), generator=g).cuda()
plan = generate.get_plan(wn, 1)
clk = torch.zeros(512, dtype=torch.int64, device='cuda')
plan.desc.dbg_clock = clk.data_ptr()
plan.generate(codes, cond, rf1)
c = clk.cpu().tolist()
fine = c[256:256+24]
mixs = c[300:308]
xs = c[320:325]
c = c[:256]
n = max(i for i, v in enumerate(c) if v) + 1
d = [c[i+1]-c[i] for i in range(n-1)]
print('prologue', d[0])
names = ['gate', 'bar', 'final', 'xchgA', 'mix', 'xchgB']
for l in range(20):
    print(l, {k: d[1 + 6*l + i] for i, k in enumerate(names)}, )
print('total cycles', c[n-1]-c[0])

print('gate chunk stamps (layer 3): wait / math+release / gap-to-next')
for k in range(8):
    a, b, e = fine[3*k:3*k+3]
    nxt = fine[3*k+3] if k < 7 else e
    print(k, b - a, e - b, nxt - e)

print('mix chunk stamps (layer 3): wait / math / release')
for k in range(2):
    a, b, m_, e = mixs[4*k:4*k+4]
    print(k, b - a, m_ - b, e - m_)

print('xchgA (layer 3): bar / issue / prefetch_hist / wait', [xs[i+1]-xs[i] for i in range(4)])
