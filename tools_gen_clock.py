import sys, torch, ctypes as C
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/ae-wavenet_b200')
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
n = max(i for i, v in enumerate(c) if v) + 1
d = [c[i+1]-c[i] for i in range(n-1)]
print('prologue', d[0])
names = ['gate', 'bar', 'final', 'xchgA', 'mix', 'xchgB']
for l in range(20):
    print(l, {k: d[1 + 6*l + i] for i, k in enumerate(names)}, )
print('total cycles', c[n-1]-c[0])
