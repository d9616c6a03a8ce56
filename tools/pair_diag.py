"""Diagnostic for the CTA-pair (cta_group::2) engine mode: runs one small tgemm / wgrad with exactly representable
integer data in both modes and prints the error structure (per time-tile parity = CTA rank, per column half = which
CTA staged the W rows).  Not part of the test-suite; used when bringing the mode up on hardware."""
import ctypes as C
import sys
import os

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "ae-wavenet_b200"))
import torch
from aewn import ops, _lib as L


def run_tgemm(cluster, N=256, R=64, T=1024, B=2):
    g = torch.Generator().manual_seed(1)
    x = torch.randint(-3, 4, (B, R, T), generator=g).float()
    w = torch.randint(-3, 4, (N, R), generator=g).float()
    xb = ops.to_buf(x.cuda())
    out = ops.new_buf(B, N, xb.shape[2], "cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    tiles = [ops.ntile(c0, n, out[:, c0:], t_lo=0, t_hi=T) for (c0, n) in ops.chunks(N)]
    rc = 0
    for kind, d, tag in ops.build_tgemm([ops.act_of(xb, T)], [(0, 0, R, 0)], w.cuda(), tiles, B, 0, T, err):
        d.cluster = cluster
        rc = L.lib().aewn_tgemm(C.byref(d), ops._stream())
    torch.cuda.synchronize()
    ref = torch.einsum("nr,brt->bnt", w, x)
    got = out[:, :, :T].cpu()
    diff = (got - ref).abs()
    print(f"tgemm cluster={cluster} N={N}: rc={rc} err_word={int(err.item())} max_err={float(diff.max())}")
    if float(diff.max()) > 0:
        for par in (0, 1):
            for half in (0, 1):
                tsel = torch.arange(T)[(torch.arange(T) // 128) % 2 == par]
                blk = diff[:, half * (N // 2):(half + 1) * (N // 2)][:, :, tsel]
                print(f"   time-tile parity {par} col-half {half}: max {float(blk.max()):.1f} "
                      f"frac_wrong {float((blk > 0).float().mean()):.3f} got_zero {float((got[:, half * (N // 2):(half + 1) * (N // 2)][:, :, tsel] == 0).float().mean()):.3f}")


def run_wgrad(engine, N=256, M=256, T=2048, B=2):
    ops.ENGINE_MODE = engine
    g = torch.Generator().manual_seed(2)
    G = torch.randint(-2, 3, (B, M, T), generator=g).float()
    X = torch.randint(-2, 3, (B, N, T), generator=g).float()
    Gb, Xb = ops.to_buf(G.cuda()), ops.to_buf(X.cuda())
    out = torch.zeros(M, N).cuda()
    groups = [[dict(g_act=0, x_act=1, g_row=128 * i, x_row=c0, m_valid=min(128, M - 128 * i), n_valid=n, shift=0,
                    t_lo=0, t_hi=T, out=out, out_off=128 * i * N + c0, out_rs=N, out_cs=1)
               for i in range((M + 127) // 128)] for (c0, n) in ops.chunks(N, 256)]
    items = ops.pair_items(groups)
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    rc = 0
    for kind, d, tag in ops.build_wgrad([ops.act_of(Gb, T), ops.act_of(Xb, T)], items, B, err, pair=True):
        px = d.pair_x
        rc = L.lib().aewn_wgrad(C.byref(d), ops._stream())
    torch.cuda.synchronize()
    ref = torch.einsum("bmt,bnt->mn", G, X)
    diff = (out.cpu() - ref).abs()
    print(f"wgrad engine={engine} pair_x={px} N={N}: rc={rc} err_word={int(err.item())} max_err={float(diff.max())}")
    if float(diff.max()) > 0:
        for mh in (0, 1):
            for nh in (0, 1):
                blk = diff[mh * 128:(mh + 1) * 128, nh * (N // 2):(nh + 1) * (N // 2)]
                print(f"   m-tile {mh} col-half {nh}: max {float(blk.max()):.1f} frac_wrong {float((blk > 0).float().mean()):.3f}")


def run_wgradw(widths=(256, 256), M=256, T=2048, B=2):
    g = torch.Generator().manual_seed(3)
    G = torch.randint(-2, 3, (B, M, T), generator=g).float()
    keep = [ops.to_buf(G.cuda())]                      # the descriptors hold raw pointers: keep the buffers alive
    acts = [ops.act_of(keep[0], T)]
    cks, refs = [], []
    for n in widths:
        X = torch.randint(-2, 3, (B, n, T), generator=g).float()
        out = torch.zeros(M, n).cuda()
        keep.append(ops.to_buf(X.cuda()))
        acts.append(ops.act_of(keep[-1], T))
        cks.append(dict(x_act=len(acts) - 1, x_row=0, n_valid=n, shift=0, out=out, out_off=0, out_rs=n, out_cs=1))
        refs.append((out, torch.einsum("bmt,bnt->mn", G, X)))
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    rc = 0
    for kind, d, tag in ops.build_wgradw(acts, ops.pack_wide_units(0, 0, M, 0, T, cks), B, err):
        rc = L.lib().aewn_wgradw(C.byref(d), ops._stream())
    torch.cuda.synchronize()
    worst = max(float((o.cpu() - r).abs().max()) for o, r in refs)
    print(f"wgradw widths={widths}: rc={rc} err_word={int(err.item())} max_err={worst}")
    if worst > 0:
        for ci, (o, r) in enumerate(refs):
            diff = (o.cpu() - r).abs()
            n = diff.shape[1]
            for mh in (0, 1):
                for nh in (0, 1):
                    blk = diff[mh * 128:(mh + 1) * 128, nh * (n // 2):(nh + 1) * (n // 2)]
                    if blk.numel():
                        print(f"   chunk {ci} m-half {mh} col-half {nh}: max {float(blk.max()):.1f} frac_wrong {float((blk > 0).float().mean()):.3f}")


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "tgemm"):
        for cl in (2, L.CLUSTER_PAIR_MMA):
            for N in (256, 48, 368):
                run_tgemm(cl, N=N)
    if which in ("all", "wgrad"):
        for eng in ("mcast", "pair"):
            for N in (256, 112):
                run_wgrad(eng, N=N)
    if which in ("all", "wgradw"):
        for widths in ((256, 256), (112, 112, 144), (256,), (48,)):
            print("launching wgradw", widths, flush=True)
            run_wgradw(widths)
