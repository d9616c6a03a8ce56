"""Unit parity of the two tcgen05 engines through the C ABI (ops.tgemm / ops.wgrad) against plain PyTorch fp32
(TF32 tolerance), covering: shifted segments, ragged channel counts (K not a multiple of 32, N not a multiple of 16),
all three multicast cluster sizes and the CTA-pair (cta_group::2) mode, both epilogue store paths (TMA store / st.global), accumulate via TMA reduce-add, the
16-byte origin rule (rejected host-side instead of faulting on the device)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def make_desc_run(acts, segs, w, tiles, B, t0, t1, cluster, no_tma):
    from aewn import ops, _lib as L
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    for kind, d, tag in ops.build_tgemm(acts, segs, w, tiles, B, t0, t1, err):
        d.cluster = cluster
        d.no_tma_store = no_tma
        L.check(L.lib().aewn_tgemm(C.byref(d), ops._stream()), "aewn_tgemm")
    torch.cuda.synchronize()
    assert int(err.item()) == 0


@pytest.mark.parametrize("cluster", [1, 2, 4, 102])      # 102 = AEWN_CLUSTER_PAIR_MMA (cta_group::2 MMAs)
@pytest.mark.parametrize("no_tma", [0, 1])
def test_tgemm_shifted_segments_match_torch(cluster, no_tma):
    from aewn import ops
    g = torch.Generator().manual_seed(cluster * 10 + no_tma)
    B, R, Cc, T, d, N = 3, 72, 11, 900, 8, 300
    x = torch.randn(B, R, T, generator=g)
    c = torch.randn(B, Cc, T, generator=g)
    w0, w1, wc = torch.randn(N, R, generator=g), torch.randn(N, R, generator=g), torch.randn(N, Cc, generator=g)
    kr, kc = ops.ceil_to(R, 32), ops.ceil_to(Cc, 32)
    w = torch.zeros(N, 2 * kr + kc)
    w[:, :R], w[:, kr:kr + R], w[:, 2 * kr:2 * kr + Cc] = w0, w1, wc
    xb, cb = ops.to_buf(x.cuda()), ops.to_buf(c.cuda())
    prev = torch.randn(B, N, T, generator=g)
    out = ops.to_buf(prev.cuda())
    lo = 40        # multiple of 4: with TMA stores rows of the first tile below t_lo would be zeroed, so start aligned
    tiles = [ops.ntile(c0, n, out[:, c0:], flags=ops.L.F_ACCUM if c0 else 0, t_lo=lo & ~31, t_hi=T, t_zero_lo=lo)
             for (c0, n) in ops.chunks(N)]
    make_desc_run([ops.act_of(xb, T), ops.act_of(cb, T)], [(0, -d, R, 0), (0, 0, R, kr), (1, 0, Cc, 2 * kr)], w.cuda(),
                  tiles, B, lo & ~31, T, cluster, no_tma)
    xs = torch.nn.functional.pad(x, (d, 0))[:, :, :T]
    ref = torch.einsum("nr,brt->bnt", w0, xs) + torch.einsum("nr,brt->bnt", w1, x) + torch.einsum("nc,bct->bnt", wc, c)
    ref[:, :, :lo] = 0
    ref[:, 256:] += prev[:, 256:]                     # second n-tile accumulates (TMA reduce-add / RMW)
    got = out[:, :, :T].cpu()
    sel = slice(lo & ~31, T)
    err = float((got[:, :, sel] - ref[:, :, sel]).abs().max()) / float(ref.abs().max())
    assert err < 3e-3, err
    assert torch.equal(got[:, :, :lo & ~31], prev[:, :, :lo & ~31])      # untouched before the first tile


def test_unaligned_shift_is_rejected_on_the_host():
    from aewn import ops
    x = ops.to_buf(torch.randn(1, 32, 256).cuda())
    out = ops.new_buf(1, 16, 256, "cuda")
    w = torch.zeros(16, 32).cuda()
    with pytest.raises(RuntimeError, match="multiple of 4"):
        ops.tgemm([ops.act_of(x, 256)], [(0, -1, 32, 0)], w, [ops.ntile(0, 16, out, t_lo=0, t_hi=256)], 1, 0, 256)


@pytest.mark.parametrize("engine", ["mcast", "pair"])
@pytest.mark.parametrize("pair,N,chunk", [(False, 150, 128), (True, 150, 128), (True, 368, 384), (False, 300, 384),
                                          (True, 368, 256)])
def test_wgrad_matches_torch(pair, N, chunk, engine, monkeypatch):
    from aewn import ops, _lib as L
    monkeypatch.setattr(ops, "ENGINE_MODE", engine)   # "pair": X-sharing item pairs run cta_group::2 MMAs (pair_x = 2)
    g = torch.Generator().manual_seed(3)
    B, M, T, shift, t_lo = 2, 200, 1500, -8, 12
    G = torch.randn(B, M, T, generator=g)
    X = torch.randn(B, N, T, generator=g)
    Gb, Xb = ops.to_buf(G.cuda()), ops.to_buf(X.cuda())
    out = torch.zeros(M, N, 2).cuda()                 # conv-weight layout (out, in, tap): write tap 1
    groups = [[dict(g_act=0, x_act=1, g_row=128 * i, x_row=c0, m_valid=min(128, M - 128 * i), n_valid=n, shift=shift,
                    t_lo=t_lo, t_hi=T, out=out, out_off=128 * i * N * 2 + c0 * 2 + 1, out_rs=2 * N, out_cs=2)
               for i in range(2)] for (c0, n) in ops.chunks(N, chunk)]
    items = ops.pair_items(groups) if pair else [it for grp in groups for it in grp]
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    for kind, d, tag in ops.build_wgrad([ops.act_of(Gb, T), ops.act_of(Xb, T)], items, B, err, pair=pair):
        L.check(L.lib().aewn_wgrad(C.byref(d), ops._stream()), "aewn_wgrad")
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    Xs = torch.nn.functional.pad(X, (-shift, 0))[:, :, :T]          # X[u + shift]
    ref = torch.einsum("bmt,bnt->mn", G[:, :, t_lo:], Xs[:, :, t_lo:])
    got = out.cpu()
    assert float((got[:, :, 1] - ref).abs().max()) / float(ref.abs().max()) < 3e-3
    assert float(got[:, :, 0].abs().max()) == 0.0


@pytest.mark.parametrize("f,s,L", [(25, 5, 63), (16, 4, 295), (16, 2, 40)])
def test_polyphase_conv_transpose_matches_torch(f, s, L):
    """Upsampling.tconv (wavenet.py:154-155: ConvTranspose1d, padding = f - s) as polyphase GEMMs: forward, data and
    weight gradients vs plain PyTorch fp32."""
    from aewn import ops
    g = torch.Generator().manual_seed(f * 100 + s)
    x = torch.randn(3, 48, L, generator=g)
    w = torch.randn(48, 40, f, generator=g) * 0.2
    b = torch.randn(40, generator=g)
    xc, wc, bc = [t.clone().cuda().requires_grad_(True) for t in (x, w, b)]
    y = ops.tap_conv_transpose(xc, wc, bc, s, f - s)
    xr, wr, br = [t.clone().requires_grad_(True) for t in (x, w, b)]
    ref = torch.nn.functional.conv_transpose1d(xr, wr, br, stride=s, padding=f - s)
    assert y.shape == ref.shape
    assert float((y.cpu() - ref).abs().max()) / float(ref.abs().max()) < 3e-3
    gy = torch.randn(ref.shape, generator=g)
    (y * gy.cuda()).sum().backward()
    (ref * gy).sum().backward()
    ops.check_device_errors()
    for a, r, name in ((xc.grad, xr.grad, "x"), (wc.grad, wr.grad, "w"), (bc.grad, br.grad, "b")):
        assert float((a.cpu() - r).abs().max()) / float(r.abs().max()) < 5e-3, name


@pytest.mark.parametrize("M,widths", [(256, (256, 256)), (200, (112, 112, 139)), (40, (48,)), (256, (256, 100, 16))])
def test_wgradw_wide_units_match_torch(M, widths):
    """aewn_wgradw: one CTA pair per unit (M <= 256 rows of G) against up to three column chunks with their own operand
    rows, time shifts and (strided / transposed) outputs; per-unit split-K."""
    from aewn import ops, _lib as L
    g = torch.Generator().manual_seed(11)
    B, T, t_lo = 2, 1700, 8
    G = torch.randn(B, M, T, generator=g)
    Gb = ops.to_buf(G.cuda())
    acts, keep = [ops.act_of(Gb, T)], []
    cks, refs = [], []
    for i, n in enumerate(widths):
        X = torch.randn(B, n + 5, T, generator=g)            # operand with extra rows: the chunk starts at row 3
        shift = (0, -8, 4)[i % 3]
        transposed = i % 2 == 1
        out = torch.zeros((n, M) if transposed else (M, n), device="cuda")
        keep.append(ops.to_buf(X.cuda()))                # descriptors hold raw pointers: keep the buffers alive
        acts.append(ops.act_of(keep[-1], T))
        cks.append(dict(x_act=len(acts) - 1, x_row=3, n_valid=n, shift=shift, out=out, out_off=0,
                        out_rs=1 if transposed else n, out_cs=M if transposed else 1))
        Xs = torch.nn.functional.pad(X, (max(0, -shift), max(0, shift)))
        Xs = Xs[:, :, :T] if shift <= 0 else Xs[:, :, shift:shift + T]          # X[u + shift], zero outside [0, T)
        ref = torch.einsum("bmt,bnt->mn", G[:, :, t_lo:], Xs[:, 3:3 + n, t_lo:])
        refs.append((out, ref.t() if transposed else ref))
    units = ops.pack_wide_units(0, 0, M, t_lo, T, cks)
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    for kind, d, tag in ops.build_wgradw(acts, units, B, err):
        L.check(L.lib().aewn_wgradw(C.byref(d), ops._stream()), "aewn_wgradw")
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    for out, ref in refs:
        assert float((out.cpu() - ref).abs().max()) / float(ref.abs().max()) < 3e-3
