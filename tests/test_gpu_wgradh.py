"""aewn_wgradh (csrc/wgradh.cu): the wide-unit weight gradient on fp16 channels-last operands, against a float64 contraction of
the SAME fp16 values.  What is left is fp32 accumulation order over ~1e5 products and the split-K atomics: 2e-5 of the
result's max-abs (a wrong MN-major descriptor, K-block advance or tap shift is an O(1) error)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def run_case(B, T, Mc, Nc, shifts, t_lo, t_hi, seed, scale=1.0):
    from aewn import ops, _lib as L
    dev = torch.device("cuda")
    gen = torch.Generator().manual_seed(seed)
    Tp = (T + 4 + 31) // 32 * 32
    Mp, Np = (Mc + 63) // 64 * 64, (Nc + 63) // 64 * 64
    g = torch.zeros(B, Tp, Mp, dtype=torch.float16, device=dev)
    x = torch.zeros(B, Tp, Np, dtype=torch.float16, device=dev)
    g[:, :T, :Mc] = (torch.randn(B, T, Mc, generator=gen) * 0.5).half().to(dev)
    x[:, :T, :Nc] = torch.randn(B, T, Nc, generator=gen).half().to(dev)
    g[:, t_hi:] = 0                               # contract: rows of the last K block beyond t_hi read as zero
    outs = [torch.zeros(Mc, Nc, len(shifts), device=dev)]
    chunks = []
    for tap, sh in enumerate(shifts):
        for (c0, n) in ops.chunks(Nc):
            chunks.append(dict(x_act=1, x_row=c0, n_valid=n, shift=sh, out=outs[0], out_off=c0 * len(shifts) + tap,
                               out_rs=Nc * len(shifts), out_cs=len(shifts)))
    inv = torch.tensor([1.0 / scale], device=dev)
    launches = []
    for m0 in range(0, Mc, 256):
        units = ops.pack_wide_units(0, m0, min(256, Mc - m0), t_lo, t_hi, [dict(c, out_off=c["out_off"] + m0 * Nc * len(shifts))
                                                                          for c in chunks])
        launches += ops.build_wgradh([ops.act16_of(g), ops.act16_of(x)], units, B, inv.data_ptr(), ops.err_word(dev), tag="t")
    ops.run_launches(launches)
    torch.cuda.synchronize()
    assert int(ops.err_word(dev).item()) == 0
    gd, xd = g.double(), x.double()
    ref = torch.zeros(Mc, Nc, len(shifts), dtype=torch.float64, device=dev)
    for tap, sh in enumerate(shifts):
        lo, hi = max(t_lo, -sh), min(t_hi, Tp - sh)
        ref[:, :, tap] = torch.einsum("btm,btn->mn", gd[:, lo:hi, :Mc], xd[:, lo + sh:hi + sh, :Nc]) / scale
    err = float((outs[0].double() - ref).abs().max()) / float(ref.abs().max())
    return err


@pytest.mark.parametrize("case", [
    # B, T, M channels, N channels, tap shifts, t_lo, t_hi
    (2, 3000, 256, 368, (-4, 0), 7, 3000),            # arch.basic widths, unaligned t_lo, dilation 4
    (1, 700, 512, 144, (0,), 0, 700),                  # two M units, one narrow chunk
    (3, 1100, 256, 80, (-1, 0), 1, 1090),              # dilation 1 (no 16-byte rule), t_hi inside the tensor
    (2, 5000, 128, 512, (-512, 0), 600, 5000),         # m_valid 128, dilation 512
], ids=["basic", "two_units", "dil1", "dil512"])
def test_wgradh_matches_float64_contraction(case):
    assert run_case(*case, seed=5) < 2e-5


def test_wgradh_applies_the_inverse_scale():
    assert run_case(2, 2000, 256, 128, (0,), 0, 2000, seed=9, scale=4096.0) < 2e-5
