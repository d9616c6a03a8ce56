"""GPU parity of the VQ / VQ-EMA step (indices bit-exact) and of the encoder conv stack, through the C-ABI kernels."""
import ctypes
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel_err(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-12)


def c_oracle_assign(ze, emb, metric):
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
    ze = np.ascontiguousarray(ze.cpu().numpy())
    emb = np.ascontiguousarray(emb.cpu().numpy())
    B, d, N = ze.shape
    ind = np.zeros(B * N, dtype=np.int64)
    dist = np.zeros(B * N, dtype=np.float32)
    lib.vq_assign_oracle(ze.ctypes.data_as(ctypes.c_void_p), emb.ctypes.data_as(ctypes.c_void_p), B, d, N, emb.shape[0],
                         metric, ind.ctypes.data_as(ctypes.c_void_p), dist.ctypes.data_as(ctypes.c_void_p))
    return torch.from_numpy(ind).reshape(B, N), torch.from_numpy(dist).reshape(B, N)


@pytest.mark.parametrize("which,metric", [("vqema", 1), ("vq", 0)])
def test_vq_kernel_bit_exact_vs_c_oracle_and_reference_indices(golden_dir, which, metric):
    """Same ze, emb as the reference golden: indices equal the reference's, and indices + min_dist are BIT-identical to
    the C oracle (same fixed fp32 order)."""
    from aewn.vqema_bn import _VQAssignFn
    g = torch.load(os.path.join(golden_dir, "vq.pt"))[which]
    ze, emb = g["ze"].cuda(), g["emb"].cuda()
    K, d = emb.shape
    hist, z_sum, n_sum = torch.zeros(K).cuda(), torch.empty(K, d).cuda(), torch.empty(K).cuda()
    zq, min_dist, min_ind, ze_norm = _VQAssignFn.apply(ze, emb, metric, hist, z_sum, n_sum, True)
    torch.cuda.synchronize()
    o_ind, o_dist = c_oracle_assign(g["ze"], g["emb"], metric)
    assert torch.equal(min_ind.cpu(), o_ind)
    assert torch.equal(min_dist.cpu(), o_dist)                     # bit-exact distances
    assert torch.equal(min_ind.cpu(), g["min_ind"])                # == the reference's argmin on these inputs
    assert torch.equal(zq.cpu(), g["emb"][g["min_ind"].flatten()].reshape(*g["min_ind"].shape, d).permute(0, 2, 1))
    cnt = torch.bincount(g["min_ind"].flatten(), minlength=K).float()
    assert torch.equal(hist.cpu(), cnt) and torch.equal(n_sum.cpu(), cnt)
    if which == "vqema":
        assert torch.allclose(z_sum.cpu(), g["z_sum"], atol=1e-5)


def test_vq_kernel_random_seeds_bit_exact_vs_c_oracle():
    from aewn.vqema_bn import _VQAssignFn
    for seed, (B, d, N, K) in enumerate([(16, 32, 65, 4096), (3, 64, 33, 512), (2, 20, 7, 100), (1, 48, 1, 37)]):
        g = torch.Generator().manual_seed(100 + seed)
        ze = torch.randn(B, d, N, generator=g)
        emb = torch.randn(K, d, generator=g) * 0.5
        emb[K // 2] = emb[K // 3]                                  # an exact duplicate code: first index must win
        for metric in (0, 1):
            zq, md, mi, _ = _VQAssignFn.apply(ze.cuda(), emb.cuda(), metric, None, None, None, False)
            o_ind, o_dist = c_oracle_assign(ze, emb, metric)
            assert torch.equal(mi.cpu(), o_ind) and torch.equal(md.cpu(), o_dist), (seed, metric)


def test_vqema_module_matches_reference_golden(golden_dir):
    from aewn import vqema_bn, ops
    g = torch.load(os.path.join(golden_dir, "vq.pt"))["vqema"]
    torch.manual_seed(2507)
    bn = vqema_bn.VQEMA(96, 32, 0.25, 0.99, 4096, True).cuda()
    z = g["z"].cuda().requires_grad_(True)
    out = bn(z)
    assert rel_err(bn.ze, g["ze"]) < 5e-3                          # TF32 1x1 conv
    # indices can differ from the fp32 reference only where TF32 rounding of ze flips a near-tie
    agree = float((bn.min_ind.cpu() == g["min_ind"]).float().mean())
    assert agree > 0.97, agree
    # commitment gradient: compare through ze on the vectors whose code agrees with the fp32 reference (a flipped
    # near-tie legitimately changes that vector's gradient)
    same = (bn.min_ind.cpu() == g["min_ind"])                                     # (B, N)
    (g_ze,) = torch.autograd.grad((bn.min_dist * bn.gamma).sum(), bn.ze, retain_graph=True)
    ze_ref = g["ze"].clone().requires_grad_(True)
    from oracle import torch_oracle as orc
    md_ref, _, _ = orc.vq_assign(ze_ref, g["emb"], "scaled_l2")
    (g_ze_ref,) = torch.autograd.grad((md_ref * 0.25).sum(), ze_ref)
    m = same.unsqueeze(1).expand_as(g_ze_ref)
    err = (g_ze.cpu() - g_ze_ref)[m].abs().max() / g_ze_ref[m].abs().max()
    assert float(err) < 2e-2, float(err)
    z.grad = None
    (out * g["gout"].cuda()).sum().backward()
    ops.check_device_errors()
    assert rel_err(z.grad, g["z_grad_st"]) < 2e-2                  # straight-through: d out / d ze = I
    assert rel_err(bn.linear.weight.grad, g["lin_grad_st"]) < 2e-2
    # every flipped index moves one count between two codes: |d denom|_1 <= 2 * (1 - gamma) * flips
    flips = int((bn.min_ind.cpu() != g["min_ind"]).sum())
    assert float((bn.ema_denom.cpu() - g["ema_denom"]).abs().sum()) <= 2 * 0.01 * flips + 1e-5


def test_encoder_matches_reference_golden(golden_dir):
    from aewn import geometry as vc, wave_encoder, ops
    from oracle import torch_oracle as orc
    g = torch.load(os.path.join(golden_dir, "encoder_small.pt"))
    enc = wave_encoder.Encoder(13, 64, vc.VirtualConv(filter_info=400, stride=160, name="MFCC"))
    enc.load_state_dict(g["state_dict"])
    enc = enc.cuda()
    x = g["x"].cuda().requires_grad_(True)
    y = enc(x)
    assert y.shape == g["y"].shape
    assert rel_err(y, g["y"]) < 5e-3
    fz = [float(enc.metrics[f"enc_az_{i}"]) for i in range(9)]
    assert np.allclose(fz, g["frac_zero"], atol=0.02)
    gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(3)).cuda()
    (y * gy).sum().backward()
    ops.check_device_errors()
    # backward oracle: out-of-place restatement on CPU
    sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in g["state_dict"].items()}
    xc = g["x"].clone().requires_grad_(True)
    yc, _ = orc.encoder_forward(sd, xc)
    (yc * gy.cpu()).sum().backward()
    # A ReLU whose pre-activation is ~0 can flip between TF32 and fp32 arithmetic, which changes the gradient of every
    # element upstream of that unit discretely: require near-perfect direction and a small fraction of outliers.
    def close(a, b, name):
        a, b = a.detach().cpu().double().flatten(), b.detach().cpu().double().flatten()
        cos = float(a @ b / (a.norm() * b.norm()))
        frac_bad = float(((a - b).abs() > 3e-2 * b.abs().max()).double().mean())
        assert cos > 0.995 and frac_bad < 0.07, (name, cos, frac_bad)
    close(x.grad, xc.grad, "x")
    for k, p in enc.named_parameters():
        close(p.grad, sd[k].grad, k)
