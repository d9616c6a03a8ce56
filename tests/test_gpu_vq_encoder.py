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
    """VQEMA.forward (vqema_bn.py:125-214) through the module: the projection runs in exact fp32 (aewn_conv1x1_f32), so
    the code indices are IDENTICAL to the reference's, and with them every statistic that is a function of the indices."""
    from aewn import vqema_bn, ops
    from oracle import torch_oracle as orc
    g = torch.load(os.path.join(golden_dir, "vq.pt"))["vqema"]
    torch.manual_seed(2507)
    bn = vqema_bn.VQEMA(96, 32, 0.25, 0.99, 4096, True).cuda()
    assert torch.equal(bn.linear.weight.cpu(), g["lin_w"]) and torch.equal(bn.emb.cpu(), g["emb"])
    z = g["z"].cuda().requires_grad_(True)
    out = bn(z)
    assert rel_err(bn.ze, g["ze"]) < 2e-6                          # fp32 fmaf chain vs the reference's fp32 conv
    assert torch.equal(bn.min_ind.cpu(), g["min_ind"])             # index work: bit-identical (north star)
    assert torch.equal(out.detach().cpu(), g["out"])               # the gathered codes are copies of codebook rows
    assert torch.allclose(bn.min_dist.detach().cpu(), g["min_dist"], rtol=1e-5, atol=1e-7)
    assert torch.equal(bn.ind_hist.cpu(), g["ind_hist"]) and torch.equal(bn.n_sum.cpu(), g["n_sum"])
    assert torch.allclose(bn.z_sum.cpu(), g["z_sum"], atol=1e-5)
    assert torch.allclose(bn.ema_numer.cpu(), g["ema_numer"], atol=1e-6)       # vqema_bn.py:190-195
    assert torch.allclose(bn.ema_denom.cpu(), g["ema_denom"], atol=1e-7)
    assert sorted(bn.uniq.cpu().tolist()) == sorted(g["min_ind"].unique().tolist())
    # commitment gradient (VQEMALoss total = gamma * sum(min_dist), vqema_bn.py:237,246) down to z
    # (the golden's linear.weight.grad holds BOTH backward passes, as the reference run accumulated them)
    (bn.min_dist * bn.gamma).sum().backward(retain_graph=True)
    assert rel_err(z.grad, g["z_grad_commit"]) < 1e-4
    z.grad = None
    (out * g["gout"].cuda()).sum().backward()
    ops.check_device_errors()
    assert rel_err(z.grad, g["z_grad_st"]) < 1e-5                  # straight-through: d out / d ze = I, fp32 projection
    assert rel_err(bn.linear.weight.grad, g["lin_grad_st"]) < 1e-5
    # update_codebook (vqema_bn.py:216-222): emb = ema_numer / ema_denom, detached
    bn.update_codebook()
    ref_emb = g["ema_numer"] / g["ema_denom"].unsqueeze(1)
    assert rel_err(bn.emb, ref_emb) < 1e-4 and not bn.emb.requires_grad     # z_sum is summed with fp32 atomics
    # eval mode (the reference raises UnboundLocalError there, vqema_bn.py:144,212-214; the drop-in returns the codes):
    # nearest code under the NEW codebook, no statistics touched
    hist0, numer0 = bn.ind_hist.clone(), bn.ema_numer.clone()
    bn.eval()
    with torch.no_grad():
        out_e = bn(g["z"].cuda())
    ind_e, _ = c_oracle_assign(bn.ze.detach(), bn.emb, 1)
    assert torch.equal(bn.min_ind.cpu(), ind_e)
    assert torch.equal(out_e.cpu(), bn.emb.cpu()[ind_e.flatten()].reshape(*ind_e.shape, 32).permute(0, 2, 1))
    assert torch.equal(bn.ind_hist, hist0) and torch.equal(bn.ema_numer, numer0)


def test_vq_module_forward_matches_reference_golden(golden_dir):
    """VQ.forward (vq_bn.py:28-61) on CUDA through the module: squared-L2 codes, straight-through gradient, usage
    histogram and the circ_inds ring of recent index sets with its write_pos."""
    from aewn import vq_bn, ops
    g = torch.load(os.path.join(golden_dir, "vq.pt"))["vq"]
    bn = vq_bn.VQ(96, 64, 0.25, 512)
    with torch.no_grad():
        bn.linear.weight.copy_(g["lin_w"])
        bn.emb.copy_(g["emb"])
    bn = bn.cuda()
    z = g["z"].cuda().requires_grad_(True)
    out = bn(z)
    assert rel_err(bn.ze, g["ze"]) < 2e-6
    assert torch.equal(bn.min_ind.cpu(), g["min_ind"])
    assert torch.equal(out.detach().cpu(), g["out"])
    assert torch.allclose(bn.min_dist.detach().cpu(), g["min_dist"], rtol=1e-5, atol=1e-6)
    assert torch.equal(bn.ind_hist.cpu(), g["ind_hist"])
    ni = g["min_ind"].numel()
    assert bn.circ_inds.shape == (100, ni) and bn.write_pos == 1
    assert torch.equal(bn.circ_inds[0].cpu(), g["min_ind"].flatten()) and int((bn.circ_inds[1:] != -1).sum()) == 0
    gout = torch.randn(out.shape, generator=torch.Generator().manual_seed(5)).cuda()
    (out * gout).sum().backward()
    ops.check_device_errors()
    # ReplaceGrad (vq_bn.py:42): the gradient of the output lands on ze unchanged -> z.grad = W^T gout, dW = gout z^T
    w = g["lin_w"][:, :, 0].double()
    assert rel_err(z.grad, torch.einsum("nk,bnt->bkt", w, gout.cpu().double())) < 1e-5
    assert rel_err(bn.linear.weight.grad[:, :, 0], torch.einsum("bnt,bkt->nk", gout.cpu().double(), g["z"].double())) < 1e-5
    assert bn.emb.grad is None or float(bn.emb.grad.abs().max()) == 0.0      # StopGrad on the codebook (vq_bn.py:37)
    with torch.no_grad():
        bn(g["z"].cuda())
    assert bn.write_pos == 2 and torch.equal(bn.circ_inds[1].cpu(), g["min_ind"].flatten())
    assert torch.equal(bn.ind_hist.cpu(), 2 * g["ind_hist"])


def test_stop_grad_and_replace_grad_functions():
    """StopGradFn / ReplaceGradFn (vqema_bn.py:7-45): identity forward; StopGrad returns a zero gradient, ReplaceGrad
    moves the gradient of its first output onto its second input."""
    from aewn.vqema_bn import ReplaceGrad, StopGrad
    a = torch.randn(3, 5, device="cuda", requires_grad=True)
    b = torch.randn(3, 5, device="cuda", requires_grad=True)
    y = StopGrad()(a)
    assert torch.equal(y, a)
    (y * 3.0).sum().backward()
    assert float(a.grad.abs().max()) == 0.0
    a.grad = None
    s, t = ReplaceGrad()(a, b)
    assert torch.equal(s, a) and torch.equal(t, b)
    ws, wt = torch.randn(3, 5, device="cuda"), torch.randn(3, 5, device="cuda")
    ((s * ws).sum() + (t * wt).sum()).backward()
    assert float(a.grad.abs().max()) == 0.0 and torch.allclose(b.grad, ws + wt)


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
