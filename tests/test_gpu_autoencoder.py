"""BASELINE cfg3: the VQ-VAE-EMA autoencoder train step (Encoder -> VQEMA -> WaveNet, aewn/autoencoder.py) on the
kernel path.

* small: against the golden written by oracle/ae_harness.py from the UNMODIFIED reference modules (forward values,
  VQ indices, EMA statistics, losses, every parameter gradient and the gradient w.r.t. the mel input);
* full size (par/arch.vqvae-ema.json, batch 16, window 16384): size-independent properties -- VQ indices bit-exact
  against the C oracle on the step's own ze, code counts sum to B*N, z_sum equals the per-code sum of ze, the EMA
  update follows vqema_bn.py:190-195, the loss is finite and an Adam step changes every parameter.

Tolerances: TF32 contractions (see test_gpu_decoder.py): forward 5e-3 of max-abs, gradients 0.15 max-abs / 0.99 cosine
on the tiny fixture (sums of few signed terms)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from test_gpu_decoder import HP, cosine, rel_err  # noqa: E402
from test_gpu_vq_encoder import c_oracle_assign  # noqa: E402


def build_small(g):
    from aewn.autoencoder import AutoEncoder
    ae = AutoEncoder(HP(g["hp"]), g["n_mel"], g["enc_n_out"], "vqvae-ema", g["hp"]["n_lc_in"], 0.25, 0.99, g["K"], True)
    ae.init_geometry(g["W"])
    geo = g["geo"]
    assert (ae.enc_in_len, ae.enc_in_mel_len, ae.embed_len, ae.dec_in_len) == (
        geo["enc_in_len"], geo["enc_in_mel_len"], geo["embed_len"], geo["dec_in_len"])
    assert ae.trim_dec_in.tolist() == geo["trim_dec_in"] and ae.trim_dec_out.tolist() == geo["trim_dec_out"]
    assert ae.decoder.trim_ups_out.tolist() == geo["trim_ups_out"]
    ae.encoder.load_state_dict(g["state_dict"]["encoder"], strict=True)
    ae.bottleneck.load_state_dict(g["state_dict"]["bottleneck"], strict=True)
    ae.decoder.load_state_dict(g["state_dict"]["decoder"], strict=True)
    return ae.cuda().train()


def test_autoencoder_small_step_matches_reference_golden(golden_dir):
    from aewn import ops
    g = torch.load(os.path.join(golden_dir, "autoencoder_small.pt"))
    ae = build_small(g)
    mels = g["mels"].cuda().requires_grad_(True)
    wav_dec = g["wav_dec"].cuda()
    pred, target, com, rec = ae.run(mels, wav_dec, g["spk"].cuda(), g["jit"].cuda())
    bn = ae.bottleneck
    assert rel_err(bn.ze, g["ze"]) < 5e-3
    # indices: equal wherever the reference's own decision is not a near-tie (ze differs by TF32 rounding)
    d = torch.cdist(g["ze"].permute(0, 2, 1), g["state_dict"]["bottleneck"]["emb"].unsqueeze(0).expand(g["ze"].shape[0], -1, -1))
    top2 = d.topk(2, dim=2, largest=False).values
    decided = (top2[..., 1] - top2[..., 0]) > 1e-2 * top2[..., 1]
    assert torch.equal(bn.min_ind.cpu()[decided], g["min_ind"][decided])
    same = torch.equal(bn.min_ind.cpu(), g["min_ind"])
    assert rel_err(pred, g["quant"][..., :-1]) < 5e-3 or not same
    assert abs(float(rec) - float(g["rec"])) < 2e-3 or not same
    assert abs(float(com) - float(g["com"])) < 5e-3 * float(g["com"]) or not same
    assert same, "VQ indices differ from the reference on this fixture"
    for k in ("z_sum", "n_sum", "ema_numer", "ema_denom"):
        assert rel_err(getattr(bn, k), g[k]) < 5e-3, k
    (com + rec).backward()
    ops.check_device_errors()
    assert rel_err(mels.grad, g["mel_grad"]) < 0.15 and cosine(mels.grad, g["mel_grad"]) > 0.995
    mods = dict(encoder=ae.encoder, bottleneck=ae.bottleneck, decoder=ae.decoder)
    worst = []
    for part, grads in g["grads"].items():
        for k, p in mods[part].named_parameters():
            ref = grads[k]
            assert p.grad is not None, (part, k)
            if float(ref.abs().max()) == 0:
                assert float(p.grad.abs().max()) == 0, (part, k)
                continue
            worst.append((rel_err(p.grad, ref), cosine(p.grad, ref), part, k))
    worst.sort(reverse=True)
    print("worst gradient errors", worst[:5])
    # (max-abs bound: the envelope MEASURED in test_gpu_decoder.py -- cuDNN's own TF32 path against the same kind of fp32
    # golden, times two -- lands at this order for sums of few signed terms; the direction bound is the tight one)
    for e, c, part, k in worst:
        assert e < 0.15 and c > 0.995, (part, k, e, c)
    fz = [float(m.frac_zero_act) for m in ae.encoder.net]
    assert max(abs(a - b) for a, b in zip(fz, g["frac_zero"])) < 2e-3


ARCH_VQVAE_EMA = dict(filter_sz=2, n_lc_out=128, lc_upsample_strides=[5, 4, 4, 4], lc_upsample_filt_sizes=[25, 16, 16, 16],
                      n_res=368, n_dil=256, n_skp=256, n_post=256, n_quant=256, n_blocks=2, n_block_layers=10,
                      n_global_embed=10, n_speakers=40, bias=True, n_lc_in=32)


def test_cfg3_full_size_train_step_properties(golden_dir):
    import json
    from aewn import ops
    from aewn.autoencoder import AutoEncoder
    geo = json.load(open(os.path.join(golden_dir, "geometry.json")))["cfg3_vqvae_ema_W16384"]
    torch.manual_seed(2507)
    B, W, K, d = 16, 16384, 4096, 32
    ae = AutoEncoder(HP(ARCH_VQVAE_EMA), 39, 768, "vqvae-ema", d, 0.25, 0.99, K, True)
    ae.init_geometry(W)
    assert (ae.enc_in_len, ae.enc_in_mel_len, ae.embed_len, ae.dec_in_len) == (23280, 144, 65, 18430)
    assert ae.trim_dec_in.tolist() == geo["trim_dec_in"] and ae.trim_dec_out.tolist() == geo["trim_dec_out"]
    ae = ae.cuda().train()
    opt = torch.optim.Adam(ae.parameters(), lr=2e-5, fused=True)          # par/train.basic.json:6, checkpoint.py:49
    g = torch.Generator().manual_seed(1234)
    mels = torch.randn(B, 39, ae.enc_in_mel_len, generator=g).cuda()
    wav_dec = torch.randint(0, 256, (B, ae.dec_in_len), generator=g).float().cuda()
    spk = torch.randint(0, 40, (B,), generator=g).cuda()
    jit = torch.arange(ae.embed_len).unsqueeze(0).repeat(B, 1).cuda()
    bn = ae.bottleneck
    numer0, denom0 = bn.ema_numer.clone(), bn.ema_denom.clone()
    before = [p.detach().clone() for p in ae.parameters()]
    pred, target, com, rec = ae.run(mels, wav_dec, spk, jit)
    assert pred.shape == (B, 256, W - 1) and target.shape == (B, W - 1)
    assert torch.isfinite(com) and torch.isfinite(rec) and 4.0 < float(rec) < 7.5       # ~ln(256) at init
    # --- VQ: bit-exact against the C oracle on the step's own ze
    o_ind, o_dist = c_oracle_assign(bn.ze.detach().float().contiguous(), bn.emb.detach(), 1)
    N = bn.ze.shape[2]
    assert torch.equal(o_ind, bn.min_ind.cpu())
    assert torch.equal(o_dist, bn.min_dist.detach().cpu())
    # --- EMA statistics (vqema_bn.py:172-195)
    assert float(bn.n_sum.sum()) == B * N
    flat = bn.min_ind.flatten()
    z_ref = torch.zeros(K, d, device="cuda").index_add_(0, flat, bn.ze.detach().permute(0, 2, 1).reshape(-1, d))
    assert rel_err(bn.z_sum, z_ref) < 1e-5
    assert rel_err(bn.ema_numer, 0.99 * numer0 + 0.01 * bn.z_sum) < 1e-6
    assert rel_err(bn.ema_denom, 0.99 * denom0 + 0.01 * bn.n_sum) < 1e-6
    # --- oracle VALUES at the full cfg3 size (not only properties).  The encoder works on 144 mel frames, so the CPU
    # oracle runs it for all 16 items: ze within the TF32 envelope, codes equal wherever the oracle's own decision is not
    # a near-tie.  The decoder is checked on the trailing receptive field of item 0, conditioned on the codes the GPU
    # step selected (wav_dec arrives pre-trimmed: wav_cond_offset = [0, dec_in_len]).
    from oracle import torch_oracle as orc
    sd_e = {k: v.detach().cpu() for k, v in ae.encoder.state_dict().items()}
    enc_ref, _ = orc.encoder_forward(sd_e, mels.cpu())
    ze_ref = torch.nn.functional.conv1d(enc_ref, bn.linear.weight.detach().cpu())
    assert rel_err(bn.ze, ze_ref) < 5e-3
    # codes: at initialisation the scaled-L2 distances of a vector to its best codes differ by less than the TF32
    # perturbation of ze, so index equality with the fp32 encoder is not defined; what is: the code the GPU step chose is
    # optimal under the ORACLE's ze up to twice the largest distance perturbation the ze difference causes
    emb_c = bn.emb.detach().cpu()
    dist_ref = orc.scaled_l2(ze_ref, emb_c)                                  # (B, K, N)
    delta = float((orc.scaled_l2(bn.ze.detach().cpu(), emb_c) - dist_ref).abs().max())
    chosen = dist_ref.gather(1, bn.min_ind.cpu().unsqueeze(1)).squeeze(1)
    assert bool((chosen - dist_ref.min(dim=1).values <= 2 * delta + 1e-7).all()) and delta < 5e-3
    sd_d = {k: v.detach().cpu() for k, v in ae.decoder.state_dict().items()}
    T0, rf, n_out = ae.dec_in_len, 2046, 64
    zq0 = ae.encoding_bn.detach().cpu()[0:1]
    cond = orc.conditioning(sd_d, ARCH_VQVAE_EMA, zq0, spk.cpu()[0:1], jit.cpu()[0:1], ae.decoder.trim_ups_out.tolist())
    assert cond.shape[2] == T0
    s0 = T0 - (rf + n_out)
    onehot = torch.nn.functional.one_hot(wav_dec.cpu()[0:1].long(), 256).permute(0, 2, 1).float()[:, :, s0:]
    sig = torch.nn.functional.conv1d(onehot, sd_d["base_layer.weight"], sd_d["base_layer.bias"])
    cw = cond[:, :, s0:]
    skp_sum = 0
    leads = [l.leads.tolist() for l in ae.decoder.conv_layers]
    for li, dd in enumerate(orc.dilations(ARCH_VQVAE_EMA)):
        sig, skp = orc.grcc_layer(sig, cw, orc.sub(sd_d, f"conv_layers.{li}"), dd, leads[li], li == 19)
        skp_sum = skp_sum + skp
    post1 = torch.nn.functional.conv1d(torch.relu(skp_sum), sd_d["post1.weight"], sd_d["post1.bias"])
    ref_q = torch.nn.functional.conv1d(torch.relu(post1), sd_d["post2.weight"], sd_d["post2.bias"])
    got_q = pred.detach()[0:1, :, -(n_out - 1):].cpu()            # pred = quant[..., :-1]
    assert rel_err(got_q, ref_q[:, :, :-1]) < 5e-3
    # --- backward + Adam
    opt.zero_grad(set_to_none=True)
    (com + rec).backward()
    ops.check_device_errors()
    for name, p in ae.named_parameters():
        assert p.grad is not None and bool(torch.isfinite(p.grad).all()), name
    opt.step()
    changed = sum(int(not torch.equal(a, p.detach())) for a, p in zip(before, ae.parameters()))
    assert changed == len(before)
