"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/* by running the UNMODIFIED reference modules on CPU.

Run in the build container (needs /root/reference):   python oracle/make_golden.py
The fixtures are what pins oracle/torch_oracle.py, oracle/vq_oracle.c and the CUDA path on machines where the
reference tree does not exist (the GPU box).  Shims applied to the reference are listed in oracle/ref_harness.py.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402
from oracle import torch_oracle  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
torch.set_num_threads(8)

SMALL = dict(filter_sz=2, n_lc_out=16, lc_upsample_strides=[5, 4, 4, 4], lc_upsample_filt_sizes=[25, 16, 16, 16],
             n_res=48, n_dil=32, n_skp=24, n_post=40, n_quant=256, n_blocks=2, n_block_layers=4,
             n_global_embed=6, n_speakers=5, bias=True, n_lc_in=12)


def synth_inputs(B, geo, n_lc_in, n_speakers, seed, jitter=False):
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(geo["wav_len"]).float()
    wav = []
    for b in range(B):
        f1, f2 = 0.01 + 0.02 * torch.rand(1, generator=g), 0.05 + 0.1 * torch.rand(1, generator=g)
        x = 0.3 * torch.sin(f1 * t) + 0.2 * torch.sin(f2 * t + 1.0) + 0.05 * torch.randn(t.shape, generator=g)
        x = x.clamp(-1, 1).numpy()
        mu = 255
        amp = np.sign(x) * np.log1p(mu * np.abs(x)) / np.log1p(mu)          # util.mu_encode_np, util.py:62-67
        wav.append(torch.from_numpy(((amp + 1) * 0.5 * mu + 0.5).astype(np.int32)).float())
    wav = torch.stack(wav)
    lc = torch.randn(B, n_lc_in, geo["lc_len"], generator=g)
    spk = torch.randint(0, n_speakers, (B,), generator=g)
    jit = torch.arange(geo["lc_len"]).unsqueeze(0).repeat(B, 1)
    if jitter:   # jitter.py:12-33 style: indices repeat / skip neighbours
        jit = (jit + torch.randint(-1, 2, jit.shape, generator=g)).clamp(0, geo["lc_len"] - 1)
    return wav, lc, spk, jit


def golden_wavenet_small():
    m = rh.load()
    hp = rh.HP(SMALL)
    W, B = 96, 2
    torch.manual_seed(2507)
    wn, geo = rh.standalone_wavenet(hp, W)
    wn.train()
    wav, lc, spk, jit = synth_inputs(B, geo, hp.n_lc_in, hp.n_speakers, 1234, jitter=True)
    lc.requires_grad_(True)
    quant = wn(wav, lc, spk, jit)
    t0, t1 = geo["trim_dec_out"]
    loss = m["wavenet"].RecLoss()(quant[..., :-1], wav[:, t0:t1][..., 1:])
    loss.backward()
    sd = {k: v.clone() for k, v in wn.state_dict().items()}
    grads = {k: p.grad.clone() for k, p in wn.named_parameters()}
    torch.save(dict(hp=dict(hp), W=W, geo=geo, trim_ups_out=wn.trim_ups_out.tolist(), state_dict=sd,
                    wav=wav, lc=lc.detach(), spk=spk, jit=jit, quant=quant.detach(), loss=loss.detach(),
                    grads=grads, lc_grad=lc.grad.clone()), os.path.join(OUT, "wavenet_small.pt"))
    print("wavenet_small: loss", float(loss), "quant", tuple(quant.shape))


def golden_grcc_layer():
    m = rh.load()
    hp = rh.HP(SMALL)
    out = {}
    for name, dil, final in (("d1", 1, False), ("d8", 8, False), ("d2_final", 2, True)):
        torch.manual_seed(11 + dil)
        vc = m["vconv"]
        # a 2-layer chain so that cond/skip leads are non-trivial: [layer under test] -> [tail layer dil 4]
        hp2 = rh.HP(dict(SMALL))
        wvc = {}
        layer = m["wavenet"].GatedResidualCondConv(wvc, hp2, n_cond=22, stride=1, dil=dil, final_layer=final,
                                                   parent_vc=None, name="L")
        if final:
            wvc["beg_grcc"], wvc["end_grcc"] = layer.vc, layer.vc
        else:
            tail = m["wavenet"].GatedResidualCondConv(wvc, hp2, n_cond=22, stride=1, dil=4, final_layer=True,
                                                      parent_vc=layer.vc, name="T")
            wvc["beg_grcc"], wvc["end_grcc"] = layer.vc, tail.vc
        vc.compute_inputs(wvc["end_grcc"], vc.GridRange((0, 10 ** 6), (0, 50), 1))
        layer.post_init()
        T_in = layer.vc.in_len()
        g = torch.Generator().manual_seed(5)
        x = torch.randn(2, hp2.n_res, T_in, generator=g, requires_grad=True)
        cond = torch.randn(2, 22, T_in, generator=g, requires_grad=True)
        sig, skp = layer(x, cond)
        gs, gk = torch.randn(sig.shape, generator=g), torch.randn(skp.shape, generator=g)
        (sig * gs).sum().backward(retain_graph=True) if not final else None
        ((skp * gk).sum() + ((sig * gs).sum() if final else 0)).backward()
        out[name] = dict(dil=dil, final=final, leads=layer.leads.tolist(), x=x.detach(), cond=cond.detach(),
                         state_dict={k: v.clone() for k, v in layer.state_dict().items()}, sig=sig.detach(),
                         skp=skp.detach(), g_sig=gs, g_skp=gk, x_grad=x.grad.clone(), cond_grad=cond.grad.clone(),
                         grads={k: p.grad.clone() for k, p in layer.named_parameters()})
        print("grcc", name, "T_in", T_in, "leads", layer.leads.tolist())
    torch.save(dict(hp=dict(SMALL), cases=out), os.path.join(OUT, "grcc_layer.pt"))


def golden_encoder():
    m = rh.load()
    torch.manual_seed(2507)
    vc = m["vconv"].VirtualConv(filter_info=400, stride=160, name="MFCC")
    enc = m["wave_encoder"].Encoder(13, 64, parent_vc=vc)
    g = torch.Generator().manual_seed(9)
    x = torch.randn(3, 13, 40, generator=g)
    with torch.no_grad():
        y = enc(x)
    torch.save(dict(n_in=13, n_out=64, state_dict={k: v.clone() for k, v in enc.state_dict().items()}, x=x, y=y,
                    frac_zero=[float(mod.frac_zero_act) for mod in enc.net]), os.path.join(OUT, "encoder_small.pt"))
    print("encoder_small:", tuple(y.shape))


def golden_vq():
    m = rh.load()
    res = {}
    # VQEMA (scaled L2), reference shapes: n_in 768 -> d 32, K 4096, B 4, N 65
    torch.manual_seed(2507)
    bn = m["vqema_bn"].VQEMA(96, 32, 0.25, 0.99, 4096, True)
    g = torch.Generator().manual_seed(77)
    z = torch.randn(4, 96, 65, generator=g, requires_grad=True)
    emb0, numer0, denom0 = bn.emb.clone(), bn.ema_numer.clone(), bn.ema_denom.clone()
    with rh.quiet():
        out = bn(z)
    (bn.min_dist * bn.gamma).sum().backward(retain_graph=True)     # VQEMALoss total, vqema_bn.py:237,246
    z_grad_commit = z.grad.clone()
    z.grad = None
    gout = torch.randn(out.shape, generator=g)
    (out * gout).sum().backward()
    res["vqema"] = dict(lin_w=bn.linear.weight.detach().clone(), emb=emb0, ema_numer0=numer0, ema_denom0=denom0,
                        gamma=0.25, ema_gamma=0.99, z=z.detach(), ze=bn.ze.detach(), min_ind=bn.min_ind.clone(),
                        min_dist=bn.min_dist.detach(), out=out.detach(), z_sum=bn.z_sum.clone(), n_sum=bn.n_sum.clone(),
                        ema_numer=bn.ema_numer.clone(), ema_denom=bn.ema_denom.clone(), ind_hist=bn.ind_hist.clone(),
                        z_grad_commit=z_grad_commit, gout=gout, z_grad_st=z.grad.clone(),
                        lin_grad_st=bn.linear.weight.grad.clone())
    # VQ (squared L2), d 64, K 512
    torch.manual_seed(2508)
    vq = m["vq_bn"].VQ(96, 64, 0.25, 512)
    z2 = torch.randn(3, 96, 33, generator=g)
    out2 = vq(z2)
    res["vq"] = dict(lin_w=vq.linear.weight.detach().clone(), emb=vq.emb.detach().clone(), z=z2, ze=vq.ze.detach(),
                     min_dist=vq.min_dist.detach(), out=out2.detach(), ind_hist=vq.ind_hist.clone(),
                     min_ind=out2.new_zeros(0))
    # min_ind is not kept by VQ.forward; recover it from the reference's own distance expression (vq_bn.py:39-40)
    l2 = ((vq.ze.detach().unsqueeze(1) - vq.emb.detach().unsqueeze(2)) ** 2).sum(dim=2)
    res["vq"]["min_ind"] = l2.min(dim=1)[1]
    torch.save(res, os.path.join(OUT, "vq.pt"))
    print("vq: unique codes", int(bn.uniq.numel()), "of", 4 * 65)


def golden_geometry_and_init():
    m = rh.load()
    geo_out = {}
    for name, W, arch in (("cfg2_basic_W16384", 16384, rh.ARCH_BASIC), ("basic_W1024", 1024, rh.ARCH_BASIC),
                          ("small_W96", 96, SMALL)):
        torch.manual_seed(2507)
        wn, geo = rh.standalone_wavenet(rh.HP(arch), W)
        geo_out[name] = geo
        if name == "basic_W1024":
            import hashlib
            dig = {k: hashlib.sha256(v.numpy().tobytes()).hexdigest() for k, v in wn.state_dict().items()
                   if v.dtype == torch.float32}
            json.dump(dig, open(os.path.join(OUT, "init_digest_basic.json"), "w"), indent=0)
    # cfg1: MfccInverter geometry (mfcc_inverter.py:38-65) with hparams defaults, W = 4096
    hps = m["hparams"].setup_hparams("mfcc_inverter,mfcc,train", dict(n_win_batch=4096, n_batch=2))
    torch.manual_seed(2507)
    mi = m["mfcc_inverter"].MfccInverter(hps)
    geo_out["cfg1_mi_W4096"] = dict(enc_in_len=mi.enc_in_len, embed_len=mi.embed_len, dec_in_len=mi.dec_in_len,
                                    trim_dec_in=mi.trim_dec_in.tolist(), trim_dec_out=mi.trim_dec_out.tolist(),
                                    wav_cond_offset=list(mi.wavenet.wav_cond_offset),
                                    leads=[l.leads.tolist() for l in mi.wavenet.conv_layers])
    json.dump(geo_out, open(os.path.join(OUT, "geometry.json"), "w"), indent=1)
    print("geometry:", {k: v.get("dec_in_len") for k, v in geo_out.items()})


def golden_cfg1():
    """BASELINE cfg1: the reference's UNMODIFIED mfcc_inverter.MfccInverter (hparams sets mfcc_inverter,mfcc,train ==
    par/arch.mi.json: upsampling strides [5,4,4,2], 39 conditioning channels), batch 2, window 4096, seed 2507, through
    MfccInverter.run (mfcc_inverter.py:89-107: forward, RecLoss, autograd.grad w.r.t. mel) on CPU.  The 13.5 M initial
    parameters are NOT stored (54 MB): both sides build them from the same seed, pinned here by SHA-256 digests; stored
    are the inputs, the last 64 output steps of the logits, the loss and the mel gradient."""
    import hashlib
    m = rh.load()
    hps = m["hparams"].setup_hparams("mfcc_inverter,mfcc,train", dict(n_win_batch=4096, n_batch=2))
    torch.manual_seed(2507)
    mi = m["mfcc_inverter"].MfccInverter(hps)
    mi.train()
    geo = dict(wav_len=mi.enc_in_len, lc_len=mi.embed_len)
    wav, mel, voice, jit = synth_inputs(2, geo, hps.n_lc_in, hps.n_speakers, 4321)
    mel = mel.clone()
    pred, target, loss = mi.run(wav, mel, voice, jit)
    (mel_grad,) = torch.autograd.grad(loss, mel)
    arch = {k: hps[k] for k in ("filter_sz", "n_lc_out", "lc_upsample_strides", "lc_upsample_filt_sizes", "n_res", "n_dil",
                                "n_skp", "n_post", "n_quant", "n_blocks", "n_block_layers", "n_global_embed", "n_speakers",
                                "bias", "n_lc_in", "mfcc_win_sz", "mfcc_hop_sz")}
    sd = mi.wavenet.state_dict()
    torch.save(dict(arch=arch, W=4096, wav=wav, mel=mel.detach(), voice=voice, jit=jit,
                    trim_dec_out=mi.trim_dec_out.tolist(), wav_cond_offset=list(mi.wavenet.wav_cond_offset),
                    dec_in_len=mi.dec_in_len, enc_in_len=mi.enc_in_len, embed_len=mi.embed_len,
                    pred_tail=pred.detach()[:, :, -64:].clone(), loss=loss.detach(), mel_grad=mel_grad,
                    digest={k: hashlib.sha256(v.numpy().tobytes()).hexdigest() for k, v in sd.items()
                            if v.dtype == torch.float32}),
               os.path.join(OUT, "cfg1_mi.pt"))
    print("cfg1: loss", float(loss), "pred", tuple(pred.shape))


def golden_forward_test():
    """The reference's own sampler (wavenet.py:367-531) on CPU, 2 replicas, with torch.multinomial swapped for an
    inverse-CDF draw on recorded uniforms (generator seed 99) so that the run is reproducible by any implementation.
    Pins the alignment of wav / cond / base_global_rf, the output layout and every per-step distribution."""
    m = rh.load()
    hp = rh.HP(SMALL)
    torch.manual_seed(2507)
    wn, geo = rh.standalone_wavenet(hp, 96)
    wav, lc, spk, jit = synth_inputs(1, geo, hp.n_lc_in, hp.n_speakers, 1234, jitter=True)
    wn.eval()
    wn.set_n_replicas(2)
    g = torch.Generator().manual_seed(99)
    rec = []

    def draw(probs, n, replacement):
        u = torch.rand(probs.shape[0], generator=g)
        idx = torch_oracle.inverse_cdf_draw(probs, u)
        rec.append((probs.clone(), u))
        return idx.unsqueeze(1)

    real = torch.multinomial
    torch.multinomial = draw
    try:
        with torch.no_grad(), rh.quiet():
            out = wn(wav, lc, spk, jit)
    finally:
        torch.multinomial = real
    probs = torch.stack([r[0] for r in rec])
    u = torch.stack([r[1] for r in rec])
    torch.save(dict(hp=dict(hp), W=96, geo=geo, state_dict={k: v.clone() for k, v in wn.state_dict().items()},
                    wav=wav, lc=lc, spk=spk, jit=jit, n_rep=2, base_global_rf=int(wn.base_global_rf),
                    uniforms=u, probs=probs.half(), out=out), os.path.join(OUT, "forward_test.pt"))
    print("forward_test: out", tuple(out.shape), "steps", len(rec))


def golden_autoencoder():
    """cfg3 in miniature: Encoder -> VQEMA -> WaveNet wired by oracle/ae_harness.py from the unmodified reference
    modules; one forward + backward of (commitment + reconstruction) loss.  Also records the cfg3 geometry at the
    BASELINE size (W = 16384) for the product-side wiring (aewn/autoencoder.py) to be checked against."""
    from oracle import ae_harness as ah
    hp = rh.HP(dict(SMALL))
    W, B, n_mel, enc_out, K = 96, 3, 13, 64, 256
    torch.manual_seed(2507)
    enc, bn, dec, geo = ah.build(hp, n_mel, enc_out, hp.n_lc_in, 0.25, 0.99, K, W)
    for mod in (enc, bn, dec):
        mod.train()
    g = torch.Generator().manual_seed(4321)
    fake_geo = dict(wav_len=geo["dec_in_len"], lc_len=geo["embed_len"])
    wav_dec, _, spk, jit = synth_inputs(B, fake_geo, hp.n_lc_in, hp.n_speakers, 4321, jitter=True)
    mels = torch.randn(B, n_mel, geo["enc_in_mel_len"], generator=g, requires_grad=True)
    sd = {name: {k: v.clone() for k, v in mod.state_dict().items()}
          for name, mod in (("encoder", enc), ("bottleneck", bn), ("decoder", dec))}
    quant, com, rec = ah.train_forward(enc, bn, dec, geo, mels, wav_dec, spk, jit)
    (com + rec).backward()
    grads = {name: {k: p.grad.clone() for k, p in mod.named_parameters()}
             for name, mod in (("encoder", enc), ("bottleneck", bn), ("decoder", dec))}
    torch.save(dict(hp=dict(hp), W=W, n_mel=n_mel, enc_n_out=enc_out, K=K, geo=geo, state_dict=sd, mels=mels.detach(),
                    wav_dec=wav_dec, spk=spk, jit=jit, quant=quant.detach(), com=com.detach(), rec=rec.detach(),
                    min_ind=bn.min_ind.clone(), min_dist=bn.min_dist.detach(), ze=bn.ze.detach(),
                    z_sum=bn.z_sum.clone(), n_sum=bn.n_sum.clone(), ema_numer=bn.ema_numer.clone(),
                    ema_denom=bn.ema_denom.clone(), grads=grads, mel_grad=mels.grad.clone(),
                    frac_zero=[float(m_.frac_zero_act) for m_ in enc.net]), os.path.join(OUT, "autoencoder_small.pt"))
    print("autoencoder_small: com", float(com), "rec", float(rec), "uniq", int(bn.min_ind.unique().numel()),
          {k: v for k, v in geo.items() if k != "leads"})
    # cfg3 geometry at full size (par/arch.vqvae-ema.json: 39 mel channels, 768 encoder channels, d = 32, K = 4096)
    torch.manual_seed(2507)
    _, _, _, geo3 = ah.build(rh.HP(dict(rh.ARCH_BASIC, n_lc_in=32)), 39, 768, 32, 0.25, 0.99, 4096, 16384)
    path = os.path.join(OUT, "geometry.json")
    allgeo = json.load(open(path)) if os.path.exists(path) else {}
    allgeo["cfg3_vqvae_ema_W16384"] = geo3
    json.dump(allgeo, open(path, "w"), indent=1)


def golden_loader():
    """Loader arithmetic from the reference's OWN functions (importable: pure numpy / torch): util.mu_encode_np,
    util.mu_decode_np, util.mu_encode_torch, util.mu_decode_torch (util.py:62-96) and jitter.Jitter (jitter.py:3-33) under a
    fixed numpy seed.  (mfcc.ProcessWav needs librosa, which is absent here: no MFCC golden -- see oracle/loader_oracle.py.)"""
    sys.path.insert(0, "/root/reference")
    import jitter as ref_jitter
    import util as ref_util
    rs = np.random.RandomState(7)
    x = np.clip(rs.randn(4096) * 0.25, -1, 1).astype(np.float32)
    x[:9] = [0.0, 1.0, -1.0, 1e-4, -1e-4, 0.5, -0.5, 0.999, -0.999]
    q = np.arange(256, dtype=np.int32)
    out = dict(x=torch.from_numpy(x), enc_np=torch.from_numpy(ref_util.mu_encode_np(x, 256)),
               enc_torch=ref_util.mu_encode_torch(torch.from_numpy(x), 256),
               dec_np=torch.from_numpy(ref_util.mu_decode_np(q, 256).astype(np.float32)),
               dec_torch=ref_util.mu_decode_torch(torch.from_numpy(q).long(), 256), jitter=[])
    for seed, prob, win in ((0, 0.12, 58), (3, 0.12, 117), (11, 0.3, 40)):
        np.random.seed(seed)
        idx = ref_jitter.Jitter(prob)(win)
        np.random.seed(seed)
        u = np.random.random_sample(win - 2)          # the draws Jitter.__call__ consumed
        out["jitter"].append(dict(seed=seed, prob=prob, win=win, index=torch.from_numpy(idx.astype(np.int64)),
                                  uniforms=torch.from_numpy(u)))
    torch.save(out, os.path.join(OUT, "loader.pt"))
    print("loader golden written")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "loader":
        golden_loader()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "cfg1":
        golden_cfg1()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "autoencoder":
        golden_autoencoder()
        sys.exit(0)
    golden_wavenet_small()
    golden_forward_test()
    golden_grcc_layer()
    golden_encoder()
    golden_vq()
    golden_geometry_and_init()
    golden_autoencoder()
    golden_loader()
    print("goldens written to", OUT)
