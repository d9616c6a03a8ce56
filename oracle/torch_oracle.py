"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain PyTorch fp32, functional style) of the reference hot path.

This is the oracle that travels to the GPU box (the reference tree does not).  Each function cites the reference lines
it restates.  It is pinned two ways (tests/test_oracle.py): against the real reference modules when /root/reference
is present, and against the golden vectors under tests/golden/ that oracle/make_golden.py produced FROM the real
reference.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import it; the product package
never does (it has no CPU compute path at all).

All tensors are float32 on the device of the inputs (CPU in the tests; bench.py's GPU-library baseline runs the same
functions on cuda, i.e. eager PyTorch/cuDNN); weights come in a flat dict keyed like the reference state_dict
(SURVEY.md 8b), e.g. ``conv_layers.3.conv_signal.weight``.
"""
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------ decoder
def grcc_layer(x, cond, p, dil, leads, final_layer):
    """GatedResidualCondConv.forward, wavenet.py:91-111.  ``p``: dict with conv_signal.weight/.bias, conv_gate.*,
    proj_signal.weight, proj_gate.weight, dil_skp.weight[, dil_res.weight]; ``leads`` = (cond_lead, skip_lead, lw)."""
    cl, sl, lw = int(leads[0]), int(leads[1]), int(leads[2])
    c = cond[:, :, cl:]
    filt = F.conv1d(x, p["conv_signal.weight"], p.get("conv_signal.bias"), dilation=dil) + \
        F.conv1d(c, p["proj_signal.weight"])
    gate = F.conv1d(x, p["conv_gate.weight"], p.get("conv_gate.bias"), dilation=dil) + \
        F.conv1d(c, p["proj_gate.weight"])
    z = torch.tanh(filt) * torch.sigmoid(gate)
    skp = F.conv1d(z[:, :, sl:], p["dil_skp.weight"])
    if final_layer:
        sig = x[:, :, lw:]
    else:
        sig = F.conv1d(z, p["dil_res.weight"]) + x[:, :, lw:]
    return sig, skp


def sub(sd, prefix):
    """view of a flat state dict below ``prefix.``"""
    n = len(prefix) + 1
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix + ".")}


def jitter_gather(lc_sparse, jitter_index):
    """wavenet.py:330-336: torch.take on the FLAT lc_sparse with an index that only carries the batch
    offset, i.e. lc_jitter[b, c, t] = lc_sparse.flatten()[jitter[b, t] + b * T]  for every channel c."""
    B, D1, T = lc_sparse.shape
    flat_idx = jitter_index + (torch.arange(B, device=jitter_index.device) * jitter_index.shape[1]).unsqueeze(1)
    g = lc_sparse.reshape(-1)[flat_idx]  # (B, T)
    return g.unsqueeze(1).expand(-1, D1, -1)


def conditioning(sd, hp, lc_sparse, speaker_inds, jitter_index, trim_ups_out):
    """wavenet.py:330-343: jitter -> lc_conv -> 4x ConvTranspose1d -> trim -> concat speaker embedding."""
    lc = jitter_gather(lc_sparse, jitter_index)
    lc = F.conv1d(lc, sd["lc_conv.weight"], sd.get("lc_conv.bias"))
    for i, (f, s) in enumerate(zip(hp["lc_upsample_filt_sizes"], hp["lc_upsample_strides"])):
        lc = F.conv_transpose1d(lc, sd[f"lc_upsample.{i}.tconv.weight"], sd.get(f"lc_upsample.{i}.tconv.bias"),
                                stride=s, padding=f - s)  # wavenet.py:154-155
    if trim_ups_out is not None:                                          # forward_test does not trim (:385-391)
        lc = lc[:, :, int(trim_ups_out[0]):int(trim_ups_out[1])]
    one_hot = F.one_hot(speaker_inds.long(), hp["n_speakers"]).to(lc.dtype)  # wavenet.py:135
    gc = F.linear(one_hot, sd["cond.speaker_embedding.weight"], sd.get("cond.speaker_embedding.bias"))
    return torch.cat((lc, gc.unsqueeze(2).expand(-1, -1, lc.shape[2])), dim=1)


def dilations(hp):
    return [2 ** bl for _ in range(hp["n_blocks"]) for bl in range(hp["n_block_layers"])]  # wavenet.py:234-236


def wavenet_forward_train(sd, hp, geo, wav, lc_sparse, speaker_inds, jitter_index, return_intermediates=False):
    """WaveNet.forward_train, wavenet.py:323-364.  geo: dict(trim_ups_out, wav_cond_offset, leads[L][4], n_win_batch)."""
    cond = conditioning(sd, hp, lc_sparse, speaker_inds, jitter_index, geo["trim_ups_out"])
    o0, o1 = geo["wav_cond_offset"]
    wav_onehot = F.one_hot(wav.long(), hp["n_quant"]).permute(0, 2, 1).float()[:, :, o0:o1]  # :348-349
    sig = F.conv1d(wav_onehot, sd["base_layer.weight"], sd.get("base_layer.bias"))
    skp_sum = torch.zeros(wav.shape[0], hp["n_skp"], geo["n_win_batch"], device=wav.device)
    dils = dilations(hp)
    inter = []
    for li, d in enumerate(dils):
        sig, skp = grcc_layer(sig, cond, sub(sd, f"conv_layers.{li}"), d, geo["leads"][li], li == len(dils) - 1)
        skp_sum = skp_sum + skp
        if return_intermediates:
            inter.append(sig)
    post1 = F.conv1d(F.relu(skp_sum), sd["post1.weight"], sd.get("post1.bias"))
    quant = F.conv1d(F.relu(post1), sd["post2.weight"], sd.get("post2.bias"))
    if return_intermediates:
        return quant, dict(cond=cond, sigs=inter, skp_sum=skp_sum)
    return quant


def stack_window_logits(sd, hp, codes, cond):
    """Logits of ONE output step from a full receptive-field window (wavenet.py:459-477 in 'full' mode):
    codes (n, RF+1) long, cond (n, C, RF+1)  ->  (n, Q).  Leads follow init_leads (wavenet.py:53-75) for a one-step
    output: cond_lead_l = sum_{j<=l} d_j, skip_lead_l = RF - cond_lead_l, lw = d_l."""
    dils = dilations(hp)
    rf = sum(dils)
    assert codes.shape[1] == rf + 1 and cond.shape[2] == rf + 1
    sig = F.conv1d(F.one_hot(codes, hp["n_quant"]).permute(0, 2, 1).to(cond.dtype), sd["base_layer.weight"],
                   sd.get("base_layer.bias"))
    skp_sum, lead = 0, 0
    for li, d in enumerate(dils):
        lead += d
        sig, skp = grcc_layer(sig, cond, sub(sd, f"conv_layers.{li}"), d, (lead, rf - lead, d), li == len(dils) - 1)
        skp_sum = skp_sum + skp
    post1 = F.conv1d(F.relu(skp_sum), sd["post1.weight"], sd.get("post1.bias"))
    return F.conv1d(F.relu(post1), sd["post2.weight"], sd.get("post2.bias")).squeeze(2)


def inverse_cdf_draw(probs, u):
    """Index of the first bin whose cumulative probability exceeds u * total; probs (n, Q), u (n,).  Distributed like
    torch.multinomial(probs, 1) (wavenet.py:479), but reproducible from the uniforms."""
    cdf = probs.double().cumsum(-1)
    idx = (cdf <= (u.double() * cdf[:, -1]).unsqueeze(1)).sum(-1)
    return idx.clamp(max=probs.shape[-1] - 1)


def wavenet_forward_test(sd, hp, wav_cond_offset, wav, lc_sparse, speaker_inds, jitter_index, n_rep, uniforms,
                         n_steps=None):
    """WaveNet.forward_test, wavenet.py:367-531, restated without the ring-buffer bookkeeping: every step evaluates the
    stack on the last RF+1 samples (what the reference's incremental buffers hold, :455-509).  wav (1, T_wav) codes;
    uniforms (steps, n_rep) replace torch.multinomial's RNG.  Returns (wav_out (n_rep+1, T) float, probs (steps, n_rep, Q))
    with row 0 = the input and rows 1.. = input up to base_global_rf, then drawn samples up to the end of cond."""
    assert wav.shape[0] == 1
    rf1 = sum(dilations(hp)) + 1                                           # base_global_rf (:298, vc.in_len())
    cond = conditioning(sd, hp, lc_sparse, speaker_inds, jitter_index, None)   # (1, C, n_ts)
    n_ts = cond.shape[2]
    codes = wav[0, int(wav_cond_offset[0]):].long()
    out = codes.unsqueeze(0).repeat(n_rep + 1, 1)
    end = n_ts if n_steps is None else min(n_ts, rf1 + n_steps)
    probs_all = []
    for cur in range(rf1, end):                                            # :455; the draw lands at index cur (:480)
        win = out[1:, cur - rf1:cur]
        logits = stack_window_logits(sd, hp, win, cond[:, :, cur - rf1:cur].expand(n_rep, -1, -1))
        probs = F.softmax(logits, dim=-1)                                  # :478
        out[1:, cur] = inverse_cdf_draw(probs, uniforms[cur - rf1])
        probs_all.append(probs)
    return out.float(), torch.stack(probs_all)


def rec_loss(quant_pred, target_wav):
    """RecLoss.forward, wavenet.py:541-552."""
    log_pred = F.log_softmax(quant_pred, dim=1)
    return -torch.gather(log_pred, 1, target_wav.long().unsqueeze(1)).mean()


def decoder_loss(sd, hp, geo, wav, lc_sparse, speaker_inds, jitter_index):
    """MfccInverter.run alignment, mfcc_inverter.py:95-101: pred = quant[..., :-1], target = wav[trim_dec_out][1:]."""
    quant = wavenet_forward_train(sd, hp, geo, wav, lc_sparse, speaker_inds, jitter_index)
    t0, t1 = geo["trim_dec_out"]
    return rec_loss(quant[..., :-1], wav[:, t0:t1][..., 1:]), quant


# ------------------------------------------------------------------------------------------------ encoder
ENC_FILTERS = [3, 3, 4, 3, 3, 1, 1, 1, 1]     # wave_encoder.py:59
ENC_STRIDES = [1, 1, 2, 1, 1, 1, 1, 1, 1]     # wave_encoder.py:60
ENC_RESIDUAL = [False, True, False, True, True, True, True, True, True]  # wave_encoder.py:61


def conv_relu_res(x, w, b, stride, do_res):
    """ConvReLURes.forward, wave_encoder.py:34-50, in the OUT-OF-PLACE form (the reference's in-place
    add on the ReLU output breaks autograd).  Returns (act, frac_zero_act)."""
    k = w.shape[2]
    act = F.relu(F.conv1d(x, w, b, stride=stride))
    if do_res:
        lw = (k - 1) // 2            # vconv.VirtualConv wings for an integer filter size
        rw = (k - 1) - lw
        act = act + x[:, :, lw:x.shape[2] - rw]
    frac_zero = (act == 0.0).sum().double() / act.nelement()
    return act, frac_zero


def encoder_forward(sd, x):
    """Encoder.forward, wave_encoder.py:94-103 (9 ConvReLURes layers)."""
    fracs = []
    for i in range(9):
        x, fz = conv_relu_res(x, sd[f"net.{i}.conv.weight"], sd[f"net.{i}.conv.bias"], ENC_STRIDES[i], ENC_RESIDUAL[i])
        fracs.append(fz)
    return x, fracs


# ------------------------------------------------------------------------------------------------ VQ / VQ-EMA
def scaled_l2(ze, emb):
    """scaled_l2_norm, vqema_bn.py:67-76, evaluated exactly as VQEMA.forward does (vqema_bn.py:138-139): broadcast
    (B,1,d,N) against (1,K,d,1), reduce over d.  Returns (B, K, N)."""
    z = ze.unsqueeze(1)
    q = emb.unsqueeze(2).unsqueeze(0)
    num = ((z - q) ** 2).sum(dim=2).sqrt()
    den = (z ** 2).sum(dim=2).sqrt() + (q ** 2).sum(dim=2).sqrt()
    return num / den


def sq_l2(ze, emb):
    """vq_bn.py:39: ((ze.unsqueeze(1) - emb.unsqueeze(2)) ** 2).sum(dim=2) -> (B, K, N)."""
    return ((ze.unsqueeze(1) - emb.unsqueeze(2)) ** 2).sum(dim=2)


def vq_assign(ze, emb, metric):
    """Nearest code per (b, n): returns (min_dist (B,N), min_ind (B,N) int64, zq (B,d,N)); first index wins ties."""
    dist = scaled_l2(ze, emb) if metric == "scaled_l2" else sq_l2(ze, emb)
    min_dist, min_ind = dist.min(dim=1)
    zq = emb.index_select(0, min_ind.flatten()).reshape(*min_ind.shape, emb.shape[1]).permute(0, 2, 1)
    return min_dist, min_ind, zq


def vqema_stats(ze, min_ind, k):
    """EMA statistics, vqema_bn.py:172-188: z_sum[k,:] = sum of ze vectors assigned to k, n_sum[k] = their count."""
    d = ze.shape[1]
    flat = min_ind.flatten()
    z_sum = torch.zeros(k, d, device=ze.device).index_add_(0, flat, ze.permute(0, 2, 1).reshape(-1, d))
    n_sum = torch.zeros(k, device=ze.device).index_add_(0, flat, torch.ones(flat.numel(), device=ze.device))
    return z_sum, n_sum


def vqema_forward(lin_w, emb, z, ema_numer, ema_denom, ema_gamma):
    """VQEMA.forward in training mode, vqema_bn.py:125-214.  Returns dict with ze, min_dist, min_ind, zq, z_sum, n_sum,
    new ema_numer/ema_denom.  Straight-through: output value = zq, d(out)/d(ze) = I (ReplaceGrad, :33-45)."""
    ze = F.conv1d(z, lin_w)
    min_dist, min_ind, zq = vq_assign(ze, emb, "scaled_l2")
    z_sum, n_sum = vqema_stats(ze.detach(), min_ind, emb.shape[0])
    out = zq.detach() + (ze - ze.detach())  # value == zq exactly, d(out)/d(ze) = I
    return dict(ze=ze, min_dist=min_dist, min_ind=min_ind, zq=zq, out=out, z_sum=z_sum, n_sum=n_sum,
                ema_numer=ema_gamma * ema_numer + (1.0 - ema_gamma) * z_sum,
                ema_denom=ema_gamma * ema_denom + (1.0 - ema_gamma) * n_sum)


def vq_forward(lin_w, emb, z):
    """VQ.forward, vq_bn.py:28-61 (squared-L2 metric; emb is a Parameter but enters through StopGrad)."""
    ze = F.conv1d(z, lin_w)
    min_dist, min_ind, zq = vq_assign(ze, emb.detach(), "sq_l2")
    out = zq.detach() + (ze - ze.detach())  # value == zq exactly, d(out)/d(ze) = I
    return dict(ze=ze, min_dist=min_dist, min_ind=min_ind, zq=zq, out=out)


# ------------------------------------------------------------------------------------------------ autoencoder (cfg3)
def autoencoder_step(sd, hp, geo, mels, wav_dec, speaker_inds, jitter_index, vq_gamma, ema_gamma):
    """The VQ-VAE-EMA train step as oracle/ae_harness.py wires it from the reference modules
    (autoencoder_model.py:206-259): Encoder -> VQEMA -> WaveNet on the pre-trimmed wav_dec, losses
    com = vq_gamma * sum(min_dist) (VQEMALoss total, vqema_bn.py:237,246) and rec = RecLoss (wavenet.py:541-552).
    sd: dict(encoder=..., bottleneck=..., decoder=...) of state dicts (tensors that need gradients must require them).
    geo: ae_harness geometry (trim_ups_out, dec_in_len, trim_dec_out, leads).  Returns a dict."""
    encoding, fracs = encoder_forward(sd["encoder"], mels)
    bsd = sd["bottleneck"]
    vq = vqema_forward(bsd["linear.weight"], bsd["emb"], encoding, bsd["ema_numer"], bsd["ema_denom"], ema_gamma)
    dgeo = dict(trim_ups_out=geo["trim_ups_out"], wav_cond_offset=[0, geo["dec_in_len"]], leads=geo["leads"],
                n_win_batch=geo["trim_dec_out"][1] - geo["trim_dec_out"][0])
    quant = wavenet_forward_train(sd["decoder"], hp, dgeo, wav_dec, vq["out"], speaker_inds, jitter_index)
    t0, t1 = geo["trim_dec_out"]
    pred, target = quant[..., :-1], wav_dec[:, t0:t1][..., 1:]
    com = (vq["min_dist"] * vq_gamma).sum()
    rec = rec_loss(pred, target)
    return dict(quant=quant, com=com, rec=rec, frac_zero=fracs, **{k: vq[k] for k in
                ("ze", "min_ind", "min_dist", "z_sum", "n_sum", "ema_numer", "ema_denom")})
