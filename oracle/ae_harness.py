"""TEST INFRASTRUCTURE ONLY -- the VQ-VAE-EMA autoencoder train step (BASELINE cfg3 / cfg4) wired from the UNMODIFIED
reference modules on CPU.

`autoencoder_model.AutoEncoder` cannot be constructed at the reference's HEAD (it still calls the old
`WaveNet(**dec_params)` constructor), so this harness does by hand what `AutoEncoder._initialize`
(autoencoder_model.py:44-89), `post_init` / `_init_geometry` (:90-146), `forward` (:206-225) and `run` (:227-259) do,
with the current `WaveNet(hps, parent_vc)` constructor:

    mfcc_vc = VirtualConv(filter_info=400, stride=160)                         (data.py: the MFCC window / hop)
    encoder = Encoder(n_mel_chan, enc_n_out, parent_vc=mfcc_vc)                (autoencoder_model.py:52, :91)
    bn      = VQEMA(enc_n_out, bn_n_out, vq_gamma, vq_ema_gamma, K, training)  (:64-66)
    decoder = WaveNet(hps, parent_vc=encoder.vc['end'])                        (:83-87)
    geometry: do, di, ei, mi, eo, uo and the three trims                       (:119-146)
    decoder.post_init(W);  decoder.wav_cond_offset = [0, dec_in_len]           (AutoEncoder.forward
                                                                                passes the PRE-TRIMMED wav_dec, :221-223)
    step: encoding = encoder(mels); enc_bn = bn(encoding); quant = decoder(wav_dec, enc_bn, voice, jitter)
          loss = VQEMALoss(quant[..., :-1], wav_dec[:, trim_dec_out][..., 1:])  (commitment term only, vqema_bn.py:246)
               + RecLoss(...)  (added explicitly so the decoder and encoder get a reconstruction gradient)

Only usable where /root/reference exists; oracle/make_golden.py uses it to write tests/golden/autoencoder_small.pt and
the cfg3 geometry.  Shims: those of oracle/ref_harness.py (F4/F5/F6) plus the out-of-place residual add below (F7).
"""
import torch

from . import ref_harness as rh


def _patch_encoder_residual(wave_encoder):
    """F7: ConvReLURes.forward adds the residual IN PLACE into the ReLU output (wave_encoder.py:41-44), which makes
    autograd raise in backward.  The oracle-side forward below is the same arithmetic out of place;
    its forward values are bit-identical to the reference's (checked in tests/test_oracle.py)."""
    if getattr(wave_encoder.ConvReLURes, "_aewn_out_of_place", False):
        return

    def forward(self, x):
        pre = self.conv(x)
        act = self.relu(pre)
        if self.do_res:
            l_off, r_off = int(self.residual_offsets[0]), int(self.residual_offsets[1])
            act = act + x[:, :, l_off:r_off or None]
        self.frac_zero_act = (act == 0.0).sum().double() / act.nelement()
        return act

    wave_encoder.ConvReLURes.forward = forward
    wave_encoder.ConvReLURes._aewn_out_of_place = True


def build(hps, n_mel_chan, enc_n_out, bn_n_out, vq_gamma, vq_ema_gamma, vq_n_embed, n_win_batch):
    """Returns (encoder, bn, decoder, geo) built from the reference modules; geo follows _init_geometry."""
    m = rh.load()
    vconv, we, vqema, wavenet = m["vconv"], m["wave_encoder"], m["vqema_bn"], m["wavenet"]
    _patch_encoder_residual(we)
    mfcc_vc = vconv.VirtualConv(filter_info=400, stride=160, parent=None, name="MFCC")
    encoder = we.Encoder(n_mel_chan, enc_n_out, parent_vc=mfcc_vc)
    bn = vqema.VQEMA(enc_n_out, bn_n_out, vq_gamma, vq_ema_gamma, vq_n_embed, True)
    decoder = wavenet.WaveNet(hps, parent_vc=encoder.vc["end"])

    w = n_win_batch
    end_enc_vc, end_ups_vc = encoder.vc["end"], decoder.vc["last_upsample"]
    beg_grcc_vc, end_grcc_vc = decoder.vc["beg_grcc"], decoder.vc["end_grcc"]
    do = vconv.GridRange((0, 100000 if w < 50000 else 10 ** 7), (0, w), 1)
    di = vconv.input_range(beg_grcc_vc, end_grcc_vc, do)
    ei = vconv.input_range(mfcc_vc, end_grcc_vc, do)
    mi = vconv.input_range(mfcc_vc.child, end_grcc_vc, do)
    eo = vconv.output_range(mfcc_vc, end_enc_vc, ei)
    uo = vconv.output_range(mfcc_vc, end_ups_vc, ei)
    geo = dict(enc_in_len=ei.sub_length(), enc_in_mel_len=mi.sub_length(), embed_len=eo.sub_length(),
               dec_in_len=di.sub_length(),
               trim_dec_in=[di.sub[0] - ei.sub[0], di.sub[1] - ei.sub[0]],
               trim_ups_out=[di.sub[0] - uo.sub[0], di.sub[1] - uo.sub[0]],
               trim_dec_out=[do.sub[0] - di.sub[0], do.sub[1] - di.sub[0]])
    decoder.trim_ups_out = torch.tensor(geo["trim_ups_out"], dtype=torch.long)
    decoder.post_init(w)
    geo["wav_cond_offset_post_init"] = [int(v) for v in decoder.wav_cond_offset]
    decoder.wav_cond_offset = [0, geo["dec_in_len"]]           # pre-trimmed wav_dec convention
    geo["leads"] = [layer.leads.tolist() for layer in decoder.conv_layers]
    return encoder, bn, decoder, geo


def train_forward(encoder, bn, decoder, geo, mels, wav_dec, voice, jitter):
    """One forward of the harness; returns (quant, com_loss, rec_loss).  total = com_loss + rec_loss."""
    m = rh.load()
    encoding = encoder(mels)
    with rh.quiet():
        enc_bn = bn(encoding)
    quant = decoder(wav_dec, enc_bn, voice, jitter)
    t0, t1 = geo["trim_dec_out"]
    pred, target = quant[..., :-1], wav_dec[:, t0:t1][..., 1:]
    com = m["vqema_bn"].VQEMALoss(bn)(pred, target)
    rec = m["wavenet"].RecLoss()(pred, target)
    return quant, com, rec
