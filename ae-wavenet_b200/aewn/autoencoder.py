"""The VQ-VAE(-EMA) autoencoder wiring around the hot-path modules: what autoencoder_model.AutoEncoder does with them
(autoencoder_model.py:44-89 construction, :90-146 geometry, :206-225 forward, :227-259 run).

The reference's own class cannot be constructed at its HEAD (it still calls the pre-refactoring ``WaveNet(**dec_params)``
constructor), so the BASELINE configs 3-5 ("full VQ-VAE-EMA autoencoder train step") need a caller that
wires Encoder -> bottleneck -> WaveNet the way that class intends.  This is that caller and nothing more: no
checkpointing, no k-means codebook initialisation, no data handling (all out of scope).

Differences from the reference class, both deliberate and visible:
  * ``decoder.wav_cond_offset`` is set to ``[0, dec_in_len]`` after ``post_init`` -- ``AutoEncoder.forward`` hands the
    decoder the PRE-TRIMMED ``wav_dec`` (:221-223) whereas ``WaveNet.post_init`` derives an offset relative to the
    encoder-side grid (wavenet.py:272-276); the reference would silently mis-slice.
  * ``VQEMALoss`` / ``VQLoss`` ignore the reconstruction term (vqema_bn.py:244-246); ``run`` returns the commitment
    loss and ``RecLoss`` separately so that a train step can back-propagate their sum.
"""
import torch
from torch import nn

from .compat import vconv
from .vq_bn import VQ
from .vqema_bn import VQEMA, VQEMALoss
from .wave_encoder import Encoder
from .wavenet import RecLoss, WaveNet


class AutoEncoder(nn.Module):
    def __init__(self, hps, n_mel_chan, enc_n_out, bn_type="vqvae-ema", bn_n_out=32, vq_gamma=0.25, vq_ema_gamma=0.99,
                 vq_n_embed=4096, training=True, mfcc_win=400, mfcc_hop=160):
        """hps: decoder hyper-parameters (the ``dec_*`` keys of par/arch.*.json without the prefix, plus n_speakers,
        bias; ``n_lc_in`` is overwritten with bn_n_out, autoencoder_model.py:86)."""
        super().__init__()
        self.mfcc_vc = vconv.VirtualConv(filter_info=mfcc_win, stride=mfcc_hop, parent=None, name="MFCC")
        self.encoder = Encoder(n_mel_chan, enc_n_out, parent_vc=self.mfcc_vc)
        if bn_type == "vqvae-ema":
            self.bottleneck = VQEMA(enc_n_out, bn_n_out, vq_gamma, vq_ema_gamma, vq_n_embed, training)
            self.objective = VQEMALoss(self.bottleneck)
        elif bn_type == "vqvae":
            self.bottleneck = VQ(enc_n_out, bn_n_out, vq_gamma, vq_n_embed)
            self.objective = None        # VQLoss depends on an undefined symbol in the reference
        else:
            raise ValueError('bn_type must be "vqvae-ema" or "vqvae" (ae / vae bottlenecks are out of scope)')
        self.bn_type = bn_type
        hps = type(hps)(hps) if isinstance(hps, dict) else hps
        hps["n_lc_in"] = bn_n_out
        self.decoder = WaveNet(hps, parent_vc=self.encoder.vc["end"])
        self.vc = self.decoder.vc
        self.rec_loss = RecLoss()

    def init_geometry(self, batch_win_size):
        """autoencoder_model.py:95-146."""
        w = batch_win_size
        mfcc_vc = self.encoder.vc["beg"].parent
        end_enc_vc, end_ups_vc = self.encoder.vc["end"], self.decoder.vc["last_upsample"]
        beg_grcc_vc, end_grcc_vc = self.decoder.vc["beg_grcc"], self.decoder.vc["end_grcc"]
        do = vconv.GridRange((0, 10 ** 7), (0, w), 1)
        di = vconv.input_range(beg_grcc_vc, end_grcc_vc, do)
        ei = vconv.input_range(mfcc_vc, end_grcc_vc, do)
        mi = vconv.input_range(mfcc_vc.child, end_grcc_vc, do)
        eo = vconv.output_range(mfcc_vc, end_enc_vc, ei)
        uo = vconv.output_range(mfcc_vc, end_ups_vc, ei)
        self.enc_in_len = ei.sub_length()
        self.enc_in_mel_len = mi.sub_length()
        self.embed_len = eo.sub_length()
        self.dec_in_len = di.sub_length()
        self.trim_dec_in = torch.tensor([di.sub[0] - ei.sub[0], di.sub[1] - ei.sub[0]], dtype=torch.long)
        self.decoder.trim_ups_out = torch.tensor([di.sub[0] - uo.sub[0], di.sub[1] - uo.sub[0]], dtype=torch.long)
        self.trim_dec_out = torch.tensor([do.sub[0] - di.sub[0], do.sub[1] - di.sub[0]], dtype=torch.long)
        self.decoder.post_init(w)
        self.decoder.wav_cond_offset = [0, int(self.dec_in_len)]      # wav_dec arrives pre-trimmed (see module doc)
        self._trim_out = (int(self.trim_dec_out[0]), int(self.trim_dec_out[1]))

    def forward(self, mels, wav_dec, voice_inds, jitter_index):
        """mels (B, M, T_mel), wav_dec (B, dec_in_len) -> quant (B, Q, W)   (autoencoder_model.py:206-225)."""
        encoding = self.encoder(mels)
        self.encoding_bn = self.bottleneck(encoding)
        return self.decoder(wav_dec, self.encoding_bn, voice_inds, jitter_index)

    def run(self, mels, wav_dec, voice_inds, jitter_index):
        """autoencoder_model.py:227-259 without the diagnostic autograd.grad: returns (pred, target, com_loss,
        rec_loss); the objective's metrics dict is refreshed like the reference's."""
        t0, t1 = self._trim_out
        quant = self.forward(mels, wav_dec, voice_inds, jitter_index)
        pred, target = quant[..., :-1], wav_dec[:, t0:t1][..., 1:]
        com = self.objective(pred, target) if self.objective is not None else pred.new_zeros(())
        rec = self.rec_loss(pred, target)
        return pred, target, com, rec
