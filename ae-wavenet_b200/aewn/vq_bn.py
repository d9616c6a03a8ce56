"""Drop-in replacement for the reference's vq_bn.py VQ (:7-61) on the fused VQ kernel (squared-L2 metric).
VQLoss (:63-147) depends on a symbol (`L2Error`) that exists nowhere in the reference tree and is
therefore not provided."""
import torch
from torch import nn

from . import ops
from .compat import xavier_init
from .vqema_bn import METRIC_SQ_L2, ReplaceGrad, StopGrad, _VQAssignFn
from .wavenet import _require_cuda


class VQ(nn.Module):
    def __init__(self, n_in, n_out, vq_gamma, vq_n_embed):
        super().__init__()
        self.d = n_out
        self.gamma = vq_gamma
        self.k = vq_n_embed
        self.linear = nn.Conv1d(n_in, self.d, 1, bias=False)
        self.sg = StopGrad()
        self.rg = ReplaceGrad()
        self.ze = None
        self.min_dist = None
        self.register_buffer("ind_hist", torch.zeros(self.k))
        self.circ_inds = None
        self.emb = nn.Parameter(data=torch.empty(self.k, self.d))
        nn.init.xavier_uniform_(self.emb, gain=1)
        xavier_init(self.linear)

    def forward(self, z):
        _require_cuda(z)
        ze = ops.conv1x1_f32(z, self.linear.weight)       # exact fp32 (see vqema_bn.VQEMA.forward)
        self.ze = ze
        zq, min_dist, min_ind, ze_norm = _VQAssignFn.apply(ze, self.emb, METRIC_SQ_L2, self.ind_hist, None, None, True)
        self.min_dist = min_dist
        self.min_ind = min_ind
        # diagnostics, vq_bn.py:44-59: ring buffer of the last 100 index sets
        ni = min_ind.nelement()
        if self.circ_inds is None:
            self.write_pos = 0
            self.circ_inds = ze.new_full((100, ni), -1, dtype=torch.long)
        self.circ_inds[self.write_pos, 0:ni] = min_ind.flatten(0)
        self.circ_inds[self.write_pos, ni:] = -1
        self.write_pos = (self.write_pos + 1) % 100
        self.uniq = min_ind.unique(sorted=False)
        self.ze_norm = ze_norm
        self.emb_norm = (self.emb ** 2).sum(dim=1).sqrt()
        return zq
