"""Drop-in replacements for the reference's vqema_bn.py: StopGrad/ReplaceGrad (:7-64), scaled_l2_norm (:67-76),
VQEMA (:79-222), VQEMALoss (:225-266) -- with the nearest-code search, gather, usage histogram and EMA statistics in
ONE fused kernel (aewn_vq_fwd) instead of two (B, K, d, N) temporaries."""
import ctypes as C

import torch
from torch import nn

from . import _lib as L
from . import ops
from .compat import xavier_init
from .wavenet import _require_cuda


class StopGradFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src):
        return src

    @staticmethod
    def backward(ctx, src):
        return src.new_zeros(src.size())


class StopGrad(nn.Module):
    def forward(self, src):
        return StopGradFn.apply(src)


class ReplaceGradFn(torch.autograd.Function):
    """forward: identity on (src, trg); backward: src gets zero, trg gets g_src + g_trg  (vqema_bn.py:33-45)."""

    @staticmethod
    def forward(ctx, src, trg):
        assert src.size() == trg.size()
        return src, trg

    @staticmethod
    def backward(ctx, src_grad, trg_grad):
        return src_grad.new_zeros(src_grad.size()), src_grad + trg_grad


class ReplaceGrad(nn.Module):
    def forward(self, src, trg):
        return ReplaceGradFn.apply(src, trg)


def scaled_l2_norm(z, q):
    """vqema_bn.py:67-76 (kept for callers that import it; the kernel path does not use it)."""
    num = ((z - q) ** 2).sum(dim=2).sqrt()
    den = (z ** 2).sum(dim=2).sqrt() + (q ** 2).sum(dim=2).sqrt()
    return num / den


METRIC_SQ_L2, METRIC_SCALED_L2 = 0, 1


class _VQAssignFn(torch.autograd.Function):
    """(ze (B,d,N), emb (K,d)) -> zq (B,d,N) [straight-through: d zq / d ze = I], min_dist (B,N) [differentiable
    w.r.t. ze through the commitment-gradient kernel]; emb receives no gradient (StopGrad, vqema_bn.py:133)."""

    @staticmethod
    def forward(ctx, ze, emb, metric, hist, z_sum, n_sum, want_norm):
        B, d, N = ze.shape
        K = emb.shape[0]
        zec = ze.detach().contiguous()
        embc = emb.detach().contiguous()
        dev = ze.device
        min_ind = torch.empty(B, N, dtype=torch.int64, device=dev)
        min_dist = torch.empty(B, N, device=dev)
        zq = torch.empty(B, d, N, device=dev)
        ze_norm = torch.empty(B, N, device=dev) if want_norm else None
        vp = lambda t: C.c_void_p(t.data_ptr() if t is not None else None)
        L.check(L.lib().aewn_vq_fwd(
            vp(zec), C.c_longlong(zec.stride(0)), C.c_longlong(zec.stride(1)), vp(embc), C.c_int(metric), vp(min_ind),
            vp(min_dist), vp(zq), C.c_longlong(zq.stride(0)), C.c_longlong(zq.stride(1)), vp(hist), vp(z_sum), vp(n_sum),
            vp(ze_norm), C.c_int(B), C.c_int(d), C.c_int(N), C.c_int(K), ops._stream()), "aewn_vq_fwd")
        ctx.save_for_backward(zec, embc, min_ind)
        ctx.metric = metric
        ctx.mark_non_differentiable(min_ind)
        if ze_norm is not None:
            ctx.mark_non_differentiable(ze_norm)
        return zq, min_dist, min_ind, ze_norm

    @staticmethod
    def backward(ctx, g_zq, g_min, _gi, _gn):
        zec, embc, min_ind = ctx.saved_tensors
        B, d, N = zec.shape
        g_ze = g_zq.contiguous().clone() if g_zq is not None else torch.zeros_like(zec)
        if g_min is not None:
            gm = g_min.contiguous()
            vp = lambda t: C.c_void_p(t.data_ptr())
            L.check(L.lib().aewn_vq_commit_bwd(
                vp(zec), C.c_longlong(zec.stride(0)), C.c_longlong(zec.stride(1)), vp(embc), vp(min_ind), vp(gm),
                C.c_int(ctx.metric), vp(g_ze), C.c_longlong(g_ze.stride(0)), C.c_longlong(g_ze.stride(1)), C.c_int(1),
                C.c_int(B), C.c_int(d), C.c_int(N), ops._stream()), "aewn_vq_commit_bwd")
        return g_ze, None, None, None, None, None, None


class VQEMA(nn.Module):
    """vqema_bn.py:79-222.  Buffers, attributes and EMA semantics follow the reference; `defer_ema` (new) lets the
    data-parallel engine all-reduce z_sum / n_sum before the EMA line (mathematically identical)."""

    def __init__(self, n_in, n_out, vq_gamma, vq_ema_gamma, vq_n_embed, training):
        super().__init__()
        self.training = training
        self.d = n_out
        self.gamma = vq_gamma
        self.ema_gamma = vq_ema_gamma
        self.ema_gamma_comp = 1.0 - self.ema_gamma
        self.k = vq_n_embed
        self.linear = nn.Conv1d(n_in, self.d, 1, bias=False)
        self.sg = StopGrad()
        self.rg = ReplaceGrad()
        self.ze = None
        self.register_buffer("emb", torch.empty(self.k, self.d))
        nn.init.xavier_uniform_(self.emb, gain=10)
        if self.ema_gamma >= 1.0 or self.ema_gamma <= 0:
            raise RuntimeError("VQEMA must use an EMA-gamma value in (0, 1)")
        if self.training:
            self.min_dist = None
            self.circ_inds = None
            self.register_buffer("ind_hist", torch.zeros(self.k))
            self.register_buffer("ema_numer", torch.empty(self.k, self.d))
            self.register_buffer("ema_denom", torch.empty(self.k))
            self.register_buffer("z_sum", torch.empty(self.k, self.d))
            self.register_buffer("n_sum", torch.empty(self.k))
            self.register_buffer("n_sum_ones", torch.ones(self.k))
            self.ema_numer = self.emb * self.ema_gamma_comp
            self.ema_denom = self.n_sum_ones * self.ema_gamma_comp
        xavier_init(self.linear)
        self.defer_ema = False
        # True: diagnostics with data-dependent shapes (`uniq = min_ind.unique()`, vqema_bn.py:157) are replaced by
        # static-shape device tensors (`n_unique`), so that the step can be captured into a CUDA graph (aewn/train.py)
        self.static_diagnostics = False

    def forward(self, z):
        _require_cuda(z)
        ze = ops.conv1x1_f32(z, self.linear.weight)       # 1x1 conv n_in -> d in exact fp32: the code indices depend on it
        self.ze = ze
        train = self.training
        zq, min_dist, min_ind, ze_norm = _VQAssignFn.apply(
            ze, self.emb, METRIC_SCALED_L2, self.ind_hist if train else None, self.z_sum if train else None,
            self.n_sum if train else None, train)
        self.min_dist = min_dist
        self.min_ind = min_ind
        if train:
            if self.static_diagnostics:
                self.uniq = None
                self.n_unique = (self.n_sum > 0).sum()       # number of codes used by this batch, as a device scalar
            else:
                self.uniq = min_ind.unique(sorted=False)
            self.ze_norm = ze_norm
            self.emb_norm = (self.emb ** 2).sum(dim=1).sqrt()
            if not self.defer_ema:
                self.apply_ema(self.z_sum, self.n_sum)
        return zq      # value = gathered codes, gradient goes straight through to ze (ReplaceGrad semantics)

    def apply_ema(self, z_sum, n_sum):
        """vqema_bn.py:190-195."""
        vp = lambda t: C.c_void_p(t.data_ptr() if t is not None else None)
        numer, denom = self.ema_numer.contiguous(), self.ema_denom.contiguous()
        L.check(L.lib().aewn_ema_update(vp(numer), vp(denom), vp(z_sum.contiguous()), vp(n_sum.contiguous()),
                                        C.c_float(self.ema_gamma), vp(None), C.c_int(self.k), C.c_int(self.d),
                                        ops._stream()), "aewn_ema_update")
        self.ema_numer, self.ema_denom = numer, denom

    def update_codebook(self):
        """vqema_bn.py:216-222: emb = ema_numer / ema_denom[:, None]."""
        vp = lambda t: C.c_void_p(t.data_ptr() if t is not None else None)
        emb = torch.empty_like(self.emb)
        L.check(L.lib().aewn_ema_update(vp(self.ema_numer.contiguous()), vp(self.ema_denom.contiguous()), vp(None),
                                        vp(None), C.c_float(self.ema_gamma), vp(emb), C.c_int(self.k), C.c_int(self.d),
                                        ops._stream()), "aewn_ema_update")
        self.emb = emb
        self.emb.detach_()


class VQEMALoss(nn.Module):
    """vqema_bn.py:225-266 (total loss = commitment term only, :246)."""

    def __init__(self, bottleneck):
        super().__init__()
        self.bn = bottleneck
        self.logsoftmax = nn.LogSoftmax(1)

    def forward(self, quant_pred, target_wav):
        com_loss_embeds = self.bn.min_dist * self.bn.gamma
        log_pred = self.logsoftmax(quant_pred)
        log_pred_target = torch.gather(log_pred, 1, target_wav.long().unsqueeze(1))
        rec_loss_ts = -log_pred_target
        total_loss = com_loss_embeds.sum()
        h = self.bn.ind_hist
        n = h / h.sum()
        ent = -(n * torch.where(n == 0, torch.zeros_like(n), torch.log2(n))).sum()      # util.entropy, util.py:98-105
        peak, peak_idx = log_pred.max(dim=1)
        if self.bn.uniq is None:       # static-shape diagnostics (bn.static_diagnostics): device scalars instead of ints
            nunq = self.bn.n_unique
            pk_nuq = torch.zeros(log_pred.shape[1], device=log_pred.device).index_fill_(0, peak_idx.flatten(), 1.0).sum()
        else:
            nunq, pk_nuq = self.bn.uniq.nelement(), peak_idx.unique().nelement()
        # detached: metrics that keep the autograd graph alive pin its AccumulateGrad nodes across steps (aewn/train.py)
        self.metrics = {
            "rec": rec_loss_ts.mean().detach(), "com": com_loss_embeds.mean().detach(),
            "min_ze": self.bn.ze_norm.min(), "max_ze": self.bn.ze_norm.max(),
            "min_emb": self.bn.emb_norm.min(), "max_emb": self.bn.emb_norm.max(),
            "hst_ent": ent, "nunq": nunq,
            "pk_m": peak.to(torch.float).mean().detach(), "pk_nuq": pk_nuq,
            "pk_sd": peak.to(torch.float).std().detach(),
        }
        return total_loss
