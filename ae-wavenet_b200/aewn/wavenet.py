"""Drop-in replacements for the reference's wavenet.py classes, backed by libaewn.so (sm_100a).

Same class names, constructor signatures, sub-module names/creation order (=> identical RNG stream, parameters()
order and state_dict keys), buffers and caller-visible attributes as the reference:
  GatedResidualCondConv  wavenet.py:15-111      Conditioning  :114-140     Upsampling :142-165
  Conv1dWrap             wavenet.py:167-177     WaveNet       :179-364     RecLoss    :536-552
The nn.Conv1d / nn.Linear sub-modules are parameter containers only: the dilated gated-residual stack and the
base layer never call them -- they run in the CUDA kernels (ops.py).  There is no CPU fallback: calling forward on
CPU tensors raises.
"""
import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from . import _lib as L
from .compat import vconv, xavier_init


class _HP(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


LAYER_KEYS = ops.LAYER_KEYS


def _require_cuda(*tensors):
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError("aewn: the B200 hot path has no CPU implementation; move the module and its inputs to "
                               "a CUDA device (the oracle under oracle/ is test infrastructure, not a fallback)")


class GatedResidualCondConv(nn.Module):
    """wavenet.py:15-111.  forward(x (B,R,T), cond (B,C,Tc)) -> (sig (B,R,T-d), skp (B,S,T-d-skip_lead))."""

    def __init__(self, wavenet_vc, hps, n_cond, stride, dil, final_layer=False, parent_vc=None, name=None):
        super().__init__()
        self.wavenet_vc = wavenet_vc
        self.final_layer = final_layer
        self.dil = dil
        self.conv_signal = nn.Conv1d(hps.n_res, hps.n_dil, hps.filter_sz, dilation=dil, bias=hps.bias)
        self.conv_gate = nn.Conv1d(hps.n_res, hps.n_dil, hps.filter_sz, dilation=dil, bias=hps.bias)
        self.proj_signal = nn.Conv1d(n_cond, hps.n_dil, kernel_size=1, bias=False)
        self.proj_gate = nn.Conv1d(n_cond, hps.n_dil, kernel_size=1, bias=False)
        self.dil_skp = nn.Conv1d(hps.n_dil, hps.n_skp, kernel_size=1, bias=False)
        if not final_layer:
            self.dil_res = nn.Conv1d(hps.n_dil, hps.n_res, kernel_size=1, bias=False)
        if hps.filter_sz != 2:
            raise ValueError("aewn: the fused GRCC kernels implement filter_sz == 2 (the reference's only setting)")
        dil_filter_sz = (hps.filter_sz - 1) * dil + 1
        self.vc = vconv.VirtualConv(filter_info=(dil_filter_sz - 1, 0), parent=parent_vc, name=name)
        self.apply(xavier_init)

    # -- geometry bookkeeping, wavenet.py:45-89 ------------------------------------------------------------------
    def post_init(self):
        self.register_buffer("leads", torch.empty(4, dtype=torch.long))
        self.init_leads()
        self.set_full()

    def init_leads(self):
        cond_lead, r_off = vconv.output_offsets(self.wavenet_vc["beg_grcc"], self.vc)
        assert r_off == 0
        if self.vc == self.wavenet_vc["end_grcc"]:
            skip_lead = 0
        else:
            skip_lead, r_off = vconv.output_offsets(self.vc.child, self.wavenet_vc["end_grcc"])
            assert r_off == 0
        self.leads[0] = cond_lead
        self.leads[1] = skip_lead
        self.leads[2] = self.vc.l_wing_sz
        self.leads[3] = 0
        self._leads_host = [int(cond_lead), int(skip_lead), int(self.vc.l_wing_sz), 0]   # no D2H sync in forward
        self.global_rf = self.vc.in_len()
        self.local_rf = self.vc.filter_size()

    def set_incremental(self):
        self.cond, self.skip, self.lw = 3, 3, 2

    def set_full(self):
        self.cond, self.skip, self.lw = 0, 1, 2

    def param_dict(self):
        p = {}
        for k in LAYER_KEYS:
            mod, attr = k.split(".")
            m = getattr(self, mod, None)
            t = getattr(m, attr, None) if m is not None else None
            if t is not None:
                p[k] = t
        return p

    def forward(self, x, cond):
        _require_cuda(x, cond)
        lh = getattr(self, "_leads_host", None)
        if lh is None:
            lh = self._leads_host = [int(v) for v in self.leads.tolist()]
        cl, sl = lh[self.cond], lh[self.skip]
        p = self.param_dict()
        keys = list(p.keys())
        sig, skp = _LayerFn.apply(x, cond, self, cl, sl, keys, *[p[k] for k in keys])
        return sig, skp


class _LayerFn(torch.autograd.Function):
    """One stand-alone GRCC layer on the kernel path (used when a caller invokes a layer module directly)."""

    @staticmethod
    def forward(ctx, x, cond, mod, cl, sl, keys, *weights):
        d = mod.dil
        B, R, T_in = x.shape
        D, S, Cc = mod.conv_signal.out_channels, mod.dil_skp.out_channels, cond.shape[1]
        T_out = T_in - d
        if T_out <= sl or cond.shape[2] < cl + T_out:
            raise RuntimeError("aewn: GatedResidualCondConv input shorter than its receptive field / conditioning")
        geom = ops.StackGeom([d], T_in, skip_start=d + sl, last_is_final=mod.final_layer)
        p = dict(zip(keys, weights))
        plan = ops.get_plan(B, R, D, S, Cc, geom, [p], x.device, relu_last=False)
        plan.generation += 1
        with torch.no_grad():
            plan.sig[0][:, :, :T_in] = x
            if 0 in plan.xs:
                plan.xs[0][:, :, d:T_in] = x[:, :, :T_out]
            plan.cond[:, :Cc, d:T_in] = cond[:, :, cl:cl + T_out]
            plan.forward(save=True)
            skp = plan.skp[:, :, geom.RF:T_in].clone()
            sig = x[:, :, d:].clone() if mod.final_layer else plan.sig[1][:, :, d:T_in].clone()
        ctx.plan, ctx.keys = plan, keys
        ctx.gen = plan.generation
        ctx.dims = (B, R, D, S, Cc, T_in, d, cl, sl, cond.shape[2])
        ctx.final = mod.final_layer
        return sig, skp

    @staticmethod
    def backward(ctx, g_sig, g_skp):
        plan = ctx.plan
        geom = plan.geom
        if plan.generation != ctx.gen:
            raise RuntimeError("aewn: workspace was reused by a later forward before this backward ran")
        B, R, D, S, Cc, T_in, d, cl, sl, Tc = ctx.dims
        bw = plan.bwd()
        gs = bw["g_skp"]
        gs.zero_()
        gs[:, :, geom.RF:T_in] = g_skp
        if not ctx.final:
            gl = bw["g_last"]
            gl.zero_()
            gl[:, :, d:T_in] = g_sig
        gx, g_cond, grads = plan.backward()
        g_x = gx[:, :, :T_in].clone()
        if ctx.final:
            g_x[:, :, d:] += g_sig          # final layer: sig = x[:, :, lw:] (wavenet.py:105-106)
        g_c = torch.zeros(B, Cc, Tc, device=g_x.device)
        g_c[:, :, cl:cl + T_in - d] = g_cond[:, :, d:T_in]
        return (g_x, g_c, None, None, None, None) + tuple(grads[0][k].clone() for k in ctx.keys)


class Conditioning(nn.Module):
    """wavenet.py:114-140: concatenate up-sampled local conditioning with a learned speaker embedding."""

    def __init__(self, n_speakers, n_embed, bias=True):
        super().__init__()
        self.n_speakers = n_speakers
        self.speaker_embedding = nn.Linear(n_speakers, n_embed, bias)
        self.register_buffer("eye", torch.eye(n_speakers))
        self.apply(xavier_init)

    def forward(self, lc, speaker_inds):
        # Linear(one_hot(i)) == weight[:, i] + bias: a row gather instead of a one-hot matmul
        gc = self.speaker_embedding.weight.t()[speaker_inds.long()]
        if self.speaker_embedding.bias is not None:
            gc = gc + self.speaker_embedding.bias
        return torch.cat((lc, gc.unsqueeze(2).expand(-1, -1, lc.shape[2])), dim=1)


class Upsampling(nn.Module):
    """wavenet.py:142-165."""

    def __init__(self, n_chan, filter_sz, stride, parent_vc, bias=True, name=None):
        super().__init__()
        end_padding = stride - 1
        self.vc = vconv.VirtualConv(filter_info=filter_sz, stride=stride, padding=(end_padding, end_padding),
                                    is_downsample=False, parent=parent_vc, name=name)
        self.tconv = nn.ConvTranspose1d(n_chan, n_chan, filter_sz, stride, padding=filter_sz - stride, bias=bias)
        self.apply(xavier_init)
        self._stride, self._padding = stride, filter_sz - stride

    def forward(self, lc):
        _require_cuda(lc)
        if ops.FRONTEND == "kernels":   # polyphase transposed conv on the tcgen05 engine (opt-in, see ops.FRONTEND)
            return ops.tap_conv_transpose(lc, self.tconv.weight, self.tconv.bias, self._stride, self._padding)
        return self.tconv(lc)


class Conv1dWrap(nn.Conv1d):
    """wavenet.py:167-177."""

    def __init__(self, name, parent_vc, **kwargs):
        super().__init__(**kwargs)
        self.apply(xavier_init)
        self.vc = vconv.VirtualConv(filter_info=kwargs["kernel_size"], stride=kwargs["stride"], name=name,
                                    parent=parent_vc)

    def forward(self, x):
        _require_cuda(x)
        if ops.FRONTEND == "kernels" and self.dilation[0] == 1 and self.padding[0] == 0 and self.groups == 1:
            return ops.tap_conv(x, self.weight, self.bias, stride=self.stride[0], mode=0)
        return super().forward(x)


_OLD_API_KEYS = ("filter_sz", "n_lc_out", "lc_upsample_strides", "lc_upsample_filt_sizes", "n_res", "n_dil", "n_skp",
                 "n_post", "n_quant", "n_blocks", "n_block_layers", "n_global_embed", "n_speakers", "n_lc_in", "bias")


class WaveNet(nn.Module):
    """wavenet.py:179-364.  Accepts the current ctor ``WaveNet(hps, parent_vc=None)`` and the keyword form the stale
    ``AutoEncoder`` still uses (autoencoder_model.py:83-87: ``WaveNet(**dec_params, parent_vc=..., n_lc_in=...)``)."""
    __constants__ = ["conv_layers"]

    def __init__(self, hps=None, parent_vc=None, **old_api):
        super().__init__()
        if hps is None:
            unknown = set(old_api) - set(_OLD_API_KEYS)
            if unknown:
                raise TypeError(f"WaveNet() got unexpected keyword arguments {sorted(unknown)}")
            hps = _HP(old_api)
            hps.setdefault("bias", True)
        elif old_api:
            raise TypeError("WaveNet(): pass either hps or keyword hyper-parameters, not both")
        self.n_blocks = hps.n_blocks
        self.n_block_layers = hps.n_block_layers
        self.n_skp = hps.n_skp
        self.n_res = hps.n_res
        self.n_quant = hps.n_quant
        self.n_dil = hps.n_dil
        self.bias = hps.bias
        post_jitter_filt_sz = 3
        self.lc_conv = Conv1dWrap(f"LC_Conv(filter_size={post_jitter_filt_sz})", parent_vc, in_channels=hps.n_lc_in,
                                  out_channels=hps.n_lc_out, kernel_size=post_jitter_filt_sz, stride=1, bias=hps.bias)
        self.vc = dict()
        self.vc["beg"] = self.lc_conv.vc
        cur_vc = self.vc["beg"]
        self.lc_upsample = nn.Sequential()
        for i, (filt_sz, stride) in enumerate(zip(hps.lc_upsample_filt_sizes, hps.lc_upsample_strides)):
            mod = Upsampling(hps.n_lc_out, filt_sz, stride, cur_vc,
                             name=f"Upsampling_{i}(filter_sz={filt_sz}, stride={stride})")
            self.lc_upsample.add_module(str(i), mod)
            cur_vc = mod.vc
        self.vc["last_upsample"] = cur_vc
        self.cond = Conditioning(hps.n_speakers, hps.n_global_embed)
        self.base_layer = Conv1dWrap("Base Layer", cur_vc, in_channels=hps.n_quant, out_channels=hps.n_res,
                                     kernel_size=1, stride=1, dilation=1, bias=self.bias)
        self.base_layer.vc.do_trim_input = True
        cur_vc = self.base_layer.vc
        self.conv_layers = nn.ModuleList()
        n_cond = hps.n_lc_out + hps.n_global_embed
        self.n_cond = n_cond
        for b in range(self.n_blocks):
            for bl in range(self.n_block_layers):
                dil = 2 ** bl
                final_layer = (b + 1 == self.n_blocks and bl + 1 == self.n_block_layers)
                grc = GatedResidualCondConv(self.vc, hps, n_cond=n_cond, stride=1, dil=dil, final_layer=final_layer,
                                            parent_vc=cur_vc, name=f"GRCC_{b},{bl}(dil={dil})")
                self.conv_layers.append(grc)
                cur_vc = grc.vc
        self.vc["beg_grcc"] = self.conv_layers[0].vc
        self.vc["end_grcc"] = self.conv_layers[-1].vc
        self.relu = nn.ReLU()
        self.post1 = Conv1dWrap("Post1", cur_vc, in_channels=hps.n_skp, out_channels=hps.n_post, kernel_size=1,
                                stride=1, bias=hps.bias)
        self.post2 = Conv1dWrap("Post2", self.post1.vc, in_channels=hps.n_post, out_channels=hps.n_quant,
                                kernel_size=1, stride=1, bias=hps.bias)
        self.logsoftmax = nn.LogSoftmax(1)
        self.vc["main"] = self.post2.vc
        self.n_replicas = 1

    # -- geometry, wavenet.py:261-311 ------------------------------------------------------------------------------
    def set_parent_vc(self, parent_vc):
        self.vc["beg"].parent = parent_vc
        parent_vc.child = self.vc["beg"]

    def post_init(self, n_win_batch=None):
        if n_win_batch is None:     # stale caller (autoencoder_model.py:89): geometry is finalised later
            n_win_batch = getattr(self, "n_win_batch", None)
            if n_win_batch is None:
                raise TypeError("WaveNet.post_init() needs n_win_batch (wavenet.py:266)")
        one_gr = vconv.GridRange((0, int(1e12)), (0, 1), 1)
        win_gr = vconv.GridRange((0, int(1e12)), (0, n_win_batch), 1)
        vconv.compute_inputs(self.vc["end_grcc"], win_gr)
        di = self.vc["beg_grcc"].input_gr
        wi = self.vc["beg"].parent.input_gr
        self.wav_cond_offset = [int(di.sub[0] - wi.sub[0]), int(di.sub[1] - wi.sub[0])]
        vconv.compute_inputs(self.vc["end_grcc"], one_gr)
        for layer in self.conv_layers:
            layer.post_init()
        self.base_global_rf = self.conv_layers[0].global_rf
        self.n_win_batch = n_win_batch

    def get_input_size(self, output_size):
        win_gr = vconv.GridRange((0, int(1e12)), (0, output_size), 1)
        vconv.compute_inputs(self.vc["end_grcc"], win_gr)
        return self.vc["beg"].parent.in_len()

    def set_n_replicas(self, n_replicas):
        self.n_replicas = n_replicas

    def set_incremental(self):
        for layer in self.conv_layers:
            layer.set_incremental()

    def set_full(self):
        for layer in self.conv_layers:
            layer.set_full()

    # -- forward ---------------------------------------------------------------------------------------------------
    def forward(self, wav, lc_sparse, speaker_inds, jitter_index):
        if self.training:
            return self.forward_train(wav, lc_sparse, speaker_inds, jitter_index)
        return self.forward_test(wav, lc_sparse, speaker_inds, jitter_index)

    def conditioning(self, lc_sparse, speaker_inds, jitter_index, trim=True):
        """wavenet.py:330-343.  The jitter gather reproduces the reference exactly, including its quirk (SURVEY.md
        F8): torch.take on the flat tensor with an index that carries only the batch offset.  ``trim=False`` is the
        inference variant (wavenet.py:385-391), which conditions on the whole upsampled sequence."""
        B, D1, T = lc_sparse.shape
        flat_idx = jitter_index + (torch.arange(B, device=jitter_index.device) * jitter_index.shape[1]).unsqueeze(1)
        lc_jitter = lc_sparse.reshape(-1)[flat_idx].unsqueeze(1).expand(-1, D1, -1)
        lc_dense = self.lc_upsample(self.lc_conv(lc_jitter))
        if trim:
            t0, t1 = int(self.trim_ups_out[0]), int(self.trim_ups_out[1])
            lc_dense = lc_dense[:, :, t0:t1]
        return self.cond(lc_dense, speaker_inds)

    def stack_geometry(self, T0):
        dils = [layer.dil for layer in self.conv_layers]
        return ops.StackGeom(dils, T0)

    def forward_train(self, wav, lc_sparse, speaker_inds, jitter_index):
        """wavenet.py:323-364."""
        _require_cuda(wav, lc_sparse)
        if isinstance(self.trim_ups_out, torch.Tensor) and self.trim_ups_out.is_cuda:
            self.trim_ups_out = self.trim_ups_out.cpu()      # read on the host once, never per step
        cond = self.conditioning(lc_sparse, speaker_inds, jitter_index)
        keys, weights = [], []
        for li, layer in enumerate(self.conv_layers):
            for k, t in layer.param_dict().items():
                keys.append((li, k))
                weights.append(t)
        none = wav.new_zeros(0)
        opt = lambda t: t if t is not None else none
        # base layer + 20 GRCC layers + ReLU + post1 + ReLU + post2, all on the kernel path
        return _DecoderCoreFn.apply(wav, cond, self, keys, torch.is_grad_enabled(), self.base_layer.weight,
                                    opt(self.base_layer.bias),
                                    self.post1.weight, opt(self.post1.bias), self.post2.weight, opt(self.post2.bias),
                                    *weights)

    def forward_test(self, wav, lc_sparse, speaker_inds, jitter_index):
        """wavenet.py:367-531: draw ``n_replicas`` continuations of one utterance, sample by sample.

        Returns (n_replicas + 1, T_wav - wav_cond_offset[0]) codes in ``wav``'s dtype: row 0 is the input, rows 1..
        copy its first ``base_global_rf`` samples, are generated up to the end of the conditioning sequence and copy
        the input after that -- exactly what the reference returns.  The whole loop is ONE persistent kernel per
        ``generate.SLICE_STEPS`` samples (csrc/gen.cu); each draw is an inverse-CDF draw from the softmax, i.e. the same
        distribution as the reference's ``torch.multinomial``.  ``self.gen_uniforms`` (n_replicas, T) may pin the
        uniforms (tests); ``self.gen_logits`` receives the logits behind every draw when ``self.keep_gen_logits``."""
        from . import generate
        _require_cuda(wav, lc_sparse)
        if wav.shape[0] != 1:
            raise ValueError("aewn: forward_test generates for one utterance per call (the reference's buffers only "
                             "line up for batch 1, wavenet.py:397,419,463)")
        off0 = int(self.wav_cond_offset[0])
        codes = wav[0, off0:].long()
        with torch.no_grad():
            cond = self.conditioning(lc_sparse, speaker_inds, jitter_index, trim=False)[0]
            plan = generate.get_plan(self, int(self.n_replicas))
            keep = bool(getattr(self, "keep_gen_logits", False))
            res = plan.generate(codes, cond, int(self.base_global_rf), uniforms=getattr(self, "gen_uniforms", None),
                                want_logits=keep)
            if keep:
                res, self.gen_logits = res
        return torch.cat([codes.unsqueeze(0).to(wav.dtype), res.to(wav.dtype)], 0)


class _DecoderCoreFn(torch.autograd.Function):
    """Base layer (embedding gather) + the whole GRCC stack + ReLU of the skip sum, on the kernel path.
    Inputs: wav (B, T_wav) float mu-law codes, cond (B, C, T0).  Output: relu(skp_sum) (B, S, W)."""

    @staticmethod
    def forward(ctx, wav, cond, net, keys, want_grad, base_w, base_b, p1w, p1b, p2w, p2b, *weights):
        o0, o1 = net.wav_cond_offset
        T0 = o1 - o0
        B, Cc = cond.shape[0], cond.shape[1]
        if cond.shape[2] != T0:
            raise RuntimeError(f"aewn: conditioning length {cond.shape[2]} != decoder input length {T0} "
                               f"(wav_cond_offset={net.wav_cond_offset}); see SURVEY.md 9.5")
        if wav.shape[1] < o1:
            raise RuntimeError(f"aewn: wav has {wav.shape[1]} samples, wav_cond_offset needs {o1}")
        R, D, S, Q = net.n_res, net.n_dil, net.n_skp, net.n_quant
        geom = net.stack_geometry(T0)
        if geom.W != net.n_win_batch:
            raise RuntimeError(f"aewn: geometry mismatch, window {geom.W} != n_win_batch {net.n_win_batch}")
        dev = wav.device
        params = [dict() for _ in range(geom.L)]
        for (li, k), w in zip(keys, weights):
            params[li][k] = w
        plan = ops.get_plan(B, R, D, S, Cc, geom, params, dev, relu_last=True)
        plan.generation += 1
        wav_c = wav.detach().float().contiguous()
        with torch.no_grad():
            plan.cond[:, :Cc, :T0] = cond
            d0 = geom.dils[0]
            dup = plan.xs.get(0)          # pre-shifted copy for a TF32 weight-gradient tap (absent with the fp16 engines)
            bw = base_w.detach().reshape(R, Q)
            L.check(L.lib().aewn_base_embed_fwd(
                L.C.c_void_p(wav_c.data_ptr()), L.C.c_longlong(wav_c.stride(0)), L.C.c_int(o0),
                L.C.c_void_p(bw.data_ptr()), L.C.c_void_p(base_b.data_ptr() if base_b.numel() else None),
                L.C.c_void_p(plan.sig[0].data_ptr()), L.C.c_longlong(plan.sig[0].stride(0)),
                L.C.c_longlong(plan.sig[0].stride(1)), L.C.c_void_p(dup.data_ptr() if dup is not None else None),
                L.C.c_int(d0), L.C.c_int(T0), L.C.c_int(B), L.C.c_int(R), L.C.c_int(Q), L.C.c_int(T0),
                L.C.c_void_p(plan.err.data_ptr()), ops._stream()), "aewn_base_embed_fwd")
            # under torch.no_grad() (or with nothing upstream requiring a gradient) nothing is kept for a backward pass:
            # the layers run in inference mode (no tanh / sigmoid / z writes: SURVEY.md 8d's inference byte count)
            save = bool(want_grad) and any(ctx.needs_input_grad)     # (needs_input_grad ignores the grad mode)
            plan.forward(save=save)
            pw = {"post1.weight": p1w, "post1.bias": p1b if p1b.numel() else None, "post2.weight": p2w,
                  "post2.bias": p2b if p2b.numel() else None}
            post = getattr(plan, "post", None)
            if post is None or not post.matches(pw):
                post = plan.post = ops.PostPlan(plan, pw)
            logits = post.forward()
            # a device-side fault (bounded wait expired, mu-law code outside [0, Q), activation outside the fp16 operand
            # range) must not go unnoticed until somebody calls ops.check_device_errors(): poison ONE logit, so that the
            # loss of this step is NaN (two 1-element kernels, no synchronisation, capturable)
            logits[0, 0, geom.RF:geom.RF + 1] += torch.where(plan.err != 0, plan.nan_const, plan.zero_const)
            out = logits[:, :, geom.RF:T0]               # (B, Q, W) view of a fresh (B, Q, Tp) buffer
        ctx.saved_for_bwd = save
        ctx.post = post
        ctx.post_has_bias = (p1b.numel() > 0, p2b.numel() > 0)
        ctx.plan, ctx.keys = plan, keys
        ctx.gen = plan.generation
        ctx.wav, ctx.o0 = wav_c, o0
        ctx.dims = (B, R, D, S, Cc, Q, T0)
        ctx.has_bias = base_b.numel() > 0
        ctx.base_shape = base_w.shape
        ctx.leaves = (p1w, p1b, p2w, p2b) + tuple(weights)      # the parameters themselves (for .grad, see backward)
        return out

    @staticmethod
    def backward(ctx, g_out):
        plan = ctx.plan
        geom = plan.geom
        if not ctx.saved_for_bwd:
            raise RuntimeError("aewn: this forward ran in inference mode (torch.no_grad()): nothing was saved for backward")
        if plan.generation != ctx.gen:
            raise RuntimeError("aewn: workspace was reused by a later forward before this backward ran "
                               "(one in-flight forward per model and configuration)")
        B, R, D, S, Cc, Q, T0 = ctx.dims
        lib = L.lib()
        # post-net backward: fills the stack's g_skp buffer (gradient w.r.t. the pre-ReLU skip sum, absolute time axis)
        pviews = ctx.post.backward(g_out)
        gx0, g_cond, grads = plan.backward()
        d_base = torch.zeros(R, Q, device=g_out.device)
        d_bias = torch.zeros(R, device=g_out.device) if ctx.has_bias else None
        L.check(lib.aewn_base_embed_bwd(
            L.C.c_void_p(gx0.data_ptr()), L.C.c_longlong(gx0.stride(0)), L.C.c_longlong(gx0.stride(1)),
            L.C.c_void_p(ctx.wav.data_ptr()), L.C.c_longlong(ctx.wav.stride(0)), L.C.c_int(ctx.o0),
            L.C.c_void_p(d_base.data_ptr()), L.C.c_void_p(d_bias.data_ptr() if d_bias is not None else None),
            L.C.c_int(B), L.C.c_int(R), L.C.c_int(Q), L.C.c_int(T0), ops._stream()), "aewn_base_embed_bwd")
        # gradient views alias the plan's flat buffer (overwritten by the next backward): autograd accumulates them
        # into .grad right away; a caller that keeps them (torch.autograd.grad) gets private copies
        b1, b2 = ctx.post_has_bias
        views = [pviews["post1.weight"], pviews["post1.bias"] if b1 else None, pviews["post2.weight"],
                 pviews["post2.bias"] if b2 else None] + [grads[li][k] for (li, k) in ctx.keys]
        live = [leaf for v, leaf in zip(views, ctx.leaves) if v is not None]
        if ops.fused_accumulate_applies(live):
            # training-engine mode (FlatGradSync(fused_accumulate=True) owns these parameters and THIS backward call
            # accumulates into them): ONE launch adds every weight gradient into the existing .grad buffers
            pairs = [(v, leaf.grad) for v, leaf in zip(views, ctx.leaves) if v is not None]
            if ops.add_into_grads(pairs):
                return (None, g_cond[:, :, :T0].clone(), None, None, None, d_base.reshape(ctx.base_shape), d_bias) + \
                       (None,) * len(views)
        out = tuple(v.clone() if v is not None else None for v in views)
        return (None, g_cond[:, :, :T0].clone(), None, None, None, d_base.reshape(ctx.base_shape), d_bias) + out


class _NLLFn(torch.autograd.Function):
    """-mean log_softmax(pred)[target] with the log-softmax, gather, mean and their backward fused into two kernels."""

    @staticmethod
    def forward(ctx, pred, target):
        B, Q, N = pred.shape
        p = pred.detach()
        if p.stride(2) != 1:
            p = p.contiguous()
        tg = target.detach().float()
        if tg.stride(1) != 1:
            tg = tg.contiguous()
        lse = torch.empty(B, N, device=p.device)
        loss_sum = torch.empty(1, device=p.device)
        err = ops.err_word(p.device)
        L.check(L.lib().aewn_nll_fwd(
            L.C.c_void_p(p.data_ptr()), L.C.c_longlong(p.stride(0)), L.C.c_longlong(p.stride(1)),
            L.C.c_void_p(tg.data_ptr()), L.C.c_longlong(tg.stride(0)), L.C.c_void_p(lse.data_ptr()),
            L.C.c_void_p(loss_sum.data_ptr()), L.C.c_int(B), L.C.c_int(Q), L.C.c_int(N), L.C.c_void_p(err.data_ptr()),
            ops._stream()), "aewn_nll_fwd")
        ctx.save_for_backward(p, tg, lse)
        return (loss_sum / float(B * N)).reshape(())

    @staticmethod
    def backward(ctx, g):
        p, tg, lse = ctx.saved_tensors
        B, Q, N = p.shape
        gx = torch.empty(B, Q, N, device=p.device)
        gl = g.detach().float().reshape(1).contiguous()
        L.check(L.lib().aewn_nll_bwd(
            L.C.c_void_p(p.data_ptr()), L.C.c_longlong(p.stride(0)), L.C.c_longlong(p.stride(1)),
            L.C.c_void_p(tg.data_ptr()), L.C.c_longlong(tg.stride(0)), L.C.c_void_p(lse.data_ptr()),
            L.C.c_void_p(gl.data_ptr()), L.C.c_float(1.0 / float(B * N)), L.C.c_void_p(gx.data_ptr()),
            L.C.c_longlong(gx.stride(0)), L.C.c_longlong(gx.stride(1)), L.C.c_int(B), L.C.c_int(Q), L.C.c_int(N),
            ops._stream()), "aewn_nll_bwd")
        return gx, None


class RecLoss(nn.Module):
    """wavenet.py:536-552.  On CUDA the log-softmax + gather + mean run as one fused kernel pair."""

    def __init__(self):
        super().__init__()
        self.logsoftmax = nn.LogSoftmax(1)

    def forward(self, quant_pred, target_wav):
        _require_cuda(quant_pred)
        rec_loss = _NLLFn.apply(quant_pred, target_wav)
        # detached: a metrics entry that keeps the autograd graph alive also keeps every AccumulateGrad node (and the
        # stream it was created on) alive across steps, which breaks CUDA-graph capture of the step (aewn/train.py)
        self.metrics = {"rec": rec_loss.detach()}
        return rec_loss
