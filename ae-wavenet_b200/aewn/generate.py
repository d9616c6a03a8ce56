"""Host side of the persistent incremental sampler (reference: wavenet.py:367-531, WaveNet.forward_test).

``GenPlan`` packs the decoder's weights once into the per-CTA row streams the kernel consumes (include/aewn.h,
"Incremental sampler"), owns the history rings / code buffers, and drives ``aewn_gen_run`` in bounded slices so a
single launch never runs for more than a fraction of a second.  All arithmetic happens in csrc/gen.cu.
"""
import ctypes as C

import torch

from . import _lib as L
from . import ops

SLICE_STEPS = 4096          # steps per launch
SMEM_BUDGET = 227 * 1024
GATE_CHUNK_ROWS = 8


def _r4(x):
    return (x + 3) & ~3


def _slice_map(n_total, cl, device):
    """Slice-padded layout (include/aewn.h): element i of a vector owned 1/cl per CTA -> padded position."""
    n = n_total // cl
    i = torch.arange(n_total, device=device)
    return (i // n) * _r4(n) + (i % n), cl * _r4(n)


def _scatter_cols(m, pos, width):
    """(rows, N) -> (rows, width) with column i moved to pos[i], zeros elsewhere."""
    out = m.new_zeros(m.shape[0], width)
    out[:, pos] = m
    return out


class GenPlan:
    """Everything that depends on the weights and the replica count, but not on the utterance."""

    def __init__(self, wn, n_rep, cluster=None):
        dev = wn.base_layer.weight.device
        if dev.type != "cuda":
            raise RuntimeError("aewn: WaveNet.forward_test needs the module on a CUDA device (no CPU fallback)")
        self.device = dev
        layers = list(wn.conv_layers)
        self.n_layers = len(layers)
        if self.n_layers > L.GEN_MAX_LAYERS:
            raise ValueError(f"aewn: forward_test supports at most {L.GEN_MAX_LAYERS} layers")
        self.R, self.D, self.S = wn.n_res, wn.n_dil, wn.n_skp
        self.P, self.Q, self.Cc = wn.post1.out_channels, wn.n_quant, wn.n_cond
        self.dils = [int(layer.dil) for layer in layers]
        self.cond_pitch = _r4(self.Cc + 1)
        self.n_rep_real = int(n_rep)
        per = 1 if n_rep == 1 else 2 if n_rep == 2 else 4
        self.n_rep = per
        self.n_groups = (int(n_rep) + per - 1) // per
        self.rows = self.n_groups * per
        d = L.GenDesc()
        d.n_layers, d.R, d.D, d.S, d.P, d.Q = self.n_layers, self.R, self.D, self.S, self.P, self.Q
        d.n_rep, d.n_groups = per, self.n_groups
        d.cond_pitch = self.cond_pitch
        off = 0
        for l, dl in enumerate(self.dils):
            d.dil[l] = dl
            d.hist_off[l] = off
            off += dl + 1
        d.hist_off[self.n_layers] = off
        self.hist_slots = off
        d.n_blocks = 2 * self.n_layers + 2
        self.err = ops.err_word(dev)
        d.err = self.err.data_ptr()
        self.desc = d
        self.cluster = self._choose_cluster(cluster)
        self._pack(wn)

    # ---------------------------------------------------------------------------------------------- geometry
    def _set_cluster(self, cl):
        d = self.desc
        d.cluster = cl
        self.Rp = cl * _r4(self.R // cl)
        self.Dp, self.Sp, self.Pp = cl * _r4(self.D // cl), cl * _r4(self.S // cl), cl * _r4(self.P // cl)
        self.KA = 2 * self.Rp + self.cond_pitch
        d.base_pitch = self.Rp
        stage = _r4(GATE_CHUNK_ROWS * self.KA) * 4
        stage = (stage + 127) & ~127
        d.stage_bytes = stage
        d.n_stages = 2
        fixed = L.lib().aewn_gen_smem_bytes(C.byref(d)) - 2 * stage
        n = min(24, (SMEM_BUDGET - fixed - 1024) // stage)
        if n < 2:
            return False
        d.n_stages = int(n)
        pairs, nres, nskp = self.D // cl, self.R // cl, self.S // cl
        off = 0
        for l in range(self.n_layers):
            final = (l == self.n_layers - 1)
            a, b = d.blocks[2 * l], d.blocks[2 * l + 1]
            a.kind, a.rows, a.rowf, a.off = 0, 2 * pairs, self.KA, off
            off += a.rows * a.rowf
            b.kind, b.rows, b.rowf, b.off = 1, (0 if final else nres) + nskp, self.Dp + 4, off
            off += b.rows * b.rowf
        p1, p2 = d.blocks[2 * self.n_layers], d.blocks[2 * self.n_layers + 1]
        p1.kind, p1.rows, p1.rowf, p1.off = 2, self.P // cl, self.Sp + 4, off
        off += p1.rows * p1.rowf
        p2.kind, p2.rows, p2.rowf, p2.off = 3, self.Q // cl, self.Pp + 4, off
        off += p2.rows * p2.rowf
        self.stream_len = _r4(off)
        d.stream_stride = self.stream_len
        return True

    def _choose_cluster(self, want):
        cands = [want] if want else [16, 8, 4, 2, 1]
        dims = (self.R, self.D, self.S, self.P, self.Q)
        # placeholders so that the occupancy query validates; real buffers are bound per utterance
        dummy = torch.zeros(64, device=self.device)
        d = self.desc
        for cl in cands:
            if any(x % cl for x in dims):
                continue
            if not self._set_cluster(cl):
                continue
            for name in ("wstream", "cond", "base_t", "hist", "wav", "uniforms"):
                setattr(d, name, dummy.data_ptr())
            d.wav_pitch, d.cond_len, d.t_begin, d.t_end, d.t_prime = 4, 2, 0, 1, 0
            n = C.c_int(0)
            L.check(L.lib().aewn_gen_max_clusters(C.byref(d), C.byref(n)), "aewn_gen_max_clusters")
            if n.value >= 1:
                self.max_clusters = n.value
                return cl
        raise RuntimeError(f"aewn: no launchable cluster size for forward_test (dims R,D,S,P,Q = {dims})")

    # ---------------------------------------------------------------------------------------------- weights
    def _pack(self, wn):
        """Per-CTA streams in consumption order; see the layout comment in include/aewn.h."""
        cl, R, D, S, Cc = self.cluster, self.R, self.D, self.S, self.Cc
        Rp, dev = self.Rp, self.device
        pairs, nres, nskp = D // cl, R // cl, S // cl
        f32 = dict(dtype=torch.float32, device=dev)
        parts = []   # each (cl, rows_per_cta * rowf)

        def bias_of(m, n):
            return m.bias.detach().float() if m.bias is not None else torch.zeros(n, **f32)

        xpos, _ = _slice_map(R, cl, dev)
        zpos, _ = _slice_map(D, cl, dev)
        spos, _ = _slice_map(S, cl, dev)
        ppos, _ = _slice_map(self.P, cl, dev)
        with torch.no_grad():
            for l, layer in enumerate(wn.conv_layers):
                rows = []
                for conv, proj in ((layer.conv_signal, layer.proj_signal), (layer.conv_gate, layer.proj_gate)):
                    w = conv.weight.detach().float()                      # (D, R, 2): tap 0 -> x[t-d], tap 1 -> x[t]
                    m = torch.zeros(D, self.KA, **f32)
                    m[:, xpos] = w[:, :, 0]
                    m[:, Rp + xpos] = w[:, :, 1]
                    m[:, 2 * Rp:2 * Rp + Cc] = proj.weight.detach().float()[:, :, 0]
                    m[:, 2 * Rp + Cc] = bias_of(conv, D)
                    rows.append(m)
                gate = torch.stack(rows, 1)                               # (D, 2, KA): [filt_j, gate_j]
                parts.append(gate.reshape(cl, pairs * 2 * self.KA))
                kb = self.Dp + 4
                mix = []
                if not layer.final_layer:
                    mix.append(_scatter_cols(layer.dil_res.weight.detach().float()[:, :, 0], zpos, kb)
                               .reshape(cl, nres * kb))
                mix.append(_scatter_cols(layer.dil_skp.weight.detach().float()[:, :, 0], zpos, kb)
                           .reshape(cl, nskp * kb))
                parts.append(torch.cat(mix, 1))
            for conv, pos, kp in ((wn.post1, spos, self.Sp), (wn.post2, ppos, self.Pp)):
                m = _scatter_cols(conv.weight.detach().float()[:, :, 0], pos, kp + 4)
                m[:, kp] = bias_of(conv, m.shape[0])
                parts.append(m.reshape(cl, -1))
            stream = torch.cat(parts, 1)
            assert stream.shape[1] <= self.stream_len, (stream.shape, self.stream_len)
            self.wstream = torch.zeros(cl, self.stream_len, **f32)
            self.wstream[:, :stream.shape[1]] = stream
            base = wn.base_layer.weight.detach().float()[:, :, 0].t()     # (Q, R)
            self.base_t = _scatter_cols(base + bias_of(wn.base_layer, R), xpos, Rp)
        self.desc.wstream = self.wstream.data_ptr()
        self.desc.base_t = self.base_t.data_ptr()

    # ---------------------------------------------------------------------------------------------- run
    def generate(self, codes, cond, t_prime, uniforms=None, want_logits=False, slice_steps=SLICE_STEPS):
        """codes (T,) int mu-law codes (the utterance, aligned with cond index 0); cond (C, n_ts) fp32.
        Returns (n_rep, T) int32 codes: [0, t_prime) copied, [t_prime, n_ts) drawn, the rest copied; optionally the
        logits (n_rep, T, Q) behind every draw."""
        dev, d = self.device, self.desc
        T = int(codes.shape[0])
        n_ts = int(cond.shape[1])
        if n_ts >= T + 1 or t_prime < 1 or t_prime > n_ts:
            raise ValueError(f"aewn: forward_test needs t_prime <= n_ts <= len(wav) (t_prime={t_prime}, n_ts={n_ts}, "
                             f"T={T})")
        wav = codes.to(device=dev, dtype=torch.int32).unsqueeze(0).repeat(self.rows, 1).contiguous()
        cond_t = torch.zeros(n_ts, self.cond_pitch, device=dev)
        cond_t[:, :self.Cc] = cond.t()
        cond_t[:, self.Cc] = 1.0
        if uniforms is None:
            uniforms = torch.rand(self.rows, T, device=dev)
        else:
            u = torch.zeros(self.rows, T, device=dev)
            u[:uniforms.shape[0]] = uniforms.to(dev)
            uniforms = u
        hist = torch.zeros(self.rows, self.hist_slots, self.Rp, device=dev)
        logits = torch.zeros(self.rows, T, self.Q, device=dev) if want_logits else None
        d.cond, d.cond_len = cond_t.data_ptr(), n_ts
        d.hist, d.wav, d.wav_pitch = hist.data_ptr(), wav.data_ptr(), T
        d.uniforms = uniforms.data_ptr()
        d.logits_out = logits.data_ptr() if want_logits else None
        d.t_prime = int(t_prime)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        t_last = n_ts - 1            # steps tau in [0, n_ts-1): the draw of step tau lands at tau+1 <= n_ts-1
        t = 0
        while t < t_last:
            d.t_begin, d.t_end = t, min(t_last, t + slice_steps)
            L.check(L.lib().aewn_gen_run(C.byref(d), stream), "aewn_gen_run")
            t = d.t_end
        ops.check_device_errors()
        self._keep = (cond_t, uniforms, hist)
        out = wav[:self.n_rep_real]
        return (out, logits[:self.n_rep_real]) if want_logits else out


_plans = {}


def get_plan(wn, n_rep):
    """Plans (descriptors, workspaces) are cached per module and parameter storage.  The packed weight stream is REBUILT
    from the live parameters on every call (a few hundred small copies, ~ms, against thousands of generated samples):
    an optimizer step or `p.data.copy_()` between two generations changes the weights without bumping `p._version`."""
    ptrs = tuple(p.data_ptr() for p in wn.parameters())
    key = (id(wn), int(n_rep))
    hit = _plans.get(key)
    if hit is not None and hit[0] == ptrs:
        hit[1]._pack(wn)
        return hit[1]
    if len(_plans) > 4:
        _plans.pop(next(iter(_plans)))
    plan = GenPlan(wn, n_rep)
    _plans[key] = (ptrs, plan)
    return plan
