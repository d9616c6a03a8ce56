"""Host-side orchestration of the libaewn.so kernels: descriptor builders, weight packing, and the autograd Functions
behind the drop-in modules.  PyTorch is used for device memory, streams and autograd plumbing only; every FLOP of the
hot path runs in the CUDA kernels behind include/aewn.h.

Time axis convention ("absolute time"): all activations of one decoder stack are (B, C, Tp) fp32 buffers
on ONE time axis tau in [0, T0); layer l (dilation d_l) produces valid values for tau >= lead_l = sum_{j<=l} d_j and
reads x[tau - d_l] and x[tau] (wavenet.py:100: Conv1d is a cross-correlation, tap 0 <-> x[t], tap 1 <-> x[t+d]).
"""
import ctypes as C

import torch

from . import _lib as L


# ------------------------------------------------------------------------------------------------- small helpers
def ceil_to(x, m):
    return (x + m - 1) // m * m


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def new_buf(B, Cc, Tp, device):
    """(B, C, Tp) fp32 workspace.  Zero-filled on creation: TMA reads the aligned-down margins of these tensors and
    multiplies them by exact zeros, so they must never hold NaN/Inf bit patterns."""
    assert Tp % 32 == 0
    return torch.zeros(B, Cc, Tp, device=device, dtype=torch.float32)


def act_of(t, t_extent=None, channels=None):
    """aewn_act describing a (B, C, Tp) contiguous buffer (or a channel-sliced view of one)."""
    assert t.dtype == torch.float32 and t.dim() == 3 and t.stride(2) == 1
    a = L.Act()
    a.ptr = t.data_ptr()
    a.t_extent = int(t.shape[2] if t_extent is None else t_extent)
    a.channels = int(t.shape[1] if channels is None else channels)
    a.batch = int(t.shape[0])
    a.row_pitch = int(t.stride(1))
    a.batch_stride = int(t.stride(0))
    return a


def ntile(w_row, n_valid, out, mode=L.EPI_LINEAR, flags=0, seg_mask=0xF, t_lo=0, t_hi=0, t_zero_lo=0, out2=None,
          out3=None, out_toff=0, dup_toff=0, dup_t_hi=0, add=None, add2=None, add_toff=0, add_t_lo=0, bias=None,
          zero_count=None, n=None):
    """Build an aewn_ntile.  `out`/`out2`/`out3`/`add`/`add2` are (B, C', Tp) views already sliced to the tile's first
    channel; strides are taken from `out` (and `add`)."""
    nt = L.NTile()
    nt.w_row = int(w_row)
    nt.n_valid = int(n_valid)
    nt.n = int(n if n is not None else max(16, ceil_to(n_valid, 16)))
    nt.mode, nt.flags, nt.seg_mask = int(mode), int(flags), int(seg_mask)
    nt.t_lo, nt.t_hi, nt.t_zero_lo = int(t_lo), int(t_hi), int(t_zero_lo)
    ref = out if out is not None else out3
    nt.out = out.data_ptr() if out is not None else None
    nt.out2 = out2.data_ptr() if out2 is not None else None
    nt.out3 = out3.data_ptr() if out3 is not None else None
    nt.out_bs, nt.out_cs = int(ref.stride(0)), int(ref.stride(1))
    nt.out_toff, nt.dup_toff, nt.dup_t_hi = int(out_toff), int(dup_toff), int(dup_t_hi)
    nt.zero_count = zero_count.data_ptr() if zero_count is not None else None
    if add is not None:
        nt.add = add.data_ptr()
        nt.add_bs, nt.add_cs = int(add.stride(0)), int(add.stride(1))
        nt.add2 = add2.data_ptr() if add2 is not None else None
    nt.add_toff, nt.add_t_lo = int(add_toff), int(add_t_lo)
    nt.bias = bias.data_ptr() if bias is not None else None
    return nt


class LaunchProfiler:
    """Optional per-launch CUDA-event timing (bench.py's roofline leg).  Events are recorded on the launching stream."""

    def __init__(self):
        self.records = []      # (tag, start_event, end_event)

    def times_ms(self):
        out = {}
        for tag, e0, e1 in self.records:
            out.setdefault(tag, []).append(e0.elapsed_time(e1))
        return out


_prof = None


def set_profiler(p):
    global _prof
    _prof = p


def _prof_begin(tag):
    if _prof is None or tag is None:
        return None
    e0 = torch.cuda.Event(enable_timing=True)
    e0.record()
    return e0


def _prof_end(tag, e0):
    if e0 is not None:
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        _prof.records.append((tag, e0, e1))


def tgemm(acts, segs, w, ntiles, batch, t_begin, t_end, err=None, tag=None):
    """One aewn_tgemm launch.  acts: list of L.Act; segs: list of (act_idx, shift, channels, w_koff); w: (rows, kpad)
    fp32 contiguous; ntiles: list of L.NTile (split into launches of <= MAX_NTILES)."""
    assert w.dtype == torch.float32 and w.is_contiguous() and w.dim() == 2
    for i in range(0, len(ntiles), L.MAX_NTILES):
        chunk = ntiles[i:i + L.MAX_NTILES]
        d = L.TGemmDesc()
        for j, a in enumerate(acts):
            d.acts[j] = a
        d.n_acts = len(acts)
        for j, (ai, sh, ch, ko) in enumerate(segs):
            d.segs[j] = L.Seg(int(ai), int(sh), int(ch), int(ko))
        d.n_segs = len(segs)
        d.w, d.w_rows, d.w_kpad = w.data_ptr(), int(w.shape[0]), int(w.shape[1])
        for j, nt in enumerate(chunk):
            d.ntiles[j] = nt
        d.n_ntiles = len(chunk)
        d.batch, d.t_begin, d.t_end = int(batch), int(t_begin), int(t_end)
        d.err = err.data_ptr() if err is not None else None
        e0 = _prof_begin(tag)
        L.check(L.lib().aewn_tgemm(C.byref(d), _stream()), "aewn_tgemm")
        _prof_end(tag, e0)


def wgrad(acts, items, batch, err=None, tag=None):
    """aewn_wgrad launches.  items: list of dicts(g_act, x_act, g_row, x_row, m_valid, n_valid, shift, t_lo, t_hi, out,
    out_off (elements), out_rs, out_cs)."""
    lib = L.lib()
    sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    for i in range(0, len(items), L.WGRAD_MAX_ITEMS):
        chunk = items[i:i + L.WGRAD_MAX_ITEMS]
        d = L.WGradDesc()
        for j, a in enumerate(acts):
            d.acts[j] = a
        d.n_acts = len(acts)
        target_units = max(1, (5 * sms // 2) // len(chunk))
        for j, it in enumerate(chunk):
            w = L.WGradItem()
            w.g_act, w.x_act, w.g_row, w.x_row = it["g_act"], it["x_act"], it["g_row"], it["x_row"]
            w.m_valid, w.n_valid = it["m_valid"], it["n_valid"]
            w.n = max(16, ceil_to(it["n_valid"], 16))
            w.shift, w.t_lo, w.t_hi = it.get("shift", 0), it["t_lo"], it["t_hi"]
            kblocks = batch * ((it["t_hi"] - it["t_lo"] + 31) // 32)
            w.n_split = max(1, min(target_units, kblocks // 48))
            w.out = it["out"].data_ptr() + 4 * int(it.get("out_off", 0))
            w.out_rs, w.out_cs = int(it["out_rs"]), int(it["out_cs"])
            d.items[j] = w
        d.n_items = len(chunk)
        d.batch = int(batch)
        d.err = err.data_ptr() if err is not None else None
        e0 = _prof_begin(tag)
        L.check(lib.aewn_wgrad(C.byref(d), _stream()), "aewn_wgrad")
        _prof_end(tag, e0)


def chunks(total, size=256):
    """[(offset, width)] covering `total` output channels in n-tiles of at most `size`."""
    return [(o, min(size, total - o)) for o in range(0, total, size)]


# ------------------------------------------------------------------------------------------------- weight packing
def _pad_k(m, k):
    return torch.nn.functional.pad(m, (0, k - m.shape[1]))


class LayerPack:
    """K-major, zero-padded operand matrices of one GRCC layer (rebuilt from the live parameters each step).

    w1  [256*J][2*KR + KC]  rows: per 128-channel block j, 128 filt rows then 128 gate rows (GATE_FWD pairs column c
                            with column 128+c);  columns: tap0 | tap1 | cond proj | bias (the bias rides on an
                            all-ones conditioning channel, so no epilogue bias add and d(bias) falls out of wgrad)
    w2  [R + S][KD]         rows: dil_res (absent in the final layer) then dil_skp
    w2t [D][KR + KS]        [Wr^T | Ws^T]            (g_z   = Wr^T g_sig + Ws^T g_skp)
    w1t [R + C][2*K2]       [tap0^T | tap1^T] over (g_f;g_g), cond rows only under the tap-1 (unshifted) block
    """

    def __init__(self, p, R, D, S, Cc, final_layer):
        dev = p["conv_signal.weight"].device
        KR, KC, KD, KS, K2 = ceil_to(R, 32), ceil_to(Cc + 1, 32), ceil_to(D, 32), ceil_to(S, 32), ceil_to(2 * D, 32)
        self.KR, self.KC, self.KD, self.KS, self.K2 = KR, KC, KD, KS, K2
        J = (D + 127) // 128
        self.J = J

        def full(wc, pj, bias):
            b = bias if bias is not None else torch.zeros(D, device=dev)
            return torch.cat([_pad_k(wc[:, :, 0], KR), _pad_k(wc[:, :, 1], KR),
                              _pad_k(torch.cat([pj[:, :, 0], b[:, None]], 1), KC)], 1)

        ff = full(p["conv_signal.weight"], p["proj_signal.weight"], p.get("conv_signal.bias"))
        gg = full(p["conv_gate.weight"], p["proj_gate.weight"], p.get("conv_gate.bias"))
        w1 = torch.zeros(256 * J, ff.shape[1], device=dev)
        for j in range(J):
            nj = min(128, D - 128 * j)
            w1[256 * j:256 * j + nj] = ff[128 * j:128 * j + nj]
            w1[256 * j + 128:256 * j + 128 + nj] = gg[128 * j:128 * j + nj]
        self.w1 = w1
        ws = p["dil_skp.weight"][:, :, 0]
        if final_layer:
            self.w2 = _pad_k(ws, KD).contiguous()
            self.w2t = torch.cat([torch.zeros(D, KR, device=dev), _pad_k(ws.t(), KS)], 1).contiguous()
        else:
            wr = p["dil_res.weight"][:, :, 0]
            self.w2 = _pad_k(torch.cat([wr, ws], 0), KD).contiguous()
            self.w2t = torch.cat([_pad_k(wr.t(), KR), _pad_k(ws.t(), KS)], 1).contiguous()
        wf, wg = p["conv_signal.weight"], p["conv_gate.weight"]
        tap = [_pad_k(torch.cat([wf[:, :, k].t(), wg[:, :, k].t()], 1), K2) for k in (0, 1)]        # (R, K2) each
        pc = _pad_k(torch.cat([p["proj_signal.weight"][:, :, 0].t(), p["proj_gate.weight"][:, :, 0].t()], 1), K2)
        self.w1t = torch.cat([torch.cat(tap, 1), torch.cat([torch.zeros(Cc, K2, device=dev), pc], 1)], 0).contiguous()


# ------------------------------------------------------------------------------------------------- stack geometry
class StackGeom:
    """Integer geometry of a GRCC stack on the absolute time axis."""

    def __init__(self, dils, T0, skip_start=None, last_is_final=True):
        self.dils = list(dils)
        self.L = len(self.dils)
        self.T0 = int(T0)
        self.Tp = ceil_to(self.T0 + 4, 32)
        self.lead = []
        acc = 0
        for d in self.dils:
            acc += d
            self.lead.append(acc)
        # RF = first absolute step that receives skip output (wavenet.py:62-66: skip_lead); in a full stack this is
        # the receptive field sum(d); a stand-alone layer may start its skip output later.
        self.RF = acc if skip_start is None else int(skip_start)
        self.last_is_final = bool(last_is_final)   # final layer has no dil_res / residual output (wavenet.py:36-37)
        self.W = self.T0 - self.RF
        if self.W <= 0 or self.RF < acc:
            raise ValueError("input shorter than the receptive field")

    def key(self):
        return (tuple(self.dils), self.T0, self.RF, self.last_is_final)

    def lead_in(self, l):
        return self.lead[l - 1] if l > 0 else 0


def needs_dup(d):
    """TMA box origins must be multiples of 4 elements: taps with d % 4 != 0 read a pre-shifted duplicate."""
    return d % 4 != 0


class StackWorkspace:
    """Persistent device buffers of one (B, widths, T0) configuration.  Reused across steps (never NaN: see new_buf)."""

    def __init__(self, B, R, D, S, Cc, geom, device):
        g = geom
        Tp = g.Tp
        # sig[l] = input of layer l; sig[L] = output of the last layer when it has a residual branch
        self.sig = [new_buf(B, R, Tp, device) for _ in range(g.L + (0 if g.last_is_final else 1))]
        self.generation = 0
        self.xs = {l: new_buf(B, R, Tp, device) for l in range(g.L) if needs_dup(g.dils[l])}
        self.th = [new_buf(B, D, Tp, device) for _ in range(g.L)]
        self.sg = [new_buf(B, D, Tp, device) for _ in range(g.L)]
        self.z = [new_buf(B, D, Tp, device) for _ in range(g.L)]
        self.skp = new_buf(B, S, Tp, device)
        self.cond = new_buf(B, Cc + 1, Tp, device)
        self.cond[:, Cc, :] = 1.0                                            # the bias channel
        self.err = torch.zeros(1, dtype=torch.int32, device=device)
        self._bwd = None
        self.B, self.R, self.D, self.S, self.Cc = B, R, D, S, Cc
        self.device = device

    def bwd(self):
        if self._bwd is None:
            B, R, D, S, Cc, dev = self.B, self.R, self.D, self.S, self.Cc, self.device
            Tp = self.sig[0].shape[2]
            self._bwd = dict(gfg=new_buf(B, 2 * D, Tp, dev), gfs=new_buf(B, 2 * D, Tp, dev),
                             gx=[new_buf(B, R, Tp, dev), new_buf(B, R, Tp, dev)], g_skp=new_buf(B, S, Tp, dev))
        return self._bwd


_workspaces = {}


def get_workspace(B, R, D, S, Cc, geom, device):
    key = (B, R, D, S, Cc, geom.key(), str(device))
    ws = _workspaces.get(key)
    if ws is None:
        if len(_workspaces) >= 2:   # bound the cache: configurations rarely alternate
            _workspaces.clear()
            torch.cuda.empty_cache()
        ws = _workspaces[key] = StackWorkspace(B, R, D, S, Cc, geom, device)
    return ws


def check_device_errors():
    """Synchronise and raise if any kernel reported a device-side fault (bounded-wait timeout, bad mu-law code)."""
    torch.cuda.synchronize()
    for ws in _workspaces.values():
        e = int(ws.err.item())
        if e != 0:
            ws.err.zero_()
            raise RuntimeError(f"aewn: device-side error word = {e} "
                               f"({'bounded wait timed out' if e == L.ERR_TIMEOUT else 'invalid input'})")


# ------------------------------------------------------------------------------------------------- stack forward
def stack_forward(ws, geom, packs, relu_last, save):
    """Run all GRCC layers.  Inputs already in the workspace: ws.sig[0] (and ws.xs[0]) = base-layer output, ws.cond.
    Result: ws.skp (B, S, Tp), valid on [RF, T0) -- with ReLU applied when relu_last (wavenet.py:359).
    save=False (inference) skips the tanh/sigmoid stores."""
    g = geom
    B, R, D, S, Cc = ws.B, ws.R, ws.D, ws.S, ws.Cc
    T0 = g.T0
    for l, d in enumerate(g.dils):
        pk = packs[l]
        final = (l == g.L - 1) and g.last_is_final
        lo = g.lead[l]
        lo4 = lo & ~3
        t_begin = lo4 & ~31
        x = ws.sig[l]
        xa, ca = act_of(x, T0), act_of(ws.cond, T0)
        if needs_dup(d):
            acts = [act_of(ws.xs[l], T0), xa, ca]
            segs = [(0, 0, R, 0), (1, 0, R, pk.KR), (2, 0, Cc + 1, 2 * pk.KR)]
        else:
            acts = [xa, ca]
            segs = [(0, -d, R, 0), (0, 0, R, pk.KR), (1, 0, Cc + 1, 2 * pk.KR)]
        tiles = []
        for j in range(pk.J):
            c0, nj = 128 * j, min(128, D - 128 * j)
            tiles.append(ntile(256 * j, nj, ws.th[l][:, c0:] if save else None, mode=L.EPI_GATE_FWD, n=256,
                               out2=ws.sg[l][:, c0:] if save else None, out3=ws.z[l][:, c0:],
                               t_lo=lo4, t_hi=T0, t_zero_lo=lo))
        tgemm(acts, segs, pk.w1, tiles, B, t_begin, T0, ws.err, tag=f"fwd_gemm1.{l}")

        tiles = []
        if not final:
            d_next = g.dils[l + 1] if l + 1 < g.L else 4
            for (c0, n) in chunks(R):
                tiles.append(ntile(c0, n, ws.sig[l + 1][:, c0:], add=x[:, c0:], t_lo=lo4, t_hi=T0, t_zero_lo=lo,
                                   out2=ws.xs[l + 1][:, c0:] if needs_dup(d_next) else None, dup_toff=d_next,
                                   dup_t_hi=T0))
        rf4 = g.RF & ~3
        flags = (L.F_ACCUM if l > 0 else 0) | (L.F_RELU if (l == g.L - 1 and relu_last) else 0)
        row0 = 0 if final else R
        for (c0, n) in chunks(S):
            tiles.append(ntile(row0 + c0, n, ws.skp[:, c0:], flags=flags, t_lo=rf4, t_hi=T0,
                               t_zero_lo=g.RF if l == 0 else 0))
        tgemm([act_of(ws.z[l], T0)], [(0, 0, D, 0)], pk.w2, tiles, B, t_begin, T0, ws.err, tag=f"fwd_gemm2.{l}")


# ------------------------------------------------------------------------------------------------- stack backward
def stack_backward(ws, geom, packs, params, g_skp, need_gx0=True, g_sig_last=None):
    """Backward of stack_forward.  g_skp: (B, S, Tp) gradient w.r.t. the (pre-ReLU) skip sum, zero
    below RF.  g_sig_last: (B, R, Tp) gradient w.r.t. the last layer's residual output (stand-alone layers only; zero
    on [lead & ~3, lead)).  Returns (g_x0 buffer (B,R,Tp) valid on [0,T0), g_cond (B,Cc,Tp), per-layer grad dicts)."""
    g = geom
    B, R, D, S, Cc = ws.B, ws.R, ws.D, ws.S, ws.Cc
    T0 = g.T0
    dev = ws.device
    bw = ws.bwd()
    gfg, gfs = bw["gfg"], bw["gfs"]
    g_cond = torch.zeros(B, Cc, g.Tp, device=dev)
    rf4 = g.RF & ~3
    grads = [None] * g.L
    g_sig = g_sig_last                   # gradient w.r.t. the output of the layer being processed
    for l in range(g.L - 1, -1, -1):
        d = g.dils[l]
        pk = packs[l]
        p = params[l]
        final = (l == g.L - 1) and g.last_is_final
        lo, lo_prev = g.lead[l], g.lead_in(l)
        lo4, lop4 = lo & ~3, lo_prev & ~3
        x = ws.sig[l]

        # (1) g_z = Wr^T g_sig + Ws^T g_skp, then the gate derivative -> gfg = [g_f ; g_g]
        if g_sig is not None:
            acts = [act_of(g_sig, T0), act_of(g_skp, T0)]
            segs = [(0, 0, R, 0), (1, 0, S, pk.KR)]
        else:
            acts = [act_of(g_skp, T0)]
            segs = [(0, 0, S, pk.KR)]
        t_store = min(lop4, lo4)
        tile = ntile(0, D, gfg, mode=L.EPI_GATE_BWD, out2=gfg[:, D:], add=ws.th[l], add2=ws.sg[l],
                     out3=gfs if needs_dup(d) else None, dup_toff=-d, dup_t_hi=T0,
                     t_lo=t_store, t_hi=T0, t_zero_lo=lo)
        tgemm(acts, segs, pk.w2t, [tile], B, t_store & ~31, T0, ws.err, tag=f"bwd_gz.{l}")

        # (2) g_x[tau] = tap1^T gfg[tau] + tap0^T gfg[tau + d] (+ g_sig[tau]);  g_cond[tau] += P^T gfg[tau]
        if needs_dup(d):
            acts = [act_of(gfs, T0 - d), act_of(gfg, T0)]
            segs = [(0, 0, 2 * D, 0), (1, 0, 2 * D, pk.K2)]
        else:
            acts = [act_of(gfg, T0)]
            segs = [(0, d, 2 * D, 0), (0, 0, 2 * D, pk.K2)]
        tiles = []
        gx = None
        if l > 0 or need_gx0:
            gx = bw["gx"][l % 2]
            for (c0, n) in chunks(R):
                tiles.append(ntile(c0, n, gx[:, c0:], add=g_sig[:, c0:] if g_sig is not None else None, add_t_lo=lo,
                                   t_lo=lop4, t_hi=T0, t_zero_lo=lo_prev))
        for (c0, n) in chunks(Cc):
            tiles.append(ntile(R + c0, n, g_cond[:, c0:], flags=L.F_ACCUM, seg_mask=2, t_lo=lo, t_hi=T0))
        tgemm(acts, segs, pk.w1t, tiles, B, lop4 & ~31, T0, ws.err, tag=f"bwd_dgrad.{l}")

        # (3) weight gradients of conv_signal/conv_gate/proj_signal/proj_gate (+ biases via the ones channel)
        gr = {}
        dwf, dwg = torch.zeros_like(p["conv_signal.weight"]), torch.zeros_like(p["conv_gate.weight"])
        dpb = torch.zeros(2 * D, Cc + 1, device=dev)
        x0_act = act_of(ws.xs[l], T0) if needs_dup(d) else act_of(x, T0)
        acts = [act_of(gfg, T0), x0_act, act_of(x, T0), act_of(ws.cond, T0)]
        sh0 = 0 if needs_dup(d) else -d
        items = []
        for h, dw in ((0, dwf), (1, dwg)):
            for i in range((D + 127) // 128):
                mv = min(128, D - 128 * i)
                base = dict(g_act=0, g_row=h * D + 128 * i, m_valid=mv, t_lo=lo4, t_hi=T0)
                for (c0, n) in chunks(R):
                    items.append(dict(base, x_act=1, x_row=c0, n_valid=n, shift=sh0, out=dw,
                                      out_off=(128 * i) * R * 2 + c0 * 2 + 0, out_rs=2 * R, out_cs=2))
                    items.append(dict(base, x_act=2, x_row=c0, n_valid=n, shift=0, out=dw,
                                      out_off=(128 * i) * R * 2 + c0 * 2 + 1, out_rs=2 * R, out_cs=2))
                for (c0, n) in chunks(Cc + 1):
                    items.append(dict(base, x_act=3, x_row=c0, n_valid=n, shift=0, out=dpb,
                                      out_off=(h * D + 128 * i) * (Cc + 1) + c0, out_rs=Cc + 1, out_cs=1))
        wgrad(acts, items, B, ws.err, tag=f"wgrad1.{l}")
        gr["conv_signal.weight"], gr["conv_gate.weight"] = dwf, dwg
        gr["proj_signal.weight"] = dpb[:D, :Cc].unsqueeze(2).contiguous()
        gr["proj_gate.weight"] = dpb[D:, :Cc].unsqueeze(2).contiguous()
        if "conv_signal.bias" in p:
            gr["conv_signal.bias"], gr["conv_gate.bias"] = dpb[:D, Cc].contiguous(), dpb[D:, Cc].contiguous()

        # (4) dWr = sum g_sig z^T  (tau >= lead_l);  dWs = sum g_skp z^T  (tau >= RF)
        dws = torch.zeros_like(p["dil_skp.weight"])
        acts = [act_of(g_skp, T0), act_of(ws.z[l], T0)]
        items = []
        for i in range((S + 127) // 128):
            for (c0, n) in chunks(D):
                items.append(dict(g_act=0, x_act=1, g_row=128 * i, x_row=c0, m_valid=min(128, S - 128 * i), n_valid=n,
                                  t_lo=rf4, t_hi=T0, out=dws, out_off=128 * i * D + c0, out_rs=D, out_cs=1))
        if not final and g_sig is not None:
            dwr = torch.zeros_like(p["dil_res.weight"])
            acts.append(act_of(g_sig, T0))
            for i in range((R + 127) // 128):
                for (c0, n) in chunks(D):
                    items.append(dict(g_act=2, x_act=1, g_row=128 * i, x_row=c0, m_valid=min(128, R - 128 * i),
                                      n_valid=n, t_lo=lo4, t_hi=T0, out=dwr, out_off=128 * i * D + c0, out_rs=D,
                                      out_cs=1))
            gr["dil_res.weight"] = dwr
        wgrad(acts, items, B, ws.err, tag=f"wgrad2.{l}")
        gr["dil_skp.weight"] = dws
        grads[l] = gr
        g_sig = gx
    return g_sig, g_cond, grads


# ------------------------------------------------------------------------------------------------- generic convs
def to_buf(x, pad_t=0):
    """Copy a (B, C, T) tensor into a fresh TMA-legal (B, C, Tp) buffer (Tp % 32 == 0, zero tail)."""
    B, Cc, T = x.shape
    buf = new_buf(B, Cc, ceil_to(T + pad_t + 4, 32), x.device)
    buf[:, :, :T] = x
    return buf


_ones_cache = {}


def ones_row(B, Tp, device):
    """(B, 1, Tp) all-ones activation: contracting a gradient against it yields the bias gradient inside wgrad."""
    key = (B, Tp, str(device))
    t = _ones_cache.get(key)
    if t is None:
        if len(_ones_cache) > 16:
            _ones_cache.clear()
        t = _ones_cache[key] = torch.ones(B, 1, Tp, device=device)
    return t


_err_cache = {}


def err_word(device):
    key = str(device)
    t = _err_cache.get(key)
    if t is None:
        t = _err_cache[key] = torch.zeros(1, dtype=torch.int32, device=device)
    return t


def pack_taps(weight, kc):
    """(N, C, k) conv weight -> K-major [N][k * kc] with tap j in columns [j*kc, j*kc + C)."""
    N, Cc, k = weight.shape
    return torch.cat([_pad_k(weight[:, :, j], kc) for j in range(k)], 1).contiguous()


class _TapConvFn(torch.autograd.Function):
    """y[b, n, u] = epi( sum_j sum_c W[n, c, j] * x[b, c, u*stride + j] + bias[n] ),  u in [0, T_out)
    with epi = identity | relu | relu-then-add-residual (wave_encoder.py:39-43).  k <= 4 taps.  Forward, data
    gradient and weight gradient all run on the tcgen05 engines; PyTorch only stages the tap-shifted copies that
    TMA's 16-byte origin rule requires (taps 1..3 of a stride-1 conv are not 16-byte aligned shifts)."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, mode, res_lw, zero_count):
        # mode: 0 linear, 1 relu, 2 relu-first + residual x[:, :, res_lw : res_lw + T_out]
        B, Cc, T = x.shape
        N, _, k = weight.shape
        assert k <= L.MAX_SEGS
        T_out = (T - k) // stride + 1
        dev = x.device
        err = err_word(dev)
        kc = ceil_to(Cc, 32)
        with torch.no_grad():
            xd = x.detach()
            taps = [to_buf(xd[:, :, j::stride][:, :, :T_out]) for j in range(k)]
            Tp = taps[0].shape[2]
            wp = pack_taps(weight.detach(), kc)
            out = new_buf(B, N, Tp, dev)
            relu_out = new_buf(B, N, Tp, dev) if mode == 2 else None
            flags = {0: 0, 1: L.F_RELU, 2: L.F_RELU_FIRST}[mode]
            tiles = []
            for (c0, n) in chunks(N):
                tiles.append(ntile(c0, n, out[:, c0:], flags=flags, t_lo=0, t_hi=T_out,
                                   bias=bias.detach()[c0:] if bias is not None else None,
                                   add=taps[res_lw][:, c0:] if mode == 2 else None,
                                   out3=relu_out[:, c0:] if mode == 2 else None, zero_count=zero_count))
            tgemm([act_of(t, T_out) for t in taps], [(j, 0, Cc, j * kc) for j in range(k)], wp, tiles, B, 0, T_out, err)
        ctx.save_for_backward(weight)
        ctx.taps, ctx.out, ctx.relu_out = taps, out, relu_out
        ctx.cfg = (B, Cc, T, N, k, T_out, stride, mode, res_lw, bias is not None)
        return out[:, :, :T_out].clone()

    @staticmethod
    def backward(ctx, g):
        (weight,) = ctx.saved_tensors
        B, Cc, T, N, k, T_out, stride, mode, res_lw, has_bias = ctx.cfg
        dev = g.device
        err = err_word(dev)
        lib = L.lib()
        Tp = ctx.out.shape[2]
        # g_pre = g * (activation > 0)
        if mode == 0:
            gp = to_buf(g)
        else:
            gp = new_buf(B, N, Tp, dev)
            mask = ctx.out if mode == 1 else ctx.relu_out
            gc = g if g.stride(2) == 1 else g.contiguous()
            L.check(lib.aewn_relu_mask_bwd(
                C.c_void_p(gc.data_ptr()), C.c_longlong(gc.stride(0)), C.c_longlong(gc.stride(1)),
                C.c_void_p(mask.data_ptr()), C.c_longlong(mask.stride(0)), C.c_longlong(mask.stride(1)),
                C.c_void_p(gp.data_ptr()), C.c_longlong(gp.stride(0)), C.c_longlong(gp.stride(1)),
                C.c_int(B), C.c_int(N), C.c_int(T_out), _stream()), "aewn_relu_mask_bwd")
        # weight / bias gradients:  dW[n, c, j] = sum_{b,u} gp[b, n, u] * tap_j[b, c, u]
        dw = torch.zeros_like(weight)
        db = torch.zeros(N, device=dev) if has_bias else None
        acts = [act_of(gp, T_out)] + [act_of(t, T_out) for t in ctx.taps]
        if has_bias:
            acts.append(act_of(ones_row(B, Tp, dev), T_out))
        items = []
        for i in range((N + 127) // 128):
            base = dict(g_act=0, g_row=128 * i, m_valid=min(128, N - 128 * i), t_lo=0, t_hi=T_out)
            for j in range(k):
                for (c0, n) in chunks(Cc):
                    items.append(dict(base, x_act=1 + j, x_row=c0, n_valid=n, out=dw,
                                      out_off=(128 * i) * Cc * k + c0 * k + j, out_rs=Cc * k, out_cs=k))
            if has_bias:
                items.append(dict(base, x_act=1 + k, x_row=0, n_valid=1, out=db, out_off=128 * i, out_rs=1, out_cs=1))
        wgrad(acts, items, B, err)
        # data gradient, one phase r of the input at a time:  g_x[s*u + r] = sum_m W[:, :, r + s*m]^T gp[u - m]
        g_x = torch.zeros(B, Cc, T, device=dev)
        kn = ceil_to(N, 32)
        for r in range(min(stride, k)):
            js = list(range(r, k, stride))
            n_u = (T - r + stride - 1) // stride
            shifted = []
            for m in range(len(js)):
                sb = new_buf(B, N, ceil_to(n_u + 4, 32), dev)
                hi = min(n_u, T_out + m)
                if hi > m:
                    sb[:, :, m:hi] = gp[:, :, :hi - m]
                shifted.append(sb)
            wt = torch.cat([_pad_k(weight.detach()[:, :, j].t(), kn) for j in js], 1).contiguous()
            gph = new_buf(B, Cc, ceil_to(n_u + 4, 32), dev)
            tiles = [ntile(c0, n, gph[:, c0:], t_lo=0, t_hi=n_u) for (c0, n) in chunks(Cc)]
            tgemm([act_of(sb, n_u) for sb in shifted], [(m, 0, N, m * kn) for m in range(len(js))], wt, tiles, B, 0, n_u,
                  err)
            g_x[:, :, r::stride] = gph[:, :, :n_u]
        if mode == 2:
            g_x[:, :, res_lw:res_lw + T_out] += g
        return g_x, dw, db, None, None, None, None


def tap_conv(x, weight, bias=None, stride=1, mode=0, res_lw=0, zero_count=None):
    return _TapConvFn.apply(x, weight, bias, stride, mode, res_lw, zero_count)
