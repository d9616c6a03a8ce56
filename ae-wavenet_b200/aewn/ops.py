"""Host-side orchestration of the libaewn.so kernels: descriptor builders, weight packing, and the autograd Functions
behind the drop-in modules.  PyTorch is used for device memory, streams and autograd plumbing only; every FLOP of the
hot path runs in the CUDA kernels behind include/aewn.h.

Time axis convention ("absolute time"): all activations of one decoder stack are (B, C, Tp) fp32 buffers
on ONE time axis tau in [0, T0); layer l (dilation d_l) produces valid values for tau >= lead_l = sum_{j<=l} d_j and
reads x[tau - d_l] and x[tau] (wavenet.py:100: Conv1d is a cross-correlation, tap 0 <-> x[t], tap 1 <-> x[t+d]).
"""
import collections
import ctypes as C
import os

import torch

from . import _lib as L


# Engine mode of the plans built from here on (tests and bench.py A/B them; prebuilt plans keep theirs):
#   "pair"  2-CTA clusters issuing ONE cta_group::2 MMA stream (M = 256), operands split between the two CTAs
#   "mcast" 2-CTA clusters sharing W / X by TMA multicast, every CTA issuing its own cta_group::1 MMAs (M = 128)
#   "auto"  what measured fastest on B200 at the cfg2 shapes (profiles/r2_engine_ab.json): "pair" for every tgemm / wgrad
#           launch, and the stack's weight gradients through the wide-unit kernel aewn_wgradw (WIDE_WGRAD)
ENGINE_MODE = os.environ.get("AEWN_ENGINE_MODE", "auto")
MERGE_DGRAD_TILES = os.environ.get("AEWN_MERGE_DGRAD", "1") == "1"
WIDE_WGRAD = os.environ.get("AEWN_WIDE_WGRAD", "1") == "1"     # "auto" only: stack weight gradients through aewn_wgradw
# Forward of a dilation layer as ONE fused launch (aewn_grcc_fwd: fp16 tensor-core operands, fp32 accumulation and
# residual stream) where the widths allow it; "0" keeps the two-launch TF32 path (aewn_tgemm) everywhere.
FUSED_FWD = os.environ.get("AEWN_FUSED_FWD", "1") == "1"


def set_fused_forward(on):
    global FUSED_FWD
    FUSED_FWD = bool(on)
    _plans.clear()


def fused_ok(R, D, S, Cc):
    """Shapes aewn_grcc_fwd accepts (include/aewn.h); everything else runs through aewn_tgemm."""
    return D in (128, 256) and R % 8 == 0 and 8 <= R <= 1024 and S % 32 == 0 and 32 <= S <= 1024 and 1 <= Cc < 1024


def set_engine_mode(mode):
    global ENGINE_MODE
    if mode not in ("pair", "mcast", "auto"):
        raise ValueError("engine mode must be 'pair', 'mcast' or 'auto'")
    ENGINE_MODE = mode
    _plans.clear()


def _use_pair(tag):
    return ENGINE_MODE in ("auto", "pair")


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


def ntile(w_row, n_valid, out, mode=L.EPI_LINEAR, flags=0, seg_mask=0x3F, t_lo=0, t_hi=0, t_zero_lo=0, out2=None,
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


def build_tgemm(acts, segs, w, ntiles, batch, t_begin, t_end, err=None, tag=None):
    """Descriptors of one logical tgemm (split into launches of <= MAX_NTILES n-tiles).  acts: list of L.Act; segs: list
    of (act_idx, shift, channels, w_koff); w: (rows, kpad) fp32 contiguous.  Returns [("tgemm", desc, tag, keepalive)]."""
    assert w.dtype == torch.float32 and w.is_contiguous() and w.dim() == 2
    out = []
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
        if chunk[-1].flags & L.F_MERGE_NEXT:             # a merged pair never straddles two launches
            chunk[-1].flags &= ~L.F_MERGE_NEXT
        for j, nt in enumerate(chunk):
            d.ntiles[j] = nt
        d.n_ntiles = len(chunk)
        d.batch, d.t_begin, d.t_end = int(batch), int(t_begin), int(t_end)
        d.err = err.data_ptr() if err is not None else None
        d.cluster = L.CLUSTER_PAIR_MMA if _use_pair(tag) else 2
        out.append(("tgemm", d, tag))
    return out


def pair_items(items_by_x):
    """items_by_x: list of lists; each inner list holds the m-tile items that read the SAME X tile.  Returns a flat list in
    which items (2i, 2i+1) share their X tile (wgrad pair_x mode: 2-CTA cluster, X multicast); odd groups are padded
    with a partner that stores nothing (m_valid = 0)."""
    out = []
    for grp in items_by_x:
        grp = list(grp)
        if len(grp) % 2:
            grp.append(dict(grp[-1], m_valid=0))
        out += grp
    return out


def build_wgrad(acts, items, batch, err=None, tag=None, pair=False):
    """Descriptors of one logical wgrad.  items: list of dicts(g_act, x_act, g_row, x_row, m_valid, n_valid, shift, t_lo,
    t_hi, out, out_off (elements), out_rs, out_cs).  pair=True: items come in X-sharing pairs (see pair_items)."""
    sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    out = []
    for i in range(0, len(items), L.WGRAD_MAX_ITEMS):
        chunk = items[i:i + L.WGRAD_MAX_ITEMS]
        d = L.WGradDesc()
        for j, a in enumerate(acts):
            d.acts[j] = a
        d.n_acts = len(acts)
        n_work = len(chunk) // 2 if pair else len(chunk)          # clusters (pairs) or CTAs per split
        n_slots = sms // 2 if pair else sms                        # clusters / CTAs resident at once
        min_kb = min(batch * ((it["t_hi"] - it["t_lo"] + 31) // 32) for it in chunk)
        # split-K factor: fill whole waves of the persistent grid (units = n_work * n_split), keep >= 48 K blocks per
        # unit so the fp32 atomics of the epilogue stay a small fraction, prefer ~2 waves (epilogue/MMA overlap)
        best, n_split = -1.0, 1
        for cand in range(1, max(1, min_kb // 48) + 1):
            units = n_work * cand
            waves = -(-units // n_slots)
            if waves > 3:
                break
            eff = units / (waves * n_slots) - (0.03 if waves == 1 else 0.0)
            if eff > best + 1e-9:
                best, n_split = eff, cand
        for j, it in enumerate(chunk):
            w = L.WGradItem()
            w.g_act, w.x_act, w.g_row, w.x_row = it["g_act"], it["x_act"], it["g_row"], it["x_row"]
            w.m_valid, w.n_valid = it["m_valid"], it["n_valid"]
            w.n = max(16, ceil_to(it["n_valid"], 16))
            w.shift, w.t_lo, w.t_hi = it.get("shift", 0), it["t_lo"], it["t_hi"]
            w.n_split = n_split
            w.out = it["out"].data_ptr() + 4 * int(it.get("out_off", 0))
            w.out_rs, w.out_cs = int(it["out_rs"]), int(it["out_cs"])
            d.items[j] = w
        d.n_items = len(chunk)
        d.batch = int(batch)
        d.err = err.data_ptr() if err is not None else None
        d.pair_x = (2 if (_use_pair(tag) and all(it.n <= 256 for it in d.items[:len(chunk)])) else 1) if pair else 0
        out.append(("wgrad", d, tag))
    return out


def pack_wide_units(g_act, g_row, m_valid, t_lo, t_hi, chunks):
    """Group column chunks (dicts: x_act, x_row, n_valid, shift, out, out_off, out_rs, out_cs) that share one G tile
    into wgradw units: <= 3 chunks, <= 512 accumulator columns (each chunk rounded up to 32), <= 256 staged rows per
    CTA (64 for a chunk of <= 128 columns, else 128).  First-fit on the chunks sorted widest first."""
    units = []
    for ch in sorted(chunks, key=lambda c: -c["n_valid"]):
        n = max(16, ceil_to(ch["n_valid"], 16))
        assert n <= 256
        cols, rows = ceil_to(n, 32), (128 if n > 128 else 64)
        for u in units:
            if len(u["chunks"]) < L.WGW_MAX_CHUNKS and u["cols"] + cols <= 512 and u["rows"] + rows <= 256:
                break
        else:
            u = dict(g_act=g_act, g_row=g_row, m_valid=m_valid, t_lo=t_lo, t_hi=t_hi, chunks=[], cols=0, rows=0)
            units.append(u)
        u["chunks"].append(dict(ch, n=n))
        u["cols"] += cols
        u["rows"] += rows
    return units


def build_wgradw(acts, units, batch, err=None, tag=None, waves=1):
    """Descriptors of one logical wide-unit weight gradient.  Split-K factors are chosen per unit, proportional to the
    unit's estimated cost per K block (operand bytes per CTA vs MMA cycles, whichever is larger), so that all CTA pairs
    of a wave finish together."""
    sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    slots = (sms // 2) * waves
    out = []
    for i in range(0, len(units), L.WGW_MAX_UNITS):
        grp = units[i:i + L.WGW_MAX_UNITS]
        cost = []
        for u in grp:
            mma = sum(2.0 * c["n"] for c in u["chunks"])                      # cycles per K block: 4 x (n / 2)
            bytes_ = 16384 + 128 * u["rows"]
            kbs = batch * ((u["t_hi"] - u["t_lo"] + 31) // 32)
            cost.append(max(mma, bytes_ / 45.0) * kbs)
        tot = sum(cost)
        share = max(1, slots // max(1, -(-len(units) // L.WGW_MAX_UNITS)))    # pairs available to this launch
        splits = [max(1, int(round(share * c / tot))) for c in cost]
        while sum(splits) > share and max(splits) > 1:
            splits[splits.index(max(splits))] -= 1
        d = L.WGradWDesc()
        for j, a in enumerate(acts):
            d.acts[j] = a
        d.n_acts = len(acts)
        for j, u in enumerate(grp):
            w = L.WGWUnit()
            w.g_act, w.g_row, w.m_valid = int(u["g_act"]), int(u["g_row"]), int(u["m_valid"])
            w.t_lo, w.t_hi = int(u["t_lo"]), int(u["t_hi"])
            kbs = batch * ((u["t_hi"] - u["t_lo"] + 31) // 32)
            w.n_split = max(1, min(splits[j], kbs // 16))
            w.n_chunks = len(u["chunks"])
            for k, c in enumerate(u["chunks"]):
                cc = L.WGWChunk()
                cc.x_act, cc.x_row, cc.n, cc.n_valid = int(c["x_act"]), int(c["x_row"]), int(c["n"]), int(c["n_valid"])
                cc.shift = int(c.get("shift", 0))
                cc.out = c["out"].data_ptr() + 4 * int(c.get("out_off", 0))
                cc.out_rs, cc.out_cs = int(c["out_rs"]), int(c["out_cs"])
                w.chunk[k] = cc
            d.units[j] = w
        d.n_units = len(grp)
        d.batch = int(batch)
        d.err = err.data_ptr() if err is not None else None
        out.append(("wgradw", d, tag))
    return out


def act16_of(t16, channels=None):
    """aewn_act16 of a (B, T, Cp) fp16 channels-last tensor."""
    a = L.Act16()
    a.ptr = t16.data_ptr()
    a.t_rows, a.channels, a.batch = int(t16.shape[1]), int(channels if channels is not None else t16.shape[2]), int(t16.shape[0])
    a.row_pitch, a.batch_stride = int(t16.stride(1)), int(t16.stride(0))
    return a


def build_wgradh(acts16, units, batch, inv_scale_ptr, err=None, tag=None):
    """Descriptors of one wide-unit weight gradient on fp16 channels-last operands (aewn_wgradh); units as for build_wgradw
    (pack_wide_units), K blocks of 64 time steps."""
    sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    slots = sms // 2
    out = []
    for i in range(0, len(units), L.WGW_MAX_UNITS):
        grp = units[i:i + L.WGW_MAX_UNITS]
        cost = []
        for u in grp:
            mma = sum(2.0 * c["n"] for c in u["chunks"])                      # cycles per K block of 64: 4 x (n / 2)
            bytes_ = 16384 + 128 * u["rows"]
            kbs = batch * ((u["t_hi"] - u["t_lo"] + 63) // 64)
            cost.append(max(mma, bytes_ / 45.0) * kbs)
        tot = sum(cost)
        share = max(1, slots // max(1, -(-len(units) // L.WGW_MAX_UNITS)))
        splits = [max(1, int(round(share * c / tot))) for c in cost]
        while sum(splits) > share and max(splits) > 1:
            splits[splits.index(max(splits))] -= 1
        d = L.WGradHDesc()
        for j, a in enumerate(acts16):
            d.acts[j] = a
        d.n_acts = len(acts16)
        for j, u in enumerate(grp):
            w = L.WGWUnit()
            w.g_act, w.g_row, w.m_valid = int(u["g_act"]), int(u["g_row"]), int(u["m_valid"])
            w.t_lo, w.t_hi = int(u["t_lo"]), int(u["t_hi"])
            kbs = batch * ((u["t_hi"] - u["t_lo"] + 63) // 64)
            w.n_split = max(1, min(splits[j], kbs // 8))
            w.n_chunks = len(u["chunks"])
            for k, c in enumerate(u["chunks"]):
                cc = L.WGWChunk()
                cc.x_act, cc.x_row, cc.n, cc.n_valid = int(c["x_act"]), int(c["x_row"]), int(c["n"]), int(c["n_valid"])
                cc.shift = int(c.get("shift", 0))
                cc.out = c["out"].data_ptr() + 4 * int(c.get("out_off", 0))
                cc.out_rs, cc.out_cs = int(c["out_rs"]), int(c["out_cs"])
                w.chunk[k] = cc
            d.units[j] = w
        d.n_units = len(grp)
        d.batch = int(batch)
        d.err = err.data_ptr() if err is not None else None
        d.inv_scale = inv_scale_ptr
        out.append(("wgradh", d, tag))
    return out


# ------------------------------------------------------------------------------------------------- fused grad accumulation
# Fused gradient accumulation (opt-in PER PARAMETER SET: dist.FlatGradSync(..., fused_accumulate=True) marks the
# parameters it owns): the decoder's backward adds ALL its weight gradients into the parameters' existing .grad tensors
# with one aewn_add_blocks launch and returns None for them -- autograd would otherwise run one clone + one add kernel
# per parameter (~240 launches per step).  A custom Function cannot see which input gradients a particular backward
# call asks for, but it can ask the engine: the fused path is taken only when the running graph task will execute the
# parameters' AccumulateGrad nodes.  That is false inside torch.autograd.grad(loss, mel, retain_graph=True)
# (mfcc_inverter.py:103) and true inside the loss.backward() that follows (chassis.py:157), so the reference's own
# two-backward caller gets each weight gradient exactly once.  No process-global switch.
def mark_fused_accumulate(params, owner):
    for q in params:
        q._aewn_fused_owner = owner


def fused_accumulate_applies(leaves):
    """True iff every leaf was marked by a live FlatGradSync AND the current backward pass accumulates into them."""
    will = getattr(torch._C, "_will_engine_execute_node", None)
    if will is None:
        return False
    try:
        for q in leaves:
            owner = getattr(q, "_aewn_fused_owner", None)
            if owner is None or owner() is None or q.grad is None:
                return False
            # (looked up per call, never cached: a cached AccumulateGrad node stays bound to the stream it was created on
            # -- e.g. the warm-up stream -- and would put a cross-stream dependency into a later CUDA-graph capture)
            with torch.enable_grad():           # backward runs with grad mode off: view_as would have no grad_fn
                node = q.view_as(q).grad_fn.next_functions[0][0]     # the leaf's AccumulateGrad
            if not will(node):
                return False
    except RuntimeError:
        return False
    return True


_grad_tables = {}


def add_into_grads(pairs):
    """pairs: [(src_view, dst)], dst contiguous, src of the same shape with contiguous inner dimensions.  Returns False
    (and does nothing) if any destination is missing or incompatible."""
    for src, dst in pairs:
        if dst is None or not dst.is_contiguous() or dst.shape != src.shape or dst.dtype != torch.float32:
            return False
    key = tuple((src.data_ptr(), dst.data_ptr()) for src, dst in pairs)
    ent = _grad_tables.get(key)
    if ent is None:
        import numpy as np
        blocks = []
        for src, dst in pairs:
            rows = int(src.shape[0]) if src.dim() > 0 else 1
            cols = int(src.numel() // max(rows, 1))
            inner_ok = src.dim() <= 1 or src[0].is_contiguous()
            if not inner_ok:
                return False
            si = int(src.stride(0)) if src.dim() > 0 else 1
            step = max(1, 8192 // max(cols, 1))            # one CTA per table entry: keep entries <= 8 K elements
            for r0 in range(0, rows, step):
                nr = min(step, rows - r0)
                blocks.append((src.data_ptr() + 4 * r0 * si, dst.data_ptr() + 4 * r0 * cols, nr, cols, si, 1, cols))
        dt = np.dtype([("src", "<u8"), ("dst", "<u8"), ("ni", "<i4"), ("nj", "<i4"), ("si", "<i8"), ("sj", "<i8"),
                       ("di", "<i8")])
        arr = np.array(blocks, dtype=dt)
        if len(_grad_tables) >= 4:
            _grad_tables.clear()
        ent = _grad_tables[key] = (torch.from_numpy(arr.view(np.uint8).reshape(-1).copy()).to(pairs[0][1].device),
                                   len(blocks))
    L.check(L.lib().aewn_add_blocks(C.c_void_p(ent[0].data_ptr()), C.c_int(ent[1]), _stream()), "aewn_add_blocks")
    return True


def run_launches(launches):
    """Enqueue prebuilt descriptors on the current stream (ctypes call only: ~2 us of host time per launch)."""
    lib = L.lib()
    st = _stream()
    for kind, d, tag in launches:
        e0 = _prof_begin(tag)
        if kind == "tgemm":
            L.check(lib.aewn_tgemm(C.byref(d), st), "aewn_tgemm")
        elif kind == "grcc":
            L.check(lib.aewn_grcc_fwd(C.byref(d), st), "aewn_grcc_fwd")
        elif kind == "dgrad16":
            L.check(lib.aewn_grcc_dgrad(C.byref(d), st), "aewn_grcc_dgrad")
        elif kind == "gz16":
            L.check(lib.aewn_grcc_gz(C.byref(d), st), "aewn_grcc_gz")
        elif kind == "cvt16s":
            L.check(lib.aewn_cvt_f16_cl_scaled(C.c_void_p(d[0]), C.c_longlong(d[1]), C.c_longlong(d[2]), C.c_void_p(d[3]),
                                               C.c_longlong(d[4]), C.c_int(d[5]), C.c_int(d[6]), C.c_int(d[7]), C.c_int(d[8]),
                                               C.c_void_p(d[9]), C.c_void_p(d[10]), st), "aewn_cvt_f16_cl_scaled")
        elif kind == "amax":
            L.check(lib.aewn_amax_pow2_scale(C.c_void_p(d[0]), C.c_longlong(d[1]), C.c_float(d[2]), C.c_void_p(d[3]),
                                             C.c_void_p(d[4]), st), "aewn_amax_pow2_scale")
        elif kind == "wgradw":
            L.check(lib.aewn_wgradw(C.byref(d), st), "aewn_wgradw")
        elif kind == "wgradh":
            L.check(lib.aewn_wgradh(C.byref(d), st), "aewn_wgradh")
        else:
            L.check(lib.aewn_wgrad(C.byref(d), st), "aewn_wgrad")
        _prof_end(tag, e0)


def tgemm(acts, segs, w, ntiles, batch, t_begin, t_end, err=None, tag=None):
    run_launches(build_tgemm(acts, segs, w, ntiles, batch, t_begin, t_end, err, tag))


def wgrad(acts, items, batch, err=None, tag=None):
    run_launches(build_wgrad(acts, items, batch, err, tag))


def chunks(total, size=256):
    """[(offset, width)] covering `total` output channels in n-tiles of at most `size`."""
    return [(o, min(size, total - o)) for o in range(0, total, size)]


# ------------------------------------------------------------------------------------------------- stack geometry
def _pad_k(m, k):
    return torch.nn.functional.pad(m, (0, k - m.shape[1]))


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

    def is_final(self, l):
        return l == self.L - 1 and self.last_is_final


def needs_dup(d):
    """TMA box origins must be multiples of 4 elements: taps with d % 4 != 0 read a pre-shifted duplicate."""
    return d % 4 != 0


LAYER_KEYS = ("conv_signal.weight", "conv_signal.bias", "conv_gate.weight", "conv_gate.bias", "proj_signal.weight",
              "proj_gate.weight", "dil_skp.weight", "dil_res.weight")


class StackPlan:
    """Everything one (batch, widths, geometry, parameter set) configuration needs, built ONCE and replayed every step:
    persistent zero-initialised workspaces, K-major packed weight buffers plus the block-copy table that
    refreshes them from the live parameters with one launch, prebuilt launch descriptors for forward and backward, and
    one flat gradient buffer.  Per step the host only issues ~7 ctypes calls per layer.

    Packed operand matrices of layer l:
      w1  [256*J][2*KR + KC]  per 128-channel block: 128 filt rows then 128 gate rows; columns tap0 | tap1 | cond | bias
      w2  [R + S][KD]         dil_res rows (absent in the final layer) then dil_skp rows
      w2t [D][KR + KS]        [Wr^T | Ws^T]
      w1t [R + C][2*K2]       [tap0^T | tap1^T] over (g_f ; g_g); cond rows only under the unshifted block
    """

    def __init__(self, B, R, D, S, Cc, geom, params, device, relu_last):
        g = self.geom = geom
        if D > 256:
            raise NotImplementedError(f"aewn: n_dil = {D} > 256: the gate-derivative kernel keeps all dilation channels "
                                      "of a time tile in one 256-column accumulator (the reference presets use 256)")
        self.B, self.R, self.D, self.S, self.Cc = B, R, D, S, Cc
        self.device = device
        self.relu_last = relu_last
        self.generation = 0
        self.param_ptrs = self._ptrs(params)
        self.has_bias = "conv_signal.bias" in params[0]
        Tp = g.Tp
        n_sig = g.L + (0 if g.last_is_final else 1)
        self.sig = [new_buf(B, R, Tp, device) for _ in range(n_sig)]       # sig[l] = input of layer l
        self.fused = FUSED_FWD and fused_ok(R, D, S, Cc)
        # AEWN_DGRAD16: 2 (default) = the fused-layer engine with fp16 operands and a per-step power-of-two scale taken from
        # max|g_skp| (10-bit mantissa like TF32, round-to-nearest; overflow is reported as AEWN_ERR_RANGE), 1 = bf16 operands
        # (no scale, 8-bit mantissa), 0 = the TF32 tgemm engine -- DESIGN.md 4.1c
        mode16 = dgrad16_mode()
        self.dgrad16 = self.fused and R % 16 == 0 and Cc <= 240 and (2 * D) % 64 == 0 and mode16 in ("1", "2")
        self.dgrad16_scaled = self.dgrad16 and mode16 == "2"
        # AEWN_WGRAD16 (default 1, needs the scaled fp16 gradient copy): ALL weight gradients of the stack on aewn_wgradh,
        # from fp16 channels-last copies: x16 (then kept per layer), cond16, z16 (written by the forward instead of the fp32
        # z), the scaled copies of [g_f; g_g], g_x and g_skp -- DESIGN.md 4.2c
        self.wgrad16 = self.dgrad16_scaled and os.environ.get("AEWN_WGRAD16", "1") == "1" and D % 64 == 0 and \
            g.last_is_final and ENGINE_MODE == "auto" and WIDE_WGRAD      # (the narrow TF32 units read the fp32 z / [g_f; g_g])
        # AEWN_GZ16 (default 1; needs the fp16 weight-gradient copies): the gate derivative's GEMM on the fused-layer engine from the
        # scaled fp16 copies of g_x / g_skp (aewn_grcc_gz) instead of the TF32 tgemm launch reading the fp32 tensors
        self.gz16 = self.wgrad16 and os.environ.get("AEWN_GZ16", "1") == "1"
        # pre-shifted duplicates of the layer inputs for dilations 1 and 2: only the TF32 weight gradients read them
        self.xs = {} if self.wgrad16 else {l: new_buf(B, R, Tp, device) for l in range(g.L) if needs_dup(g.dils[l])}
        # saved for the backward pass: tanh and sigmoid (fp32), or -- fused forward -- ONE word per element holding the two
        # gate-derivative factors {fp16 a = sg (1 - th^2), fp16 b = th sg (1 - sg)} in `th` (half the bytes; `sg` unused)
        self.th = [new_buf(B, D, Tp, device) for _ in range(g.L)]
        self.sg = None if self.fused else [new_buf(B, D, Tp, device) for _ in range(g.L)]
        self.z = None if self.wgrad16 else [new_buf(B, D, Tp, device) for _ in range(g.L)]
        self.z16 = [torch.zeros(B, Tp, D, device=device, dtype=torch.float16) for _ in range(g.L)] if self.wgrad16 else None
        self.skp = new_buf(B, S, Tp, device)
        self.cond = new_buf(B, Cc + 1, Tp, device)
        self.cond[:, Cc, :] = 1.0                                            # the bias channel
        self.err = torch.zeros(1, dtype=torch.int32, device=device)
        self.nan_const = torch.full((1,), float("nan"), device=device)      # see _DecoderCoreFn.forward (error poison)
        self.zero_const = torch.zeros(1, device=device)
        self.KR, self.KC, self.KD = ceil_to(R, 32), ceil_to(Cc + 1, 32), ceil_to(D, 32)
        self.KS, self.K2 = ceil_to(S, 32), ceil_to(2 * D, 32)
        self.J = (D + 127) // 128
        if self.fused:
            # fp16 channels-last operand copies: the layer input is kept per layer when the fp16 weight
            # gradients read it again, else it ping-pongs between two buffers (the
            # backward pass reads the fp32 tensors), the conditioning copy is shared by all layers
            self.KR16, self.KC16 = ceil_to(R, 64), ceil_to(Cc + 1, 64)
            self.n_x16 = (g.L + 1) if self.wgrad16 else 2                  # kept per layer when the backward reads them
            self.x16 = [torch.zeros(B, Tp, self.KR16, device=device, dtype=torch.float16) for _ in range(self.n_x16)]
            self.c16 = torch.zeros(B, Tp, self.KC16, device=device, dtype=torch.float16)
        self._build_packs(params)
        self.fwd_train = self._build_forward(save=True)
        self.fwd_infer = None
        self._bwd = None

    # ---------------------------------------------------------------------------------------------- parameters
    @staticmethod
    def _ptrs(params):
        return tuple(p[k].data_ptr() for p in params for k in LAYER_KEYS if k in p)

    def matches(self, params):
        return self._ptrs(params) == self.param_ptrs

    def _build_packs(self, params):
        g, R, D, S, Cc, dev = self.geom, self.R, self.D, self.S, self.Cc, self.device
        KR, KC, KD, KS, K2, J = self.KR, self.KC, self.KD, self.KS, self.K2, self.J
        KP = 2 * KR + KC
        self.w1, self.w2, self.w2t, self.w1t = [], [], [], []
        self.w1h, self.w2h, self.w1t16, self.w2t16 = [], [], [], []
        blocks, hblocks, bblocks = [], [], []
        fused = self.fused

        def blkb(src, s_off, dst, d_row, d_col, ni, nj, si, sj):       # same table entry, bf16 destination
            step = max(1, 8192 // max(nj, 1))
            for r0 in range(0, ni, step):
                bblocks.append((src.data_ptr() + 4 * (s_off + r0 * si),
                                dst.data_ptr() + 2 * ((d_row + r0) * dst.shape[1] + d_col), min(step, ni - r0), nj, si, sj,
                                dst.shape[1]))

        def blkh(src, s_off, dst, d_row, d_col, ni, nj, si, sj):       # same table entry, fp16 destination
            step = max(1, 8192 // max(nj, 1))
            for r0 in range(0, ni, step):
                hblocks.append((src.data_ptr() + 4 * (s_off + r0 * si),
                                dst.data_ptr() + 2 * ((d_row + r0) * dst.shape[1] + d_col), min(step, ni - r0), nj, si, sj,
                                dst.shape[1]))

        def blk(src, s_off, dst, d_row, d_col, ni, nj, si, sj):
            step = max(1, 8192 // max(nj, 1))              # one CTA per table entry: keep entries <= 8 K elements
            for r0 in range(0, ni, step):
                blocks.append((src.data_ptr() + 4 * (s_off + r0 * si),
                               dst.data_ptr() + 4 * ((d_row + r0) * dst.shape[1] + d_col), min(step, ni - r0), nj, si, sj,
                               dst.shape[1]))

        for l in range(g.L):
            p = params[l]
            final = g.is_final(l)
            # forward operand matrices: fp16 for the fused layer kernel, else fp32 (TF32) for aewn_tgemm
            if fused:
                KR16, KC16 = self.KR16, self.KC16
                w1 = torch.zeros(256 * J, 2 * KR16 + KC16, device=dev, dtype=torch.float16)
                w2 = torch.zeros((0 if final else R) + S, D, device=dev, dtype=torch.float16)
                fblk, c_tap1, c_cond = blkh, KR16, 2 * KR16
            else:
                w1 = torch.zeros(256 * J, KP, device=dev)
                w2 = torch.zeros((0 if final else R) + S, KD, device=dev)
                fblk, c_tap1, c_cond = blk, KR, 2 * KR
            w2t = torch.zeros(D, KR + KS, device=dev)
            w1t = torch.zeros(R + Cc, 2 * K2, device=dev)
            wf, wg = p["conv_signal.weight"], p["conv_gate.weight"]          # (D, R, 2)
            pf, pg = p["proj_signal.weight"], p["proj_gate.weight"]          # (D, Cc, 1)
            for j in range(J):
                nj_ = min(128, D - 128 * j)
                for h, (wc, pj, bk) in enumerate(((wf, pf, "conv_signal.bias"), (wg, pg, "conv_gate.bias"))):
                    row = 256 * j + 128 * h
                    fblk(wc, 128 * j * R * 2 + 0, w1, row, 0, nj_, R, 2 * R, 2)         # tap 0
                    fblk(wc, 128 * j * R * 2 + 1, w1, row, c_tap1, nj_, R, 2 * R, 2)    # tap 1
                    fblk(pj, 128 * j * Cc, w1, row, c_cond, nj_, Cc, Cc, 1)             # cond projection
                    if bk in p:
                        fblk(p[bk], 128 * j, w1, row, c_cond + Cc, nj_, 1, 1, 1)        # bias on the ones channel
            ws = p["dil_skp.weight"]                                                   # (S, D, 1)
            if not final:
                wr = p["dil_res.weight"]                                               # (R, D, 1)
                fblk(wr, 0, w2, 0, 0, R, D, D, 1)
                fblk(ws, 0, w2, R, 0, S, D, D, 1)
                blk(wr, 0, w2t, 0, 0, D, R, 1, D)                                      # Wr^T
            else:
                fblk(ws, 0, w2, 0, 0, S, D, D, 1)
            blk(ws, 0, w2t, 0, KR, D, S, 1, D)                                         # Ws^T
            if self.gz16:                                                              # the same two blocks in fp16
                w2t16 = torch.zeros(D, self.KR16 + ceil_to(S, 64), device=dev, dtype=torch.float16)
                if not final:
                    blkh(wr, 0, w2t16, 0, 0, D, R, 1, D)
                blkh(ws, 0, w2t16, 0, self.KR16, D, S, 1, D)
                self.w2t16.append(w2t16)
            if self.dgrad16:
                w1t = torch.zeros(R + Cc, 2 * K2, device=dev,
                                  dtype=torch.float16 if self.dgrad16_scaled else torch.bfloat16)
                tblk = blkh if self.dgrad16_scaled else blkb
            else:
                tblk = blk
            for h, wc in enumerate((wf, wg)):
                for k in (0, 1):                                                       # tap k transposed
                    tblk(wc, k, w1t, 0, k * K2 + h * D, R, D, 2, 2 * R)
            for h, pj in enumerate((pf, pg)):
                tblk(pj, 0, w1t, R, K2 + h * D, Cc, D, 1, Cc)                          # P^T under the unshifted block
            (self.w1h if fused else self.w1).append(w1)
            (self.w2h if fused else self.w2).append(w2)
            self.w2t.append(w2t)
            (self.w1t16 if self.dgrad16 else self.w1t).append(w1t)
        import numpy as np
        dt = np.dtype([("src", "<u8"), ("dst", "<u8"), ("ni", "<i4"), ("nj", "<i4"), ("si", "<i8"), ("sj", "<i8"),
                       ("di", "<i8")])
        arr = np.array(blocks, dtype=dt)
        assert dt.itemsize == C.sizeof(L.CopyBlock)
        self.n_blocks = len(blocks)
        self.block_table = torch.from_numpy(arr.view(np.uint8).reshape(-1).copy()).to(dev)
        self.n_hblocks = len(hblocks)
        if hblocks:
            self.hblock_table = torch.from_numpy(np.array(hblocks, dtype=dt).view(np.uint8).reshape(-1).copy()).to(dev)
        self.n_bblocks = len(bblocks)
        if bblocks:
            self.bblock_table = torch.from_numpy(np.array(bblocks, dtype=dt).view(np.uint8).reshape(-1).copy()).to(dev)

    def repack(self):
        L.check(L.lib().aewn_pack_blocks(C.c_void_p(self.block_table.data_ptr()), C.c_int(self.n_blocks), _stream()),
                "aewn_pack_blocks")
        if self.n_hblocks:
            L.check(L.lib().aewn_pack_blocks_f16(C.c_void_p(self.hblock_table.data_ptr()), C.c_int(self.n_hblocks),
                                                 _stream()), "aewn_pack_blocks_f16")
        if self.n_bblocks:
            L.check(L.lib().aewn_pack_blocks_bf16(C.c_void_p(self.bblock_table.data_ptr()), C.c_int(self.n_bblocks),
                                                  _stream()), "aewn_pack_blocks_bf16")

    def _to_f16_cl(self, src, dst, channels):
        """(B, C, Tp) fp32 workspace -> (B, Tp, Cp) fp16 channels-last operand copy (one launch)."""
        L.check(L.lib().aewn_cvt_f16_cl(
            C.c_void_p(src.data_ptr()), C.c_longlong(src.stride(0)), C.c_longlong(src.stride(1)),
            C.c_void_p(dst.data_ptr()), C.c_longlong(dst.stride(0)), C.c_int(dst.shape[2]), C.c_int(channels),
            C.c_int(self.geom.T0), C.c_int(self.B), C.c_int(-1), C.c_void_p(self.err.data_ptr()), _stream()),
            "aewn_cvt_f16_cl")

    def _build_forward_fused(self, save):
        """One aewn_grcc_fwd launch per layer (wavenet.py:91-111 in one kernel)."""
        g, B, R, D, S, Cc = self.geom, self.B, self.R, self.D, self.S, self.Cc
        T0, Tp = g.T0, g.Tp
        rf4 = g.RF & ~3
        out = []
        for l, dil in enumerate(g.dils):
            final = g.is_final(l)
            lo = g.lead[l]
            x, xin, xout = self.sig[l], self.x16[l % self.n_x16], self.x16[(l + 1) % self.n_x16]
            d = L.GrccFwdDesc()
            d.x16, d.x16_bs, d.x16_cp = xin.data_ptr(), int(xin.stride(0)), self.KR16
            d.c16, d.c16_bs, d.c16_cp = self.c16.data_ptr(), int(self.c16.stride(0)), self.KC16
            d.t_rows = Tp
            d.w1h, d.w1_k, d.w2h = self.w1h[l].data_ptr(), int(self.w1h[l].shape[1]), self.w2h[l].data_ptr()
            d.x32, d.x_bs, d.x_cs = x.data_ptr(), int(x.stride(0)), int(x.stride(1))
            if not final:
                d.xo32, d.xo16 = self.sig[l + 1].data_ptr(), xout.data_ptr()
                d_next = g.dils[l + 1] if l + 1 < g.L else 4
                if save and l + 1 < g.L and (l + 1) in self.xs:     # the backward pass's TF32 weight-gradient tap
                    d.dup, d.dup_toff, d.dup_t_hi = self.xs[l + 1].data_ptr(), d_next, T0
            if save:
                d.th, d.save = self.th[l].data_ptr(), 2                                  # packed derivative factors
                if self.wgrad16:
                    d.z16, d.z16_bs, d.z16_cp = self.z16[l].data_ptr(), int(self.z16[l].stride(0)), D
                else:
                    d.z = self.z[l].data_ptr()
                d.a_bs, d.a_cs = int(self.th[l].stride(0)), int(self.th[l].stride(1))
            d.skp, d.s_bs, d.s_cs = self.skp.data_ptr(), int(self.skp.stride(0)), int(self.skp.stride(1))
            last_relu = l == g.L - 1 and self.relu_last
            d.skp_mode = (3 if last_relu else 0) if l == 0 else (2 if last_relu else 1)
            d.batch, d.R, d.D, d.S, d.n_cond1, d.dil, d.final_layer = B, R, D, S, Cc + 1, dil, int(final)
            d.t_lo, d.t_zero_lo, d.t_hi = lo & ~3, lo, T0
            d.skp_t_lo, d.skp_zero_lo = rf4, g.RF
            d.err = self.err.data_ptr()
            out.append(("grcc", d, f"fwd_layer.{l}"))
        return out

    # ---------------------------------------------------------------------------------------------- forward
    def _build_forward(self, save):
        if self.fused:
            return self._build_forward_fused(save)
        g, B, R, D, S, Cc = self.geom, self.B, self.R, self.D, self.S, self.Cc
        T0 = g.T0
        out = []
        for l, d in enumerate(g.dils):
            final = g.is_final(l)
            lo = g.lead[l]
            lo4 = lo & ~3
            t_begin = lo4 & ~31
            x = self.sig[l]
            xa, ca = act_of(x, T0), act_of(self.cond, T0)
            if needs_dup(d):
                acts = [act_of(self.xs[l], T0), xa, ca]
                segs = [(0, 0, R, 0), (1, 0, R, self.KR), (2, 0, Cc + 1, 2 * self.KR)]
            else:
                acts = [xa, ca]
                segs = [(0, -d, R, 0), (0, 0, R, self.KR), (1, 0, Cc + 1, 2 * self.KR)]
            tiles = []
            for j in range(self.J):
                c0, nj = 128 * j, min(128, D - 128 * j)
                tiles.append(ntile(256 * j, nj, self.th[l][:, c0:] if save else None, mode=L.EPI_GATE_FWD, n=256,
                                   out2=self.sg[l][:, c0:] if save else None, out3=self.z[l][:, c0:],
                                   t_lo=lo4, t_hi=T0, t_zero_lo=lo))
            out += build_tgemm(acts, segs, self.w1[l], tiles, B, t_begin, T0, self.err, tag=f"fwd_gemm1.{l}")
            tiles = []
            if not final:
                d_next = g.dils[l + 1] if l + 1 < g.L else 4
                for (c0, n) in chunks(R):
                    tiles.append(ntile(c0, n, self.sig[l + 1][:, c0:], add=x[:, c0:], t_lo=lo4, t_hi=T0, t_zero_lo=lo,
                                       out2=self.xs[l + 1][:, c0:] if (l + 1 < g.L and needs_dup(d_next)) else None,
                                       dup_toff=d_next, dup_t_hi=T0))
            rf4 = g.RF & ~3
            flags = (L.F_ACCUM if l > 0 else 0) | (L.F_RELU if (l == g.L - 1 and self.relu_last) else 0)
            row0 = 0 if final else R
            for (c0, n) in chunks(S):
                tiles.append(ntile(row0 + c0, n, self.skp[:, c0:], flags=flags, t_lo=rf4, t_hi=T0,
                                   t_zero_lo=g.RF if l == 0 else 0))
            out += build_tgemm([act_of(self.z[l], T0)], [(0, 0, D, 0)], self.w2[l], tiles, B, t_begin, T0, self.err,
                               tag=f"fwd_gemm2.{l}")
        return out

    def forward(self, save=True):
        """Inputs already staged: sig[0] (and xs[0]) = base-layer output, cond.  Result: skp (B, S, Tp) valid on
        [RF, T0), ReLU applied when relu_last (wavenet.py:359)."""
        self.repack()
        if self.fused:
            self._to_f16_cl(self.sig[0], self.x16[0], self.R)
            self._to_f16_cl(self.cond, self.c16, self.Cc + 1)
        if save:
            run_launches(self.fwd_train)
        else:
            if self.fwd_infer is None:
                self.fwd_infer = self._build_forward(save=False)
            run_launches(self.fwd_infer)

    # ---------------------------------------------------------------------------------------------- backward
    def _grad_layout(self):
        """One flat fp32 buffer holding every weight gradient (reference layouts) + the (2D, Cc+1) proj/bias scratch."""
        g, R, D, S, Cc = self.geom, self.R, self.D, self.S, self.Cc
        off = 0
        lay = []
        for l in range(g.L):
            e = {}
            for k, shape in (("conv_signal.weight", (D, R, 2)), ("conv_gate.weight", (D, R, 2)),
                             ("dpb", (2 * D, Cc + 1)), ("dil_skp.weight", (S, D, 1)), ("dil_res.weight", (R, D, 1))):
                if k == "dil_res.weight" and g.is_final(l):
                    continue
                n = 1
                for v in shape:
                    n *= v
                e[k] = (off, shape)
                off += ceil_to(n, 4)
            lay.append(e)
        return lay, off

    def bwd(self):
        if self._bwd is not None:
            return self._bwd
        g, B, R, D, S, Cc, dev = self.geom, self.B, self.R, self.D, self.S, self.Cc, self.device
        T0, Tp = g.T0, g.Tp
        bw = dict(gfg=new_buf(B, 2 * D, Tp, dev), gfs=new_buf(B, 2 * D, Tp, dev),
                  gx=[new_buf(B, R, Tp, dev), new_buf(B, R, Tp, dev)], g_skp=new_buf(B, S, Tp, dev),
                  g_cond=new_buf(B, Cc, Tp, dev), g_last=None if g.last_is_final else new_buf(B, R, Tp, dev))
        if self.dgrad16:
            bw["g16"] = torch.zeros(B, Tp, self.K2, device=dev,                            # [g_f; g_g], channels-last
                                    dtype=torch.float16 if self.dgrad16_scaled else torch.bfloat16)
            if self.dgrad16_scaled:
                bw["gscale"] = torch.ones(2, device=dev)                   # scale, 1 / scale (aewn_amax_pow2_scale)
                bw["gscale_work"] = torch.zeros(1, device=dev, dtype=torch.int32)
            if self.wgrad16:
                # scaled fp16 channels-last copies of the gradients the dil_skp / dil_res weight gradients contract with z
                bw["g_skp16"] = torch.zeros(B, Tp, ceil_to(S, 64), device=dev, dtype=torch.float16)
                bw["gx16"] = [torch.zeros(B, Tp, self.KR16, device=dev, dtype=torch.float16) for _ in range(2)]
        lay, total = self._grad_layout()
        flat = torch.zeros(total, device=dev)
        views = [{k: flat[o:o + int(torch.tensor(sh).prod())].view(sh) for k, (o, sh) in e.items()} for e in lay]
        bw["flat"], bw["views"] = flat, views
        gfg, gfs, g_skp, g_cond = bw["gfg"], bw["gfs"], bw["g_skp"], bw["g_cond"]
        rf4 = g.RF & ~3
        launches = []
        if self.dgrad16_scaled:
            # every layer's [g_f; g_g] derives from g_skp (and the chain through g_x): its maximum sets the common scale, with
            # 2^13 of headroom above (max|g_skp| * scale in (4, 8]) and 17 binades of normal fp16 range below
            launches.append(("amax", (g_skp.data_ptr(), g_skp.numel(), 8.0, bw["gscale_work"].data_ptr(),
                                      bw["gscale"].data_ptr()), "bwd_amax"))
            if self.wgrad16:
                k16 = bw["g_skp16"]
                launches.append(("cvt16s", (g_skp.data_ptr(), int(g_skp.stride(0)), int(g_skp.stride(1)), k16.data_ptr(),
                                            int(k16.stride(0)), int(k16.shape[2]), S, T0, B, bw["gscale"].data_ptr(),
                                            self.err.data_ptr()), "bwd_amax"))
        g_sig = bw["g_last"]                     # gradient w.r.t. the output of the layer being processed
        for l in range(g.L - 1, -1, -1):
            d = g.dils[l]
            final = g.is_final(l)
            lo, lo_prev = g.lead[l], g.lead_in(l)
            lo4, lop4 = lo & ~3, lo_prev & ~3
            x = self.sig[l]
            v = views[l]
            # (1) g_z = Wr^T g_sig + Ws^T g_skp, then the gate derivative -> gfg = [g_f ; g_g]   
            if g_sig is not None:
                acts = [act_of(g_sig, T0), act_of(g_skp, T0)]
                segs = [(0, 0, R, 0), (1, 0, S, self.KR)]
            else:
                acts = [act_of(g_skp, T0)]
                segs = [(0, 0, S, self.KR)]
            t_store = min(lop4, lo4)
            tile = ntile(0, D, gfg, mode=L.EPI_GATE_BWD, out2=gfg[:, D:], add=self.th[l],
                         add2=None if self.fused else self.sg[l],
                         flags=(L.F_AB16 if self.fused else 0) | (L.F_NO_OUT32 if self.wgrad16 else 0),
                         out3=gfs if (needs_dup(d) and not self.dgrad16) else None, dup_toff=-d, dup_t_hi=T0,
                         t_lo=t_store, t_hi=T0, t_zero_lo=lo)
            if self.dgrad16:
                tile.out16, tile.out16_bs, tile.out16_cp = bw["g16"].data_ptr(), int(bw["g16"].stride(0)), self.K2
                if self.dgrad16_scaled:
                    tile.out16_scale = bw["gscale"].data_ptr()
            if self.gz16:
                gz = L.GrccGzDesc()
                if g_sig is not None:
                    gxs = bw["gx16"][(l + 1) % 2]
                    gz.gx16, gz.gx16_bs, gz.gx16_cp = gxs.data_ptr(), int(gxs.stride(0)), self.KR16
                k16 = bw["g_skp16"]
                gz.gs16, gz.gs16_bs, gz.gs16_cp, gz.t_rows = k16.data_ptr(), int(k16.stride(0)), int(k16.shape[2]), Tp
                gz.w2t16, gz.w_k, gz.w_koff_skp = self.w2t16[l].data_ptr(), int(self.w2t16[l].shape[1]), self.KR16
                gz.ab, gz.a_bs, gz.a_cs = self.th[l].data_ptr(), int(self.th[l].stride(0)), int(self.th[l].stride(1))
                gz.g16, gz.g16_bs, gz.g16_cp, gz.gg_off = bw["g16"].data_ptr(), int(bw["g16"].stride(0)), self.K2, D
                gz.batch, gz.D = B, D
                gz.t_lo, gz.t_zero_lo, gz.t_hi = t_store, lo, T0
                gz.err = self.err.data_ptr()
                launches.append(("gz16", gz, f"bwd_gz.{l}"))
            else:
                launches += build_tgemm(acts, segs, self.w2t[l], [tile], B, t_store & ~31, T0, self.err, tag=f"bwd_gz.{l}")
            # (2) g_x[tau] = tap1^T gfg[tau] + tap0^T gfg[tau + d] (+ g_sig[tau]);  g_cond[tau] += P^T gfg[tau]
            gx = bw["gx"][l % 2]
            if self.dgrad16:
                dd = L.GrccDgradDesc()
                g16 = bw["g16"]
                dd.g16, dd.g16_bs, dd.g16_cp, dd.t_rows = g16.data_ptr(), int(g16.stride(0)), self.K2, Tp
                dd.w1t16, dd.w_k = self.w1t16[l].data_ptr(), 2 * self.K2
                dd.g_sig = g_sig.data_ptr() if g_sig is not None else None
                dd.gx, dd.x_bs, dd.x_cs, dd.add_t_lo = gx.data_ptr(), int(gx.stride(0)), int(gx.stride(1)), lo
                dd.g_cond, dd.c_bs, dd.c_cs, dd.n_cond = g_cond.data_ptr(), int(g_cond.stride(0)), int(g_cond.stride(1)), Cc
                dd.batch, dd.R, dd.dil = B, R, d
                dd.t_lo, dd.t_zero_lo, dd.t_hi = lop4, lo_prev, T0
                dd.cond_t_lo, dd.cond_zero_lo = lo & ~3, lo
                dd.err = self.err.data_ptr()
                if self.dgrad16_scaled:
                    dd.g_inv_scale = bw["gscale"].data_ptr() + 4
                if self.wgrad16 and l > 0:         # the layer below contracts this gradient with its z (dil_res)
                    gx16 = bw["gx16"][l % 2]
                    dd.gx16, dd.gx16_bs, dd.gx16_cp = gx16.data_ptr(), int(gx16.stride(0)), self.KR16
                launches.append(("dgrad16", dd, f"bwd_dgrad.{l}"))
            if not self.dgrad16:
                if needs_dup(d):
                    acts = [act_of(gfs, T0 - d), act_of(gfg, T0)]
                    segs = [(0, 0, 2 * D, 0), (1, 0, 2 * D, self.K2)]
                else:
                    acts = [act_of(gfg, T0)]
                    segs = [(0, d, 2 * D, 0), (0, 0, 2 * D, self.K2)]
                tiles = []
                for (c0, n) in chunks(R):
                    tiles.append(ntile(c0, n, gx[:, c0:], add=g_sig[:, c0:] if g_sig is not None else None, add_t_lo=lo,
                                       t_lo=lop4, t_hi=T0, t_zero_lo=lo_prev))
                cond_tiles = [ntile(R + c0, n, g_cond[:, c0:], flags=L.F_ACCUM, seg_mask=2, t_lo=lo, t_hi=T0)
                              for (c0, n) in chunks(Cc)]
                # The tail tile of g_x (R - 256 = 112 columns) and the g_cond tile (138 -> 144) fit ONE 256-column accumulator and
                # their rows are adjacent in w1t: one pass over [g_f; g_g] instead of two (the cond rows hold zeros under the
                # shifted tap block, so running them over both segments adds exact zeros).  AEWN_F_MERGE_NEXT, pair mode only.
                tail, first = tiles[-1], cond_tiles[0]
                if (MERGE_DGRAD_TILES and _use_pair("bwd_dgrad") and tail.n == tail.n_valid and first.n >= 32 and
                        tail.n + first.n <= 256 and tail.w_row + tail.n == first.w_row):
                    tail.flags |= L.F_MERGE_NEXT
                tiles += cond_tiles
                launches += build_tgemm(acts, segs, self.w1t[l], tiles, B, lop4 & ~31, T0, self.err, tag=f"bwd_dgrad.{l}")
            # (3) weight gradients of conv_signal/conv_gate/proj_signal/proj_gate (+ biases via the ones channel)
            x0_act = act_of(self.xs[l], T0) if l in self.xs else act_of(x, T0)     # (TF32 engines only)
            acts = [act_of(gfg, T0), x0_act, act_of(x, T0), act_of(self.cond, T0)]
            sh0 = 0 if l in self.xs else -d
            mtiles = [(h, i) for h in (0, 1) for i in range((D + 127) // 128)]
            keys = ("conv_signal.weight", "conv_gate.weight")
            wide = ENGINE_MODE == "auto" and WIDE_WGRAD
            if wide and self.wgrad16:
                # fp16 channels-last operands (aewn_wgradh): G = the scaled copy of [g_f; g_g], X = [x16(tau - d) | x16(tau) |
                # cond16, 1]; a tap is a row shift, the K range starts at the lead itself
                acts16 = [act16_of(bw["g16"]), act16_of(self.x16[l % self.n_x16]), act16_of(self.c16)]
                units = []
                for h in (0, 1):
                    cks = []
                    for (c0, n) in chunks(R):
                        for tap, sh in enumerate((-d, 0)):
                            cks.append(dict(x_act=1, x_row=c0, n_valid=n, shift=sh, out=v[keys[h]], out_off=c0 * 2 + tap,
                                            out_rs=2 * R, out_cs=2))
                    for (c0, n) in chunks(Cc + 1):
                        cks.append(dict(x_act=2, x_row=c0, n_valid=n, shift=0, out=v["dpb"],
                                        out_off=h * D * (Cc + 1) + c0, out_rs=Cc + 1, out_cs=1))
                    units += pack_wide_units(0, h * D, D, lo, T0, cks)
                launches += build_wgradh(acts16, units, B, bw["gscale"].data_ptr() + 4, self.err, tag=f"wgrad1.{l}")
            elif wide:
                # wide units (aewn_wgradw): M = the D filt (h = 0) or gate (h = 1) rows of gfg as ONE CTA pair, against
                # [x(tau-d) | x(tau) | cond, 1] in <= 512-column units: 256 + 256, then the tails 112 + 112 + 144
                units = []
                for h in (0, 1):
                    cks = []
                    for (c0, n) in chunks(R):
                        for tap, (xa, sh) in enumerate(((1, sh0), (2, 0))):
                            cks.append(dict(x_act=xa, x_row=c0, n_valid=n, shift=sh, out=v[keys[h]], out_off=c0 * 2 + tap,
                                            out_rs=2 * R, out_cs=2))
                    for (c0, n) in chunks(Cc + 1):
                        cks.append(dict(x_act=3, x_row=c0, n_valid=n, shift=0, out=v["dpb"],
                                        out_off=h * D * (Cc + 1) + c0, out_rs=Cc + 1, out_cs=1))
                    units += pack_wide_units(0, h * D, D, lo4, T0, cks)
                launches += build_wgradw(acts, units, B, self.err, tag=f"wgrad1.{l}")
            if wide and self.wgrad16:
                # dWs, dWr: M = the D channels of z16, columns = [g_skp16 | gx16 of the layer above] (both scaled); outputs
                # transposed like the TF32 units below.  K range [lead_l, T0): g_skp16 is zero below RF.
                acts16 = [act16_of(self.z16[l]), act16_of(bw["g_skp16"])]
                cks = [dict(x_act=1, x_row=c0, n_valid=n, shift=0, out=v["dil_skp.weight"], out_off=c0 * D, out_rs=1,
                            out_cs=D) for (c0, n) in chunks(S)]
                if not final and g_sig is not None:
                    acts16.append(act16_of(bw["gx16"][(l + 1) % 2]))
                    cks += [dict(x_act=2, x_row=c0, n_valid=n, shift=0, out=v["dil_res.weight"], out_off=c0 * D, out_rs=1,
                                 out_cs=D) for (c0, n) in chunks(R)]
                launches += build_wgradh(acts16, pack_wide_units(0, 0, D, lo, T0, cks), B, bw["gscale"].data_ptr() + 4,
                                         self.err, tag=f"wgrad2.{l}")
                g_sig = gx
                continue
            if wide:
                # dWs, dWr with the roles swapped: M = the D rows of z (one CTA pair), columns = [g_skp | g_sig], so z
                # is staged once per 512 gradient rows; outputs are written transposed (out_rs = 1, out_cs = D).
                # g_skp is zero below RF (PostPlan / caller contract), so both share the K range [lead_l, T0).
                acts2 = [act_of(self.z[l], T0), act_of(g_skp, T0)]
                cks = [dict(x_act=1, x_row=c0, n_valid=n, shift=0, out=v["dil_skp.weight"], out_off=c0 * D, out_rs=1,
                            out_cs=D) for (c0, n) in chunks(S)]
                if not final and g_sig is not None:
                    acts2.append(act_of(g_sig, T0))
                    cks += [dict(x_act=2, x_row=c0, n_valid=n, shift=0, out=v["dil_res.weight"], out_off=c0 * D, out_rs=1,
                                 out_cs=D) for (c0, n) in chunks(R)]
                launches += build_wgradw(acts2, pack_wide_units(0, 0, D, lo4, T0, cks), B, self.err, tag=f"wgrad2.{l}")
                g_sig = gx
                continue

            def g_item(h, i):
                return dict(g_act=0, g_row=h * D + 128 * i, m_valid=min(128, D - 128 * i), t_lo=lo4, t_hi=T0)

            groups = []
            # (wide units of up to 384 X rows are supported by the kernel but measured no faster: they give up the second
            #  accumulator and one pipeline stage)
            for (c0, n) in chunks(R):
                for tap, (xa, sh) in enumerate(((1, sh0), (2, 0))):
                    groups.append([dict(g_item(h, i), x_act=xa, x_row=c0, n_valid=n, shift=sh, out=v[keys[h]],
                                        out_off=(128 * i) * R * 2 + c0 * 2 + tap, out_rs=2 * R, out_cs=2)
                                   for (h, i) in mtiles])
            for (c0, n) in chunks(Cc + 1):
                groups.append([dict(g_item(h, i), x_act=3, x_row=c0, n_valid=n, shift=0, out=v["dpb"],
                                    out_off=(h * D + 128 * i) * (Cc + 1) + c0, out_rs=Cc + 1, out_cs=1)
                               for (h, i) in mtiles])
            items = pair_items(groups)
            launches += build_wgrad(acts, items, B, self.err, tag=f"wgrad1.{l}", pair=True)
            # (4) dWs = sum g_skp z^T (tau >= RF);  dWr = sum g_sig z^T (tau >= lead_l)
            acts = [act_of(g_skp, T0), act_of(self.z[l], T0)]
            groups = []
            for (c0, n) in chunks(D):
                groups.append([dict(g_act=0, x_act=1, g_row=128 * i, x_row=c0, m_valid=min(128, S - 128 * i), n_valid=n,
                                    t_lo=rf4, t_hi=T0, out=v["dil_skp.weight"], out_off=128 * i * D + c0, out_rs=D,
                                    out_cs=1) for i in range((S + 127) // 128)])
            if not final and g_sig is not None:
                acts.append(act_of(g_sig, T0))
                for (c0, n) in chunks(D):
                    groups.append([dict(g_act=2, x_act=1, g_row=128 * i, x_row=c0, m_valid=min(128, R - 128 * i),
                                        n_valid=n, t_lo=lo4, t_hi=T0, out=v["dil_res.weight"], out_off=128 * i * D + c0,
                                        out_rs=D, out_cs=1) for i in range((R + 127) // 128)])
            items = pair_items(groups)
            launches += build_wgrad(acts, items, B, self.err, tag=f"wgrad2.{l}", pair=True)
            g_sig = gx
        bw["launches"] = launches
        bw["gx0"] = g_sig
        self._bwd = bw
        return bw

    def backward(self):
        """Inputs staged by the caller: bwd()['g_skp'] (zero below RF) and, for a stand-alone non-final layer,
        bwd()['g_last'].  Returns (g_x0 (B,R,Tp) valid on [0,T0), g_cond (B,Cc,Tp), per-layer dict of gradient views).
        The returned tensors alias plan-owned buffers: callers clone what they hand to autograd."""
        bw = self.bwd()
        bw["flat"].zero_()
        bw["g_cond"].zero_()
        run_launches(bw["launches"])
        D, Cc = self.D, self.Cc
        grads = []
        for l, v in enumerate(bw["views"]):
            gr = {"conv_signal.weight": v["conv_signal.weight"], "conv_gate.weight": v["conv_gate.weight"],
                  "proj_signal.weight": v["dpb"][:D, :Cc].unsqueeze(2), "proj_gate.weight": v["dpb"][D:, :Cc].unsqueeze(2),
                  "dil_skp.weight": v["dil_skp.weight"]}
            if self.has_bias:
                gr["conv_signal.bias"], gr["conv_gate.bias"] = v["dpb"][:D, Cc], v["dpb"][D:, Cc]
            if "dil_res.weight" in v:
                gr["dil_res.weight"] = v["dil_res.weight"]
            grads.append(gr)
        return bw["gx0"], bw["g_cond"], grads


class PostPlan:
    """Post-net of the decoder on the kernel path (wavenet.py:359-360): logits = post2(relu(post1(relu(skp_sum)))).
    relu(skp_sum) is already in plan.skp (fused into the last GRCC layer's epilogue); post1 fuses bias + ReLU, post2 the
    bias.  The backward pass writes the gradient w.r.t. the pre-ReLU skip sum straight into the stack's g_skp buffer."""

    def __init__(self, plan, pw):
        self.plan = plan
        g, B, S, dev = plan.geom, plan.B, plan.S, plan.device
        w1, w2 = pw["post1.weight"], pw["post2.weight"]
        self.P, self.Q = w1.shape[0], w2.shape[0]
        P, Q = self.P, self.Q
        self.pw = pw
        self.ptrs = tuple(t.data_ptr() for t in pw.values() if t is not None)
        KS, KP, KQ = ceil_to(S, 32), ceil_to(P, 32), ceil_to(Q, 32)
        self.w1 = torch.zeros(P, KS, device=dev)
        self.w2 = torch.zeros(Q, KP, device=dev)
        self.w1t = torch.zeros(S, KP, device=dev)
        self.w2t = torch.zeros(P, KQ, device=dev)
        whole = [(w1.data_ptr(), self.w1.data_ptr(), P, S, S, 1, KS), (w2.data_ptr(), self.w2.data_ptr(), Q, P, P, 1, KP),
                 (w1.data_ptr(), self.w1t.data_ptr(), S, P, 1, S, KP), (w2.data_ptr(), self.w2t.data_ptr(), P, Q, 1, P, KQ)]
        blocks = []
        for (src, dst, ni, nj, si, sj, di) in whole:       # one CTA per table entry: 32-row slices
            for r0 in range(0, ni, 32):
                blocks.append((src + 4 * r0 * si, dst + 4 * r0 * di, min(32, ni - r0), nj, si, sj, di))
        import numpy as np
        dt = np.dtype([("src", "<u8"), ("dst", "<u8"), ("ni", "<i4"), ("nj", "<i4"), ("si", "<i8"), ("sj", "<i8"),
                       ("di", "<i8")])
        self.n_blocks = len(blocks)
        self.block_table = torch.from_numpy(np.array(blocks, dtype=dt).view(np.uint8).reshape(-1).copy()).to(dev)
        T0, Tp, RF = g.T0, g.Tp, g.RF
        rf4 = RF & ~3
        self.h1 = new_buf(B, P, Tp, dev)
        b1, b2 = pw.get("post1.bias"), pw.get("post2.bias")
        tiles = [ntile(c0, n, self.h1[:, c0:], flags=L.F_RELU, bias=b1[c0:] if b1 is not None else None, t_lo=rf4,
                       t_hi=T0, t_zero_lo=RF) for (c0, n) in chunks(P)]
        self.fwd1 = build_tgemm([act_of(plan.skp, T0)], [(0, 0, S, 0)], self.w1, tiles, B, rf4 & ~31, T0, plan.err,
                                tag="post1")
        self._b2 = b2
        self._bwd = None

    def matches(self, pw):
        return tuple(t.data_ptr() for t in pw.values() if t is not None) == self.ptrs

    def forward(self):
        """Returns logits on the absolute time axis: a fresh (B, Q, Tp) tensor, valid on [RF, T0)."""
        plan = self.plan
        g, B = plan.geom, plan.B
        L.check(L.lib().aewn_pack_blocks(C.c_void_p(self.block_table.data_ptr()), C.c_int(self.n_blocks), _stream()),
                "aewn_pack_blocks")
        run_launches(self.fwd1)
        logits = torch.empty(B, self.Q, g.Tp, device=plan.device)
        rf4 = g.RF & ~3
        b2 = self._b2
        tiles = [ntile(c0, n, logits[:, c0:], bias=b2[c0:] if b2 is not None else None, t_lo=rf4, t_hi=g.T0,
                       t_zero_lo=g.RF) for (c0, n) in chunks(self.Q)]
        run_launches(build_tgemm([act_of(self.h1, g.T0)], [(0, 0, self.P, 0)], self.w2, tiles, B, rf4 & ~31, g.T0,
                                 plan.err, tag="post2"))
        return logits

    def bwd(self):
        if self._bwd is not None:
            return self._bwd
        plan = self.plan
        g, B, S, dev = plan.geom, plan.B, plan.S, plan.device
        P, Q, T0, Tp, RF = self.P, self.Q, g.T0, g.Tp, g.RF
        rf4 = RF & ~3
        sb = plan.bwd()
        bw = dict(g_logits=new_buf(B, Q, Tp, dev), g_h1=new_buf(B, P, Tp, dev))
        has_b = self._b2 is not None
        sizes = [("post1.weight", (P, S, 1)), ("post2.weight", (Q, P, 1))]
        if has_b:
            sizes += [("post1.bias", (P,)), ("post2.bias", (Q,))]
        off, lay = 0, {}
        for k, sh in sizes:
            n = 1
            for v in sh:
                n *= v
            lay[k] = (off, sh, n)
            off += ceil_to(n, 4)
        flat = torch.zeros(off, device=dev)
        bw["flat"] = flat
        bw["views"] = {k: flat[o:o + n].view(sh) for k, (o, sh, n) in lay.items()}
        v = bw["views"]
        launches = []
        # g_h1 = (W2^T g_logits) * [h1 > 0]
        tiles = [ntile(c0, n, bw["g_h1"][:, c0:], flags=L.F_MASKPOS, add=self.h1[:, c0:], t_lo=rf4, t_hi=T0, t_zero_lo=RF)
                 for (c0, n) in chunks(P)]
        launches += build_tgemm([act_of(bw["g_logits"], T0)], [(0, 0, Q, 0)], self.w2t, tiles, B, rf4 & ~31, T0, plan.err,
                                tag="post2_dgrad")
        # g_skp = (W1^T g_h1) * [relu(skp) > 0]  -> the stack's skip-gradient buffer (absolute axis, zero below RF)
        tiles = [ntile(c0, n, sb["g_skp"][:, c0:], flags=L.F_MASKPOS, add=plan.skp[:, c0:], t_lo=rf4, t_hi=T0,
                       t_zero_lo=RF) for (c0, n) in chunks(S)]
        launches += build_tgemm([act_of(bw["g_h1"], T0)], [(0, 0, P, 0)], self.w1t, tiles, B, rf4 & ~31, T0, plan.err,
                                tag="post1_dgrad")
        # weight / bias gradients (bias via the all-ones row)
        acts = [act_of(bw["g_logits"], T0), act_of(self.h1, T0), act_of(ones_row(B, Tp, dev), T0), act_of(bw["g_h1"], T0),
                act_of(plan.skp, T0)]
        groups = []
        for (g_act, M, x_act, Nx, wkey, bkey) in ((0, Q, 1, P, "post2.weight", "post2.bias"),
                                                 (3, P, 4, S, "post1.weight", "post1.bias")):
            mt = range((M + 127) // 128)
            for (c0, n) in chunks(Nx):
                groups.append([dict(g_act=g_act, x_act=x_act, g_row=128 * i, x_row=c0, m_valid=min(128, M - 128 * i),
                                    n_valid=n, t_lo=rf4, t_hi=T0, out=v[wkey], out_off=128 * i * Nx + c0, out_rs=Nx,
                                    out_cs=1) for i in mt])
            if has_b:
                groups.append([dict(g_act=g_act, x_act=2, g_row=128 * i, x_row=0, m_valid=min(128, M - 128 * i), n_valid=1,
                                    t_lo=rf4, t_hi=T0, out=v[bkey], out_off=128 * i, out_rs=1, out_cs=1) for i in mt])
        launches += build_wgrad(acts, pair_items(groups), B, plan.err, tag="post_wgrad", pair=True)
        bw["launches"] = launches
        self._bwd = bw
        return bw

    def backward(self, g_quant):
        """g_quant: gradient w.r.t. logits[:, :, RF:T0].  Fills the stack's g_skp buffer; returns the dict of gradient
        views (aliasing plan-owned memory)."""
        g = self.plan.geom
        bw = self.bwd()
        bw["g_logits"][:, :, g.RF:g.T0] = g_quant
        bw["flat"].zero_()
        run_launches(bw["launches"])
        return bw["views"]


_plans = collections.OrderedDict()
MAX_PLANS = int(os.environ.get("AEWN_MAX_PLANS", "4"))


def dgrad16_mode():
    return os.environ.get("AEWN_DGRAD16", "2")


def get_plan(B, R, D, S, Cc, geom, params, device, relu_last):
    """params: list (per layer) of dicts of the live parameter tensors.  Plans are cached per (configuration, parameter
    set) in a small LRU: two same-shaped models, or a train window alternating with an eval window, each keep their
    workspace; the least recently used plan is dropped when a new one is needed (its memory returns to PyTorch's
    caching allocator -- no empty_cache(), which would fail inside a CUDA-graph capture)."""
    key = (B, R, D, S, Cc, geom.key(), str(device), bool(relu_last), StackPlan._ptrs(params), FUSED_FWD,
           dgrad16_mode(), os.environ.get("AEWN_WGRAD16", "1"), os.environ.get("AEWN_GZ16", "1"), ENGINE_MODE, WIDE_WGRAD)
    plan = _plans.get(key)
    if plan is not None:
        _plans.move_to_end(key)
        return plan
    while len(_plans) >= max(1, MAX_PLANS):
        _plans.popitem(last=False)
    plan = _plans[key] = StackPlan(B, R, D, S, Cc, geom, params, device, relu_last)
    return plan


def check_device_errors():
    """Synchronise and raise if any kernel reported a device-side fault (bounded-wait timeout, bad mu-law code)."""
    torch.cuda.synchronize()
    words = [p.err for p in _plans.values()] + list(_err_cache.values())
    for w in words:
        e = int(w.item())
        if e != 0:
            w.zero_()
            raise RuntimeError(f"aewn: device-side error word = {e} "
                               f"({'bounded wait timed out' if e == L.ERR_TIMEOUT else 'an activation (forward) or a scaled gradient (backward; AEWN_DGRAD16=0 selects the TF32 engines) outside the fp16 operand range' if e == L.ERR_RANGE else 'invalid input'})")


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


class _TapConvTransposeFn(torch.autograd.Function):
    """ConvTranspose1d(C_in -> C_out, kernel f, stride s, padding p) on the tcgen05 engines (wavenet.py:154-155 uses
    p = f - s).  Polyphase form: with o' = o + p, r = o' mod s, q = o' // s,

        y[b, n, o] = bias[n] + sum_m sum_c W[c, n, r + s*m] * x[b, c, q - m]

    i.e. one time-major GEMM per output phase r over the M = ceil(f/s) tap-shifted copies of x (no zero-stuffed input,
    no wasted MACs).  Backward: g_x = strided correlation of g_y (again one GEMM per phase, accumulated), dW through the
    weight-gradient engine, d(bias) = sum g_y."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, padding):
        B, Ci, Lx = x.shape
        _, Co, f = weight.shape
        s, p = stride, padding
        Lo = (Lx - 1) * s - 2 * p + f
        dev = x.device
        err = err_word(dev)
        kc = ceil_to(Ci, 32)
        xd, wd = x.detach(), weight.detach()
        y = torch.empty(B, Co, Lo, device=dev)
        phases = []
        with torch.no_grad():
            for r in range(s):
                js = list(range(r, f, s))
                # outputs of this phase: o = s*q + r - p, 0 <= o < Lo
                q0 = max(0, -((r - p) // s))
                q1 = (Lo - 1 - (r - p)) // s          # inclusive
                if q1 < q0 or not js:
                    continue
                nq = q1 + 1
                Tp = ceil_to(nq + 4, 32)
                shifted = []
                for m in range(len(js)):              # xm[q] = x[q - m]
                    sb = new_buf(B, Ci, Tp, dev)
                    hi = min(nq, Lx + m)
                    if hi > m:
                        sb[:, :, m:hi] = xd[:, :, :hi - m]
                    shifted.append(sb)
                wp = torch.cat([_pad_k(wd[:, :, j].t(), kc) for j in js], 1).contiguous()      # [Co][M*kc]
                yr = new_buf(B, Co, Tp, dev)
                for g0 in range(0, len(js), L.MAX_SEGS):        # at most MAX_SEGS taps per launch; later groups accumulate
                    ms = list(range(g0, min(len(js), g0 + L.MAX_SEGS)))
                    tiles = [ntile(c0, n, yr[:, c0:], flags=L.F_ACCUM if g0 else 0,
                                   bias=bias.detach()[c0:] if (bias is not None and g0 == 0) else None, t_lo=0, t_hi=nq)
                             for (c0, n) in chunks(Co)]
                    tgemm([act_of(shifted[m], nq) for m in ms], [(i, 0, Ci, m * kc) for i, m in enumerate(ms)], wp, tiles,
                          B, 0, nq, err, tag="tconv_fwd")
                o0 = s * q0 + r - p
                y[:, :, o0::s] = yr[:, :, q0:nq]
                phases.append((r, js, q0, nq, shifted))
        ctx.save_for_backward(weight)
        ctx.cfg = (B, Ci, Lx, Co, f, s, p, Lo, bias is not None)
        ctx.phases = phases
        return y

    @staticmethod
    def backward(ctx, g):
        (weight,) = ctx.saved_tensors
        B, Ci, Lx, Co, f, s, p, Lo, has_bias = ctx.cfg
        dev = g.device
        err = err_word(dev)
        wd = weight.detach()
        kn = ceil_to(Co, 32)
        dw = torch.zeros_like(weight)
        Tx = ceil_to(Lx + 4, 32)
        gx = new_buf(B, Ci, Tx, dev)
        first = True
        for (r, js, q0, nq, shifted) in ctx.phases:
            Tp = shifted[0].shape[2]
            # g_yr[q] = g_y[s*q + r - p]
            gyr = new_buf(B, Co, Tp, dev)
            gyr[:, :, q0:nq] = g[:, :, s * q0 + r - p::s]
            # data gradient: g_x[i] += sum_m W[:, :, r + s*m] g_yr[i + m]
            M = len(js)
            adv = []
            for m in range(M):                        # gm[i] = g_yr[i + m]
                if m == 0:
                    adv.append(gyr)
                else:
                    sb = new_buf(B, Co, Tp, dev)
                    if nq > m:
                        sb[:, :, :nq - m] = gyr[:, :, m:nq]
                    adv.append(sb)
            wt = torch.cat([_pad_k(wd[:, :, j], kn) for j in js], 1).contiguous()              # [Ci][M*kn]
            for g0 in range(0, M, L.MAX_SEGS):
                ms = list(range(g0, min(M, g0 + L.MAX_SEGS)))
                tiles = [ntile(c0, n, gx[:, c0:], flags=0 if first else L.F_ACCUM, t_lo=0, t_hi=Lx)
                         for (c0, n) in chunks(Ci)]
                tgemm([act_of(adv[m], min(nq, adv[m].shape[2])) for m in ms], [(i, 0, Co, m * kn) for i, m in enumerate(ms)],
                      wt, tiles, B, 0, Lx, err, tag="tconv_dgrad")
                first = False
            # weight gradient: dW[c, n, r + s*m] = sum_{b,q} x[c, q - m] * g_yr[n, q]
            for g0 in range(0, M, L.WGRAD_MAX_ACTS - 1):      # the engine takes 6 operand tensors: g_yr + 5 taps
                ms = list(range(g0, min(M, g0 + L.WGRAD_MAX_ACTS - 1)))
                acts = [act_of(gyr, nq)] + [act_of(shifted[m], nq) for m in ms]
                items = []
                for i in range((Co + 127) // 128):
                    for k, m in enumerate(ms):
                        for (c0, n) in chunks(Ci):
                            items.append(dict(g_act=0, x_act=1 + k, g_row=128 * i, x_row=c0,
                                              m_valid=min(128, Co - 128 * i), n_valid=n, t_lo=0, t_hi=nq, out=dw,
                                              out_off=c0 * Co * f + (128 * i) * f + js[m], out_rs=f, out_cs=Co * f))
                wgrad(acts, items, B, err, tag="tconv_wgrad")
        db = g.sum(dim=(0, 2)) if has_bias else None
        return gx[:, :, :Lx].clone(), dw, db, None, None


class _Conv1x1F32Fn(torch.autograd.Function):
    """Bias-free 1x1 convolution in EXACT fp32 (aewn_conv1x1_f32: sequential fmaf, no tensor cores) with its two
    gradients.  Used for the bottleneck projection in front of the nearest-code search (vqema_bn.py:131, vq_bn.py:35):
    the code indices are index work, and TF32 operand rounding there moves codes across near-ties."""

    @staticmethod
    def forward(ctx, x, weight):
        B, K, T = x.shape
        N = weight.shape[0]
        xd = x.detach()
        if xd.stride(2) != 1:
            xd = xd.contiguous()
        w = weight.detach().reshape(N, K).contiguous()
        out = torch.empty(B, N, T, device=x.device)
        L.check(L.lib().aewn_conv1x1_f32(
            C.c_void_p(xd.data_ptr()), C.c_longlong(xd.stride(0)), C.c_longlong(xd.stride(1)), C.c_void_p(w.data_ptr()),
            C.c_longlong(K), C.c_longlong(1), C.c_void_p(out.data_ptr()), C.c_longlong(out.stride(0)),
            C.c_longlong(out.stride(1)), C.c_int(B), C.c_int(N), C.c_int(K), C.c_int(T), _stream()), "aewn_conv1x1_f32")
        ctx.save_for_backward(xd, w)
        ctx.wshape = weight.shape
        return out

    @staticmethod
    def backward(ctx, g):
        xd, w = ctx.saved_tensors
        B, K, T = xd.shape
        N = w.shape[0]
        gc = g.contiguous()
        gx = dw = None
        if ctx.needs_input_grad[0]:
            gx = torch.empty(B, K, T, device=g.device)
            L.check(L.lib().aewn_conv1x1_f32(
                C.c_void_p(gc.data_ptr()), C.c_longlong(gc.stride(0)), C.c_longlong(gc.stride(1)),
                C.c_void_p(w.data_ptr()), C.c_longlong(1), C.c_longlong(K), C.c_void_p(gx.data_ptr()),
                C.c_longlong(gx.stride(0)), C.c_longlong(gx.stride(1)), C.c_int(B), C.c_int(K), C.c_int(N), C.c_int(T),
                _stream()), "aewn_conv1x1_f32")
        if ctx.needs_input_grad[1]:
            dw = torch.empty(N, K, device=g.device)
            L.check(L.lib().aewn_conv1x1_wgrad_f32(
                C.c_void_p(gc.data_ptr()), C.c_longlong(gc.stride(0)), C.c_longlong(gc.stride(1)),
                C.c_void_p(xd.data_ptr()), C.c_longlong(xd.stride(0)), C.c_longlong(xd.stride(1)),
                C.c_void_p(dw.data_ptr()), C.c_int(B), C.c_int(N), C.c_int(K), C.c_int(T), _stream()),
                "aewn_conv1x1_wgrad_f32")
            dw = dw.reshape(ctx.wshape)
        return gx, dw


def conv1x1_f32(x, weight):
    return _Conv1x1F32Fn.apply(x, weight)


# Conditioning front-end (lc_conv + 4 transposed convs, ~0.5 % of the decoder FLOPs): "torch" = cuDNN through the
# nn.Conv1d / nn.ConvTranspose1d parameter containers, "kernels" = the polyphase GEMMs above.  The kernel path is
# parity-tested but its per-phase staging copies make the whole step ~10 % slower than cuDNN here (measured 39.5 vs
# 35.7 ms on cfg2), so it is opt-in until the staging moves into the producing epilogues.
FRONTEND = "torch"


def set_frontend(mode):
    global FRONTEND
    if mode not in ("torch", "kernels"):
        raise ValueError("frontend must be 'torch' or 'kernels'")
    FRONTEND = mode


def tap_conv_transpose(x, weight, bias=None, stride=1, padding=0):
    return _TapConvTransposeFn.apply(x, weight, bias, stride, padding)
