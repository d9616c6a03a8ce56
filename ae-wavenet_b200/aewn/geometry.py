"""Receptive-field geometry for chains of strided / transposed 1-D windows.

Own restatement of the integer algebra the reference keeps in vconv.py (GridRange :8-54, VirtualConv :57-277,
input_range/output_range :282-344, output_offsets :347-361, compute_inputs :399-428).  The drop-in modules prefer the
caller's own ``vconv`` module when it is importable (so ``mfcc_inverter.py`` can keep calling ``vconv.compute_inputs``
on our ``.vc`` objects); this module is what they fall back to when the reference tree is absent (GPU box, bench).
``tests/test_geometry.py`` checks it against the reference on random chains and against tests/golden/geometry.json.

Model: every tensor lives on a global integer coordinate line.  A tensor is the set of positions
``first, first+gs, ..., last`` ("full") plus a marked sub-interval ("sub") on the same lattice.  A window op maps
(full, sub, gs) of its input to those of its output and back.
"""
from __future__ import annotations

__all__ = ["GridRange", "VirtualConv", "input_range", "output_range", "output_offsets", "tensor_slice",
           "max_spacing", "compute_inputs"]


class GridRange:
    """Half-open coordinate ranges ``full=[f0,f1)`` and ``sub=[s0,s1)`` holding lattice points spaced ``gs``."""

    def __init__(self, full, sub, gs):
        for v in (full[0], full[1], sub[0], sub[1]):
            if not isinstance(v, int):
                raise ValueError("Only integer ranges supported")
        self.full = [full[0], full[1]]
        self.sub = [sub[0], sub[1]]
        self.gs = gs

    @staticmethod
    def _count(lo, hi, gs):
        return (hi - lo - 1) // gs + 1

    def sub_length(self):
        return self._count(self.sub[0], self.sub[1], self.gs)

    def full_length(self):
        return self._count(self.full[0], self.full[1], self.gs)

    def valid(self):
        (f0, f1), (s0, s1), g = self.full, self.sub, self.gs
        if g < 1 or not (f0 <= s0 < s1 <= f1):
            return False
        return f0 % g == (f1 - 1) % g == s0 % g == (s1 - 1) % g

    def __repr__(self):
        return (f"GridRange(full=[{self.full[0]},{self.full[1]}) sub=[{self.sub[0]},{self.sub[1]}) gs={self.gs} "
                f"n_sub={self.sub_length()} n_full={self.full_length()})")


def _snap_up(x, anchor, g):
    """smallest y >= x with (anchor - y) % g == 0"""
    return x + (anchor - x) % g


def _snap_down(x, anchor, g):
    """largest y <= x with (y - anchor) % g == 0"""
    return x - (x - anchor) % g


class VirtualConv:
    """One scanning-window op (conv, strided conv, transposed conv, STFT frame) in a doubly linked chain."""

    def __init__(self, filter_info, padding=(0, 0), stride=1, is_downsample=True, do_trim_input=False, name=None,
                 parent=None):
        self.parent = parent
        self.child = None
        if parent is not None:
            parent.child = self
        self.l_pad, self.r_pad = padding
        self.stride = stride
        self.is_downsample = is_downsample
        self.do_trim_input = do_trim_input
        self.name = name
        self.input_gr = None
        self.input_trim = None
        if isinstance(filter_info, tuple):
            self.l_wing_sz, self.r_wing_sz = filter_info
        elif isinstance(filter_info, int):
            self.l_wing_sz = (filter_info - 1) // 2
            self.r_wing_sz = (filter_info - 1) - self.l_wing_sz
        else:
            raise RuntimeError("filter_info must be either a 2-tuple of (l_wing_sz, r_wing_sz) or an integer of "
                               "filter_sz")
        if self.l_pad > self.l_wing_sz or self.r_pad > self.r_wing_sz:
            raise RuntimeError("Filter wing sizes cannot be less than the respective padding")

    # -- small accessors -------------------------------------------------------------------------------------
    def in_len(self):
        return self.input_gr.sub_length() if self.input_gr else None

    def out_len(self):
        return self.child.in_len() if self.child else None

    def filter_size(self):
        return self.l_wing_sz + 1 + self.r_wing_sz

    def get_index_trim(self):
        n = self.input_gr.sub_length()
        if self.input_trim is None:
            return [0, n]
        return [self.input_trim[0], n - self.input_trim[1]]

    def __repr__(self):
        ratio = f"{self.stride}/1" if self.is_downsample else f"1/{self.stride}"
        sub = self.input_gr.sub if self.input_gr is not None else (None, None)
        return (f"[{self.l_wing_sz}^{self.r_wing_sz}, {ratio}, {self.l_pad}--{self.r_pad}, "
                f"{'T' if self.do_trim_input else '-'}, [sub: [{sub[0]}, {sub[1]}), in: {self.in_len()}, "
                f"trim: {self.input_trim}] \"{self.name}\"]")

    # -- one-step maps on inclusive (first, last) intervals -----------------------------------------------------
    def _output_range(self, full_in, sub_in, gs_in):
        fb, fe = full_in
        sb, se = sub_in
        if self.is_downsample:
            gs_out = gs_in * self.stride
            unit = gs_in
            fb, fe = fb - self.l_pad * unit, fe + self.r_pad * unit
            need = (self.l_wing_sz + self.r_wing_sz) * unit
            if fe - fb < need or se - sb < need:
                return None
            ofb = fb + self.l_wing_sz * unit
            ofe = _snap_down(fe - self.r_wing_sz * unit, ofb, gs_out)
            osb = _snap_up(sb + self.l_wing_sz * unit, ofe, gs_out)
            ose = _snap_down(se - self.r_wing_sz * unit, ofb, gs_out)
            if ose - osb < 0:
                return None
        else:
            assert gs_in % self.stride == 0
            gs_out = gs_in // self.stride
            unit = gs_out
            pfb, pfe = fb - self.l_pad * unit, fe + self.r_pad * unit
            reach = (self.stride - 1) * unit
            sb = pfb if sb == fb else sb - reach
            se = pfe if se == fe else se + reach
            need = (self.l_wing_sz + self.r_wing_sz) * unit
            if pfe - pfb < need or se - sb < need:
                return None
            ofb, ofe = pfb + self.l_wing_sz * unit, pfe - self.r_wing_sz * unit
            osb, ose = sb + self.l_wing_sz * unit, se - self.r_wing_sz * unit
        return (ofb, ofe), (osb, ose), gs_out

    def _input_range(self, full_out, sub_out, gs_out):
        fb, fe = full_out
        sb, se = sub_out
        if self.is_downsample:
            assert gs_out % self.stride == 0
            gs_in = gs_out // self.stride
            unit = gs_in
            lw, rw = self.l_wing_sz * unit, self.r_wing_sz * unit
            pfb, pfe = fb - lw, fe + rw
            psb, pse = sb - lw, se + rw
            if pfe - pfb < lw + rw or pse - psb < lw + rw:
                return None
            ifb, ife = pfb + self.l_pad * unit, pfe - self.r_pad * unit
            isb, ise = max(psb, ifb), min(pse, ife)
        else:
            gs_in = gs_out * self.stride
            unit = gs_out
            lw, rw = self.l_wing_sz * unit, self.r_wing_sz * unit
            afb, afe = fb - lw, fe + rw
            asb, ase = sb - lw, se + rw
            ifb = afb + self.l_pad * unit
            ife = _snap_up(afe - self.r_pad * unit, ifb, gs_in)
            assert asb <= ife and ifb <= ase
            isb = _snap_up(asb, ife, gs_in)
            ise = _snap_down(ase, ifb, gs_in)
            if ise - isb < 0:
                return None
        return (ifb, ife), (isb, ise), gs_in

    def _output_offsets(self):
        if self.l_pad != 0 or self.r_pad != 0 or self.stride != 1:
            raise RuntimeError("Can only call output_offset with no padding and unit stride")
        return self.l_wing_sz, -self.r_wing_sz


def _walk(first, last, attr):
    vc = first
    while True:
        yield vc
        if vc is last:
            return
        vc = getattr(vc, attr)


def input_range(source, dest, out):
    """GridRange of the input of ``source`` needed for output ``out`` of ``dest`` (source is an ancestor of dest)."""
    full, sub, gs = (out.full[0], out.full[1] - 1), (out.sub[0], out.sub[1] - 1), out.gs
    for vc in _walk(dest, source, "parent"):
        res = vc._input_range(full, sub, gs)
        if res is None:
            raise RuntimeError("empty input range")
        full, sub, gs = res
    return GridRange((full[0], full[1] + 1), (sub[0], sub[1] + 1), gs)


def output_range(source, dest, gin):
    """GridRange of the output of ``dest`` produced from input ``gin`` of ``source``."""
    full, sub, gs = (gin.full[0], gin.full[1] - 1), (gin.sub[0], gin.sub[1] - 1), gin.gs
    for vc in _walk(source, dest, "child"):
        res = vc._output_range(full, sub, gs)
        if res is None:
            raise RuntimeError("empty output range")
        full, sub, gs = res
    return GridRange((full[0], full[1] + 1), (sub[0], sub[1] + 1), gs)


def output_offsets(source, dest):
    """(left, right) index offsets from the input of ``source`` to the output of ``dest`` (unit-stride chains)."""
    lo = ro = 0
    for vc in _walk(source, dest, "child"):
        a, b = vc._output_offsets()
        lo += a
        ro += b
    return lo, ro


def tensor_slice(ref_gcoord, subrange_gcoord):
    r0, r1 = ref_gcoord.sub
    g = ref_gcoord.gs
    t0, t1 = subrange_gcoord
    assert r0 <= t0 and t1 <= r1
    b, e = t0 - r0, t1 - r0
    assert b % g == 0 and (e - 1) % g == 0
    return b // g, (e - 1) // g + 1


def max_spacing(source, dest, initial_gs):
    gs = peak = initial_gs
    for vc in _walk(source, dest, "child"):
        if vc.is_downsample:
            gs *= vc.stride
        else:
            assert gs % vc.stride == 0
            gs //= vc.stride
        peak = max(peak, gs)
    return peak


def compute_inputs(end_vc, end_gr):
    """Set ``input_gr`` (and ``input_trim`` on trimming stages) for ``end_vc`` and all its ancestors so that the
    chain yields at least ``end_gr``; returns the output range actually produced."""
    vc, gr = end_vc, end_gr
    while True:
        gr = input_range(vc, vc, gr)
        vc.input_gr = gr
        if vc.parent is None:
            break
        vc = vc.parent
    while True:
        gr = output_range(vc, vc, vc.input_gr)
        if vc.child is None:
            break
        vc = vc.child
        if vc.do_trim_input:
            vc.input_trim = [vc.input_gr.sub[0] - gr.sub[0], gr.sub[1] - vc.input_gr.sub[1]]
        else:
            vc.input_gr = gr
    return gr
