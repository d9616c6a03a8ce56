"""The in-package geometry (aewn.geometry) must reproduce the reference's vconv numbers."""
import json
import os
import random

import pytest
import torch


class HP(dict):
    __getattr__ = dict.__getitem__


ARCH_BASIC = dict(filter_sz=2, n_lc_out=128, lc_upsample_strides=[5, 4, 4, 4], lc_upsample_filt_sizes=[25, 16, 16, 16],
                  n_res=368, n_dil=256, n_skp=256, n_post=256, n_quant=256, n_blocks=2, n_block_layers=10,
                  n_global_embed=10, n_speakers=40, bias=True, n_lc_in=64)


def standalone(hp, W):
    import aewn
    from aewn import geometry as vc
    parent = vc.VirtualConv(filter_info=1, stride=320, parent=None, name="LC-grid")
    with torch.device("meta"):
        wn = aewn.WaveNet(HP(hp), parent_vc=parent)
    end_gr = vc.GridRange((0, 10 ** 7), (0, W), 1)
    vc.compute_inputs(wn.vc["end_grcc"], end_gr)
    beg = wn.vc["beg_grcc"]
    geo = dict(wav_len=parent.in_len(), lc_len=parent.child.in_len(), dec_in_len=beg.in_len(),
               trim_dec_in=[beg.input_gr.sub[0] - parent.input_gr.sub[0], beg.input_gr.sub[1] - parent.input_gr.sub[0]],
               trim_dec_out=[end_gr.sub[0] - parent.input_gr.sub[0], end_gr.sub[1] - parent.input_gr.sub[0]])
    wn.trim_ups_out = torch.tensor([0, beg.in_len()], dtype=torch.long)
    one_gr = vc.GridRange((0, int(1e12)), (0, 1), 1)
    win_gr = vc.GridRange((0, int(1e12)), (0, W), 1)
    vc.compute_inputs(wn.vc["end_grcc"], win_gr)
    di, wi = wn.vc["beg_grcc"].input_gr, wn.vc["beg"].parent.input_gr
    geo["wav_cond_offset"] = [int(di.sub[0] - wi.sub[0]), int(di.sub[1] - wi.sub[0])]
    vc.compute_inputs(wn.vc["end_grcc"], one_gr)
    leads = []
    for layer in wn.conv_layers:
        cl, _ = vc.output_offsets(wn.vc["beg_grcc"], layer.vc)
        sl = 0 if layer.vc is wn.vc["end_grcc"] else vc.output_offsets(layer.vc.child, wn.vc["end_grcc"])[0]
        leads.append([cl, sl, layer.vc.l_wing_sz, 0])
    geo["leads"] = leads
    return geo


@pytest.mark.parametrize("name,W", [("cfg2_basic_W16384", 16384), ("basic_W1024", 1024)])
def test_decoder_geometry_matches_reference_golden(golden_dir, name, W):
    ref = json.load(open(os.path.join(golden_dir, "geometry.json")))[name]
    assert standalone(ARCH_BASIC, W) == ref


@pytest.mark.needs_reference
def test_random_chains_match_reference_vconv():
    import sys
    sys.path.insert(0, "/root/reference")
    try:
        import vconv as rv
    finally:
        sys.path.remove("/root/reference")
    from aewn import geometry as mv
    rnd = random.Random(7)

    def build(mod, spec):
        vc, out = None, []
        for (kind, f, s, pad, trim) in spec:
            vc = mod.VirtualConv(filter_info=f, stride=s, padding=pad, is_downsample=(kind == "d"), do_trim_input=trim,
                                 parent=vc, name=str(len(out)))
            out.append(vc)
        return out

    compared = 0
    for _ in range(400):
        spec = [("d", 400, 160, (0, 0), False)]
        for _i in range(rnd.randint(0, 4)):
            spec.append(("d", rnd.choice([1, 3, 4, (2, 0), (5, 1)]), rnd.choice([1, 1, 2]), (0, 0), False))
        for f, s in [(25, 5), (16, 4), (16, 4), (16, 2)][:rnd.randint(0, 4)]:
            spec.append(("u", f, s, (s - 1, s - 1), False))
        spec.append(("d", 1, 1, (0, 0), True))
        for i in range(rnd.randint(1, 6)):
            spec.append(("d", (2 ** i, 0), 1, (0, 0), False))
        W = rnd.randint(1, 3000)
        res = []
        for mod in (rv, mv):
            try:
                ch = build(mod, spec)
                gr = mod.compute_inputs(ch[-1], mod.GridRange((0, 100000), (0, W), 1))
                res.append(([(c.input_gr.full, c.input_gr.sub, c.input_gr.gs, c.input_trim, c.in_len()) for c in ch],
                            (gr.full, gr.sub, gr.gs)))
            except (RuntimeError, AssertionError) as e:
                res.append(("EXC", type(e).__name__))
        assert res[0] == res[1], (spec, W)
        compared += res[0][0] != "EXC"
    assert compared >= 40      # the rest raised the same exception in both implementations
