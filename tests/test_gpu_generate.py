"""GPU parity of the persistent incremental sampler (csrc/gen.cu) behind WaveNet.forward_test (wavenet.py:367-531).

The kernel computes in fp32 FMA (no TF32); against a float64 oracle fed the same conditioning the logits agree to
~1e-6 of the logit scale, the tests assert 2e-5.  Draws are inverse-CDF draws on supplied uniforms, so a whole generated sequence is reproducible and
is compared sample for sample with the reference's own run (golden) and with the oracle."""
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import torch_oracle as orc

pytestmark = pytest.mark.gpu


class HP(dict):
    __getattr__ = dict.__getitem__


def build(hp, W, state_dict=None, seed=2507):
    import aewn
    from aewn import geometry as vc
    torch.manual_seed(seed)
    parent = vc.VirtualConv(filter_info=1, stride=320, parent=None, name="LC-grid")
    wn = aewn.WaveNet(HP(hp), parent_vc=parent)
    vc.compute_inputs(wn.vc["end_grcc"], vc.GridRange((0, 10 ** 7), (0, W), 1))
    wn.trim_ups_out = torch.tensor([0, wn.vc["beg_grcc"].in_len()], dtype=torch.long)
    wn.post_init(W)
    if state_dict is not None:
        wn.load_state_dict(state_dict, strict=True)
    return wn.cuda().eval()


def place_uniforms(u_steps, rf1, T):
    """(steps, n_rep) step-major uniforms -> (n_rep, T) with the draw for index i at column i."""
    n = u_steps.shape[0]
    u = torch.zeros(u_steps.shape[1], T)
    u[:, rf1:rf1 + n] = u_steps.t()
    return u


def test_forward_test_reproduces_reference_run(golden_dir):
    from aewn import ops
    g = torch.load(os.path.join(golden_dir, "forward_test.pt"))
    wn = build(g["hp"], g["W"], g["state_dict"])
    assert list(wn.wav_cond_offset) == list(g["geo"]["wav_cond_offset"]) and wn.base_global_rf == g["base_global_rf"]
    rf1, T = g["base_global_rf"], g["out"].shape[1]
    wn.set_n_replicas(g["n_rep"])
    wn.gen_uniforms = place_uniforms(g["uniforms"], rf1, T)
    wn.keep_gen_logits = True
    out = wn(g["wav"].cuda(), g["lc"].cuda(), g["spk"].cuda(), g["jit"].cuda())
    ops.check_device_errors()
    assert out.shape == g["out"].shape and out.dtype == g["out"].dtype
    n = g["probs"].shape[0]
    probs = F.softmax(wn.gen_logits[:, rf1:rf1 + n].cpu(), dim=-1).permute(1, 0, 2)
    # the first step's distribution depends on the input only -- a clean numerics check
    assert float((probs[0] - g["probs"][0].float()).abs().max()) < 1e-3
    mism = (out.cpu() != g["out"])
    assert not bool(mism.any()), f"first mismatch at column {int(mism.any(0).nonzero()[0])}"
    assert float((probs - g["probs"].float()).abs().max()) < 1e-3


def test_sliced_launches_continue_the_same_sequence(golden_dir):
    """History rings and codes live in global memory: cutting the run into many launches must not change a sample."""
    from aewn import generate
    g = torch.load(os.path.join(golden_dir, "forward_test.pt"))
    wn = build(g["hp"], g["W"], g["state_dict"])
    rf1, T = g["base_global_rf"], g["out"].shape[1]
    wn.set_n_replicas(2)
    wn.gen_uniforms = place_uniforms(g["uniforms"], rf1, T)
    args = (g["wav"].cuda(), g["lc"].cuda(), g["spk"].cuda(), g["jit"].cuda())
    whole = wn(*args)
    saved = generate.SLICE_STEPS
    try:
        for steps in (37, 1):
            generate.SLICE_STEPS = steps
            plan = generate.get_plan(wn, 2)
            off0 = int(wn.wav_cond_offset[0])
            cond = wn.conditioning(args[1], args[2], args[3], trim=False)[0]
            out = plan.generate(args[0][0, off0:].long(), cond, rf1, uniforms=wn.gen_uniforms, slice_steps=steps)
            assert torch.equal(out.float(), whole[1:]), steps
            if steps == 1:
                break
    finally:
        generate.SLICE_STEPS = saved


ARCH_BASIC = dict(filter_sz=2, n_lc_out=128, lc_upsample_strides=[5, 4, 4, 4], lc_upsample_filt_sizes=[25, 16, 16, 16],
                  n_res=368, n_dil=256, n_skp=256, n_post=256, n_quant=256, n_blocks=2, n_block_layers=10,
                  n_global_embed=10, n_speakers=40, bias=True, n_lc_in=64)


@pytest.mark.parametrize("n_rep,n_res", [(1, 368), (3, 368), (2, 512)])
def test_arch_basic_generation_is_teacher_forced_consistent(n_rep, n_res):
    """par/arch.basic.json widths (and the 512-channel stress width: two float4 columns per thread), cluster of 16:
    for draws picked along the sequence, recompute the logits on CPU from the generated history (oracle, fp32) and
    check (a) the kernel's logits, (b) that the drawn code is the inverse-CDF draw of those logits."""
    from aewn import generate, ops
    hp = dict(ARCH_BASIC, n_res=n_res)
    wn = build(hp, 1024, seed=11)
    rf1 = wn.base_global_rf
    assert rf1 == 2047
    gen = torch.Generator().manual_seed(5)
    lc = torch.randn(1, hp["n_lc_in"], 14, generator=gen)
    spk = torch.randint(0, hp["n_speakers"], (1,), generator=gen)
    jit = torch.arange(14).unsqueeze(0)
    off0 = int(wn.wav_cond_offset[0])
    with torch.no_grad():
        n_ts = wn.conditioning(lc.cuda(), spk.cuda(), jit.cuda(), trim=False).shape[2]
    assert n_ts > rf1 + 200
    T = n_ts + 40
    wav = torch.randint(0, 256, (1, off0 + T), generator=gen).float()
    u = torch.rand(n_rep, T, generator=gen)
    wn.set_n_replicas(n_rep)
    wn.gen_uniforms = u
    wn.keep_gen_logits = True
    out = wn(wav.cuda(), lc.cuda(), spk.cuda(), jit.cuda())
    ops.check_device_errors()
    plan = generate.get_plan(wn, n_rep)
    assert plan.cluster == 16
    assert out.shape == (n_rep + 1, T)
    out = out.cpu().long()
    assert torch.equal(out[0], wav[0, off0:].long())
    assert torch.equal(out[1:, :rf1], out[:1, :rf1].expand(n_rep, -1))
    assert torch.equal(out[1:, n_ts:], out[:1, n_ts:].expand(n_rep, -1))
    assert int((out[1:, rf1:n_ts] != out[:1, rf1:n_ts]).sum()) > 0.9 * n_rep * (n_ts - rf1)   # really generated
    if n_rep > 1:
        assert not torch.equal(out[1], out[2])
    # float64 oracle: separates the kernel's fp32 rounding from the fp32 CPU oracle's own (both ~1e-4 of the logit
    # scale after 20 layers of 875-term dot products)
    sd = {k: (v.detach().cpu().double() if v.is_floating_point() else v.detach().cpu()) for k, v in wn.state_dict().items()}
    # conditioning comes from the module's own front-end (cuDNN, TF32 by default -- like the reference on a GPU), so
    # that the comparison isolates the sampler kernel
    with torch.no_grad():
        cond = wn.conditioning(lc.cuda(), spk.cuda(), jit.cuda(), trim=False).cpu().double()
    logits = wn.gen_logits.cpu()
    for cur in (rf1, rf1 + 1, rf1 + 2, rf1 + 77, n_ts - 1):
        ref = orc.stack_window_logits(sd, hp, out[1:, cur - rf1:cur], cond[:, :, cur - rf1:cur].expand(n_rep, -1, -1))
        got = logits[:, cur]
        err = float((got.double() - ref).abs().max()) / float(ref.abs().max())
        assert err < 2e-5, (cur, err)
        draw = orc.inverse_cdf_draw(F.softmax(got, -1), u[:, cur])
        cdf = F.softmax(got.double(), -1).cumsum(-1)
        for r in range(n_rep):
            k = int(out[1 + r, cur])
            if k != int(draw[r]):       # only acceptable when u sits on a bin edge to rounding
                edge = float(cdf[r, min(k, int(draw[r]))])
                assert abs(edge - float(u[r, cur])) < 1e-5, (cur, r, k, int(draw[r]))


def test_invalid_code_is_reported():
    from aewn import ops
    g_hp = dict(ARCH_BASIC, n_res=64, n_dil=32, n_skp=32, n_post=32, n_lc_out=16, n_blocks=1, n_block_layers=3)
    wn = build(g_hp, 64, seed=3)
    lc = torch.randn(1, g_hp["n_lc_in"], 8)
    off0 = int(wn.wav_cond_offset[0])
    wav = torch.full((1, off0 + 2000), 300.0)        # 300 is not a mu-law code; the reference raises in F.one_hot
    with pytest.raises(RuntimeError, match="invalid input"):
        wn(wav.cuda(), lc.cuda(), torch.zeros(1, dtype=torch.long).cuda(), torch.arange(8).unsqueeze(0).cuda())
