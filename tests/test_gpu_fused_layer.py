"""The fused dilation-layer kernel (aewn_grcc_fwd, one launch per layer: wavenet.py:91-111) through ops.StackPlan.

Tolerances.  The kernel feeds the tensor cores fp16 copies of x, cond, z and of the weights (round-to-nearest, 10-bit
mantissa), accumulates in fp32 and keeps the residual stream in fp32.  The first test compares it with a float64
reference that applies EXACTLY those operand roundings, so what is left is fp32 accumulation order, the ex2/rcp
approximations of tanh / sigmoid (~1e-7 absolute) and the occasional 1-ulp flip of an fp16 rounding: 2e-4 of the
tensor's max-abs, i.e. 25x tighter than the TF32 envelope (5e-3) the un-rounded comparisons use.  A wrong time step at
a lead boundary, a mis-zeroed margin column or a swapped channel block is an O(1) error under this bound."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def h(t):
    """fp16 operand rounding (round to nearest even), result kept in float64."""
    return t.float().half().double()


def make_params(R, D, S, Cc, dils, final_last, gen, dev):
    params = []
    for l, _ in enumerate(dils):
        final = final_last and l == len(dils) - 1
        p = {"conv_signal.weight": 0.06 * torch.randn(D, R, 2, generator=gen),
             "conv_signal.bias": 0.1 * torch.randn(D, generator=gen),
             "conv_gate.weight": 0.06 * torch.randn(D, R, 2, generator=gen),
             "conv_gate.bias": 0.1 * torch.randn(D, generator=gen),
             "proj_signal.weight": 0.1 * torch.randn(D, Cc, 1, generator=gen),
             "proj_gate.weight": 0.1 * torch.randn(D, Cc, 1, generator=gen),
             "dil_skp.weight": 0.08 * torch.randn(S, D, 1, generator=gen)}
        if not final:
            p["dil_res.weight"] = 0.08 * torch.randn(R, D, 1, generator=gen)
        params.append({k: v.to(dev).contiguous() for k, v in p.items()})
    return params


def reference_stack(x0, cond, params, dils, final_last, rounded, layer_inputs=None):
    """float64 restatement of the stack on the absolute time axis (wavenet.py:91-111), optionally with the
    kernel's fp16 operand roundings.  Returns per-layer (th, sg, x_next) and the skip sum; entries left of a layer's lead
    are garbage by construction and are not compared.  ``layer_inputs`` (the kernel's own fp32 layer inputs) makes every
    layer an independent check: a 1-ulp flip of an fp16 operand rounding in layer l would otherwise be amplified by every
    later layer's roundings (~2x per layer), which says nothing about layer l + 1."""
    r = h if rounded else (lambda t: t.double())
    B, R, T0 = x0.shape
    x = x0.double()
    c = cond.double()
    outs, skp_sum, lead = [], 0, 0
    RF = sum(dils)
    for l, d in enumerate(dils):
        p = params[l]
        final = final_last and l == len(dils) - 1
        lead += d
        if layer_inputs is not None:
            x = layer_inputs[l].double()
        xs = F.pad(r(x), (d, 0))[:, :, :T0]                       # x[tau - d]
        pre = []
        for wk, pk, bk in (("conv_signal.weight", "proj_signal.weight", "conv_signal.bias"),
                           ("conv_gate.weight", "proj_gate.weight", "conv_gate.bias")):
            w = r(p[wk])
            pre.append(torch.einsum("dr,brt->bdt", w[:, :, 0], xs) + torch.einsum("dr,brt->bdt", w[:, :, 1], r(x)) +
                       torch.einsum("dc,bct->bdt", r(p[pk])[:, :, 0], r(c)) + r(p[bk])[None, :, None])
        th, sg = torch.tanh(pre[0]), torch.sigmoid(pre[1])
        z = r(th * sg)
        skp_sum = skp_sum + torch.einsum("sd,bdt->bst", r(p["dil_skp.weight"])[:, :, 0], z)
        x_next = None
        if not final:
            x_next = torch.einsum("rd,bdt->brt", r(p["dil_res.weight"])[:, :, 0], z) + x
            x = x_next
        outs.append((th, sg, x_next, lead))
    return outs, skp_sum, RF


def rel(a, b):
    return float((a.double() - b.double()).abs().max()) / max(float(b.abs().max()), 1e-12)


CASES = [
    # R, D, S, Cc, dils, T0, B, final_last      (arch.basic widths; a narrow 128-channel variant; a 512-channel one)
    (368, 256, 256, 138, [1, 2, 4, 8, 16, 32], 63 + 700, 2, True),
    (64, 128, 64, 20, [1, 2, 4], 7 + 300, 3, False),
    (512, 256, 256, 138, [4, 1], 5 + 260, 1, True),
]


@pytest.mark.parametrize("case", CASES, ids=["basic368", "narrow128", "wide512"])
@pytest.mark.parametrize("save", [True, False], ids=["train", "infer"])
def test_fused_stack_matches_fp16_operand_reference(case, save):
    from aewn import ops
    R, D, S, Cc, dils, T0, B, final_last = case
    dev = torch.device("cuda")
    gen = torch.Generator().manual_seed(11)
    params = make_params(R, D, S, Cc, dils, final_last, gen, dev)
    x0 = torch.randn(B, R, T0, generator=gen).to(dev)
    cond = torch.randn(B, Cc, T0, generator=gen).to(dev)
    ops.set_fused_forward(True)
    geom = ops.StackGeom(dils, T0, last_is_final=final_last)
    plan = ops.StackPlan(B, R, D, S, Cc, geom, params, dev, relu_last=False)
    assert plan.fused
    plan.sig[0][:, :, :T0] = x0
    for l, d in enumerate(dils):                      # the TF32 backward's pre-shifted tap of layer 0 (base layer's job)
        if l == 0 and 0 in plan.xs:
            plan.xs[0][:, :, d:T0] = x0[:, :, :T0 - d]
    plan.cond[:, :Cc, :T0] = cond
    plan.forward(save=save)
    ops._plans["t"] = plan
    try:
        ops.check_device_errors()
    finally:
        ops._plans.pop("t")
    outs, skp_ref, RF = reference_stack(x0, cond, params, dils, final_last, rounded=True,
                                        layer_inputs=[plan.sig[l][:, :, :T0] for l in range(len(dils))])
    worst = {}
    last_writer = max([l for l, o in enumerate(outs) if o[2] is not None], default=-1)
    for l, (th, sg, xn, lead) in enumerate(outs):
        if save:
            # saved for the backward pass: {fp16 a, fp16 b} per element, a = sg (1 - th^2), b = th sg (1 - sg)
            ab = plan.th[l].view(torch.float16).view(B, D, -1, 2)[:, :, lead:T0].float()
            worst[f"a16_{l}"] = rel(ab[..., 0], (sg * (1 - th * th))[:, :, lead:])
            worst[f"b16_{l}"] = rel(ab[..., 1], (th * sg * (1 - sg))[:, :, lead:])
            if plan.wgrad16:                      # fp16 channels-last copy (the weight gradients' operand): 2^-11 relative
                z16 = plan.z16[l][:, lead:T0, :D].permute(0, 2, 1).float()
                worst[f"z16_{l}"] = rel(z16, (th * sg)[:, :, lead:])
                assert float(plan.z16[l][:, :lead & ~3].abs().max() if lead & ~3 else 0.0) == 0.0
            else:
                worst[f"z{l}"] = rel(plan.z[l][:, :, lead:T0], (th * sg)[:, :, lead:])
            # margins: zero between the aligned-down lead and the lead (TMA boxes of the backward pass read them)
            assert float(plan.th[l][:, :, lead & ~3:lead].abs().max() if lead & 3 else 0.0) == 0.0
        if xn is not None:
            worst[f"x{l + 1}"] = rel(plan.sig[l + 1][:, :, lead:T0], xn[:, :, lead:])
            if l == last_writer:                      # the fp16 channels-last copy the next layer would read
                got16 = plan.x16[(l + 1) % plan.n_x16][:, lead:T0, :R].permute(0, 2, 1).float()
                worst[f"x16_{l + 1}"] = rel(got16, xn[:, :, lead:])           # fp16 storage: 2^-11 relative
                assert float(plan.x16[(l + 1) % plan.n_x16][:, :, R:].abs().max() if plan.KR16 > R else 0.0) == 0.0
            if save and l + 1 < len(dils) and (l + 1) in plan.xs:
                dn = dils[l + 1]
                assert torch.equal(plan.xs[l + 1][:, :, lead + dn:T0], plan.sig[l + 1][:, :, lead:T0 - dn])
    worst["skp"] = rel(plan.skp[:, :, RF:T0], skp_ref[:, :, RF:])
    bad = {k: v for k, v in worst.items() if v > (1e-3 if k.startswith(("x16", "a16", "b16", "z16")) else 2e-4)}   # fp16 storage
    assert not bad, (bad, worst)
    # the un-rounded fp32 math is within the TF32-class envelope
    outs32, skp32, _ = reference_stack(x0, cond, params, dils, final_last, rounded=False)
    assert rel(plan.skp[:, :, RF:T0], skp32[:, :, RF:]) < 5e-3


def test_fused_forward_agrees_with_two_launch_tf32_path_through_wavenet_step():
    """Whole decoder train step at arch.basic widths (window 512, batch 2): fused fp16-operand forward + TF32 backward
    vs the all-TF32 two-launch path.  The backward pass consumes what the fused forward saved (tanh, sigmoid, z, the
    pre-shifted taps), so agreement of every gradient pins those side outputs too."""
    import aewn
    from aewn import ops
    from test_gpu_fullsize import build

    def run(fused):
        ops.set_fused_forward(fused)
        torch.manual_seed(2507)
        wn, geo = build(512)
        wn = wn.cuda().train()
        g = torch.Generator().manual_seed(3)
        B = 2
        wav = torch.randint(0, 256, (B, geo["wav_len"]), generator=g).float().cuda()
        lc = torch.randn(B, 64, geo["lc_len"], generator=g).cuda().requires_grad_(True)
        spk = torch.randint(0, 40, (B,), generator=g).cuda()
        jit = torch.arange(geo["lc_len"]).unsqueeze(0).repeat(B, 1).cuda()
        q = wn(wav, lc, spk, jit)
        o0, o1 = wn.wav_cond_offset
        loss = aewn.RecLoss()(q[..., :-1], wav[:, o1 - 512:o1][..., 1:])
        loss.backward()
        ops.check_device_errors()
        assert any(p.fused == fused for p in ops._plans.values())
        return q.detach(), float(loss), {k: p.grad.clone() for k, p in wn.named_parameters()}, lc.grad.clone()

    try:
        q1, l1, g1, lc1 = run(True)
        q0, l0, g0, lc0 = run(False)
    finally:
        ops.set_fused_forward(True)
    assert rel(q1, q0) < 5e-3
    assert abs(l1 - l0) < 2e-3
    assert rel(lc1, lc0) < 3e-2
    # each path is within 3e-2 of the fp32 gradients (8e-2 for the conditioning front-end's parameters, which sit behind
    # the whole stack and cuDNN's own TF32 backward): the two differ by at most the sum
    for k in g0:
        assert rel(g1[k], g0[k]) < (1e-1 if k.startswith("lc_") else 7e-2), k
        a, b = g1[k].double().flatten(), g0[k].double().flatten()
        assert float(a @ b / (a.norm() * b.norm() + 1e-300)) > 0.999, k


def test_fp16_operand_range_overflow_is_reported():
    """An activation beyond the fp16 range (|x| > 65504) cannot be an exact tensor-core operand: the kernel saturates the
    copy and raises the device error word (include/aewn.h AEWN_ERR_RANGE) instead of silently computing with inf."""
    from aewn import ops, _lib
    dev = torch.device("cuda")
    gen = torch.Generator().manual_seed(1)
    R, D, S, Cc, dils, T0, B = 64, 128, 64, 20, [1, 2], 3 + 200, 1
    params = make_params(R, D, S, Cc, dils, True, gen, dev)
    ops.set_fused_forward(True)
    plan = ops.StackPlan(B, R, D, S, Cc, ops.StackGeom(dils, T0), params, dev, relu_last=False)
    plan.sig[0][:, :, :T0] = 1.0e5 * torch.randn(B, R, T0, generator=gen).to(dev)
    plan.forward(save=False)
    torch.cuda.synchronize()
    assert int(plan.err.item()) == _lib.ERR_RANGE
    plan.err.zero_()
    assert torch.isfinite(plan.skp).all()


def test_amax_pow2_scale():
    """aewn_amax_pow2_scale: scale = 2^floor(log2(target / max|x|)), 1 for an all-zero tensor, NaNs ignored."""
    import ctypes as C
    from aewn import _lib, ops
    dev = torch.device("cuda")
    work = torch.zeros(1, dtype=torch.int32, device=dev)
    out = torch.zeros(2, device=dev)

    def run(x, target=8.0):
        _lib.check(_lib.lib().aewn_amax_pow2_scale(C.c_void_p(x.data_ptr()), C.c_longlong(x.numel()), C.c_float(target),
                                                   C.c_void_p(work.data_ptr()), C.c_void_p(out.data_ptr()), ops._stream()),
                   "aewn_amax_pow2_scale")
        return [float(v) for v in out.cpu()]

    gen = torch.Generator().manual_seed(0)
    x = (torch.randn(1_000_003, generator=gen) * 3e-6).to(dev)
    x[777_777] = -7.3e-5
    s, inv = run(x)
    assert s == 2.0 ** 16 and inv == 2.0 ** -16            # 8 / 7.3e-5 = 1.096e5 -> 2^16 = 65536; amax * s = 4.78 in (4, 8]
    assert run(torch.zeros(4096, device=dev)) == [1.0, 1.0]
    x[5] = float("nan")
    assert run(x)[0] == 2.0 ** 16


def test_data_gradient_on_the_fused_engine_bf16_and_scaled_fp16():
    """AEWN_DGRAD16 = 1 (bf16 operands) and 2 (the default: fp16 operands with the per-step power-of-two scale) against the
    TF32 data gradient of the tgemm engine (0), whole decoder step at arch.basic widths.  Both stay inside the backward tolerance; the scaled-fp16 copy
    carries TF32's mantissa, so it must sit closer to the TF32 result than bf16 does."""
    import aewn
    from aewn import ops
    from test_gpu_fullsize import build

    def run(mode, wgrad16="0", gz16="0"):
        os.environ["AEWN_DGRAD16"] = mode
        os.environ["AEWN_WGRAD16"] = wgrad16
        os.environ["AEWN_GZ16"] = gz16
        ops._plans.clear()
        torch.manual_seed(2507)
        wn, geo = build(512)
        wn = wn.cuda().train()
        g = torch.Generator().manual_seed(3)
        B = 2
        wav = torch.randint(0, 256, (B, geo["wav_len"]), generator=g).float().cuda()
        lc = torch.randn(B, 64, geo["lc_len"], generator=g).cuda().requires_grad_(True)
        spk = torch.randint(0, 40, (B,), generator=g).cuda()
        jit = torch.arange(geo["lc_len"]).unsqueeze(0).repeat(B, 1).cuda()
        q = wn(wav, lc, spk, jit)
        o0, o1 = wn.wav_cond_offset
        loss = aewn.RecLoss()(q[..., :-1], wav[:, o1 - 512:o1][..., 1:])
        loss.backward()
        ops.check_device_errors()
        used = [p for p in ops._plans.values()]
        assert used and all(p.dgrad16 == (mode != "0") and p.dgrad16_scaled == (mode == "2") and
                            p.wgrad16 == (mode == "2" and wgrad16 == "1") and
                            p.gz16 == (mode == "2" and wgrad16 == "1" and gz16 == "1") for p in used)
        return {k: p.grad.double().clone() for k, p in wn.named_parameters()}, lc.grad.double().clone()

    old = os.environ.get("AEWN_DGRAD16"), os.environ.get("AEWN_WGRAD16"), os.environ.get("AEWN_GZ16")
    try:
        g0, lc0 = run("0")
        g1, lc1 = run("1")
        g2, lc2 = run("2")
        g3, lc3 = run("2", wgrad16="1")          # + the weight gradients on aewn_wgradh
        g4, lc4 = run("2", wgrad16="1", gz16="1")    # + the gate derivative's GEMM on the fused engine (aewn_grcc_gz)
    finally:
        for k, v in zip(("AEWN_DGRAD16", "AEWN_WGRAD16", "AEWN_GZ16"), old):
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        ops._plans.clear()

    def nrm(a, b):
        return float((a - b).norm() / b.norm())

    e1 = {k: nrm(g1[k], g0[k]) for k in g0}
    e2 = {k: nrm(g2[k], g0[k]) for k in g0}
    assert max(e1.values()) < 6e-2 and max(e2.values()) < 6e-2, (max(e1.values()), max(e2.values()))
    assert nrm(lc1, lc0) < 6e-2 and nrm(lc2, lc0) < 6e-2
    # the data gradient feeds every tensor upstream of the last layer: compare the two variants on the stack's weights
    ks = [k for k in g0 if k.startswith("conv_layers.")]
    m1 = sum(e1[k] for k in ks) / len(ks)
    m2 = sum(e2[k] for k in ks) / len(ks)
    print(f"mean norm-wise distance to the TF32 data gradient: bf16 {m1:.2e}, scaled fp16 {m2:.2e}")
    assert m2 < m1
    # fp16 weight gradients: same data gradient as mode 2 (lc gradient identical), weights within TF32-class distance
    e3 = {k: nrm(g3[k], g0[k]) for k in g0}
    m3 = sum(e3[k] for k in ks) / len(ks)
    print(f"  + fp16 weight gradients: {m3:.2e} (worst {max(e3[k] for k in ks):.2e})")
    assert torch.equal(lc3, lc2)
    assert max(e3.values()) < 6e-2 and m3 < 2e-3
    # fp16 gate derivative: g_z from fp16 copies of g_x / g_skp instead of TF32 reads of the fp32 tensors
    e4 = {k: nrm(g4[k], g0[k]) for k in g0}
    m4 = sum(e4[k] for k in ks) / len(ks)
    print(f"  + fp16 gate derivative: {m4:.2e} (worst {max(e4[k] for k in ks):.2e}), lc {nrm(lc4, lc0):.2e}")
    assert max(e4.values()) < 6e-2 and m4 < 2e-3 and nrm(lc4, lc0) < 6e-2


def test_scaled_gradient_overflow_is_reported():
    """The fp16 gradient copies carry max|g_skp| * s in (4, 8]: a gradient more than 2^13 times larger than max|g_skp| (here:
    skip weights of 1e6 times the usual size, so that g_z = Ws^T g_skp dwarfs g_skp -- and the fp16 copy of those weights
    overflows as well) cannot be represented and must raise AEWN_ERR_RANGE instead of producing saturated or NaN gradients
    silently; with ordinary weights the same backward leaves the error word clear.  (Residual or convolution weights cannot
    drive this: large pre-activations saturate the gates and the gradients vanish instead.)"""
    from aewn import ops, _lib
    dev = torch.device("cuda")
    R, D, S, Cc, dils, T0, B = 64, 128, 64, 20, [1, 2], 3 + 300, 1

    def run(skip_gain):
        gen = torch.Generator().manual_seed(2)
        params = make_params(R, D, S, Cc, dils, True, gen, dev)
        for p in params:
            p["dil_skp.weight"].mul_(skip_gain)
        ops.set_fused_forward(True)
        plan = ops.StackPlan(B, R, D, S, Cc, ops.StackGeom(dils, T0), params, dev, relu_last=False)
        assert plan.dgrad16_scaled and plan.wgrad16
        plan.sig[0][:, :, :T0] = torch.randn(B, R, T0, generator=gen).to(dev)
        plan.cond[:, :Cc, :T0] = torch.randn(B, Cc, T0, generator=gen).to(dev)
        plan.forward(save=True)
        torch.cuda.synchronize()
        assert int(plan.err.item()) == 0                       # the forward pass itself is in range (the skip sum is fp32)
        bw = plan.bwd()
        bw["g_skp"][:, :, plan.geom.RF:T0] = torch.randn(B, S, T0 - plan.geom.RF, generator=gen).to(dev) * 1e-5
        plan.backward()
        torch.cuda.synchronize()
        e = int(plan.err.item())
        plan.err.zero_()
        return e

    assert run(1.0) == 0
    assert run(1.0e6) == _lib.ERR_RANGE
