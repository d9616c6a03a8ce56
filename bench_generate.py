"""Generation-rate micro-benchmark for the persistent sampler (WaveNet.forward_test, csrc/gen.cu).

    python bench_generate.py [--n-rep 1] [--steps 8192] [--eager-steps 40]

Prints one JSON line: samples/s of the persistent kernel at par/arch.basic.json widths (random-init weights,
synthetic conditioning), and -- as the library baseline the reference's sampler reduces to on a GPU -- the rate of an
eager PyTorch incremental loop that issues the same per-sample ops the reference does (wavenet.py:455-509: base layer,
20 gated layers on (n_rep, R, d+1) slices, post-net, softmax, multinomial) on the same device.  Not part of the
headline bench (bench.py); the numbers go into profiles/README.md.
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "ae-wavenet_b200"))

ARCH_BASIC = dict(filter_sz=2, n_lc_out=128, lc_upsample_strides=[5, 4, 4, 4], lc_upsample_filt_sizes=[25, 16, 16, 16],
                  n_res=368, n_dil=256, n_skp=256, n_post=256, n_quant=256, n_blocks=2, n_block_layers=10,
                  n_global_embed=10, n_speakers=40, bias=True, n_lc_in=64)


class HP(dict):
    __getattr__ = dict.__getitem__


def build():
    import aewn
    from aewn import geometry as vc
    torch.manual_seed(2507)
    parent = vc.VirtualConv(filter_info=1, stride=320, parent=None, name="LC-grid")
    wn = aewn.WaveNet(HP(ARCH_BASIC), parent_vc=parent)
    vc.compute_inputs(wn.vc["end_grcc"], vc.GridRange((0, 10 ** 7), (0, 1024), 1))
    wn.trim_ups_out = torch.tensor([0, wn.vc["beg_grcc"].in_len()], dtype=torch.long)
    wn.post_init(1024)
    return wn.cuda().eval()


@torch.no_grad()
def eager_incremental(wn, cond, codes, n_rep, steps):
    """The reference's per-sample op sequence with plain PyTorch ops on the GPU (incremental mode, ring buffers kept as
    growing lists of the last d+1 columns).  Timing stand-in for the reference's sampler; values are not checked."""
    dev = cond.device
    layers = list(wn.conv_layers)
    rf1 = wn.base_global_rf
    hist = [torch.zeros(n_rep, wn.n_res, layer.dil + 1, device=dev) for layer in layers]
    cur = codes[:rf1].clone().unsqueeze(0).repeat(n_rep, 1)
    last = cur[:, -1]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in range(steps):
        x = F.conv1d(F.one_hot(last, wn.n_quant).float().unsqueeze(2), wn.base_layer.weight, wn.base_layer.bias)
        c = cond[:, :, rf1 - 1 + s:rf1 + s].expand(n_rep, -1, -1)
        skp_sum = 0
        for li, layer in enumerate(layers):
            hist[li] = torch.cat((hist[li][:, :, 1:], x), 2)
            xin = hist[li]
            filt = layer.conv_signal(xin) + layer.proj_signal(c)
            gate = layer.conv_gate(xin) + layer.proj_gate(c)
            z = torch.tanh(filt) * torch.sigmoid(gate)
            skp_sum = skp_sum + layer.dil_skp(z)
            if not layer.final_layer:
                x = layer.dil_res(z) + x
        quant = F.conv1d(F.relu(F.conv1d(F.relu(skp_sum), wn.post1.weight, wn.post1.bias)), wn.post2.weight,
                         wn.post2.bias).squeeze(2)
        last = torch.multinomial(F.softmax(quant, -1), 1, True).squeeze(1)
    torch.cuda.synchronize()
    return steps * n_rep / (time.perf_counter() - t0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-rep", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8192)
    ap.add_argument("--eager-steps", type=int, default=40)
    a = ap.parse_args()
    from aewn import generate, ops
    wn = build()
    rf1 = wn.base_global_rf
    n_ts = rf1 + a.steps
    g = torch.Generator().manual_seed(1)
    cond = torch.randn(wn.n_cond, n_ts, generator=g).cuda()
    codes = torch.randint(0, 256, (n_ts + 8,), generator=g).cuda()
    plan = generate.get_plan(wn, a.n_rep)
    plan.generate(codes, cond[:, :rf1 + 64], rf1)          # warm-up (module load, L2 fill)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    out = plan.generate(codes, cond, rf1)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1])
    total_steps = n_ts - 1                                 # priming steps run the same stack (no post-net)
    rate = a.n_rep * total_steps / (ms * 1e-3)
    eager = eager_incremental(wn, cond.unsqueeze(0), codes, a.n_rep, a.eager_steps) if a.eager_steps > 0 else None
    stream_mb = plan.wstream.numel() * 4 / 1e6
    print(json.dumps({"metric": "generated audio samples/s (forward_test)", "value": rate, "unit": "samples/s",
                      "n_rep": a.n_rep, "steps": total_steps, "ms": ms, "us_per_step": 1e3 * ms / total_steps,
                      "cluster": plan.cluster, "n_stages": plan.desc.n_stages, "stage_bytes": plan.desc.stage_bytes,
                      "weights_streamed_mb_per_step": stream_mb,
                      "l2_stream_gbps": stream_mb * 1e-3 * total_steps / (ms * 1e-3),
                      "eager_torch_samples_per_s": eager,
                      "speedup_vs_eager": (rate / eager) if eager else None,
                      "generated_differs_from_input": bool((out[0, rf1:n_ts].cpu() != codes[rf1:n_ts].cpu()).any())}))


if __name__ == "__main__":
    main()
