#!/usr/bin/env python
"""bench.py -- audio samples/s through the WaveNet decoder train step (cfg2 of BASELINE.json) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

Workload: WaveNet decoder of par/arch.basic.json (20 GRCC layers, R=368 D=256 S=256 C=138),
batch 8 per GPU, window 16384 (decoder input T0 = 18430), synthetic 16 kHz mu-law codes, random-init weights
(seed 2507).  One step = H2D of the batch (e2e leg only), forward, RecLoss, backward, one flat-buffer all-reduce
(N > 1), Adam step.  `value` = B_total * W / t_step with inputs resident in HBM, timed with CUDA events, max over ranks.

`--impl reference` times the reference's own CPU path (the oracle port of wavenet.py, oracle/torch_oracle.py) on the
host cores on a bounded sample of the same workload.  oracle/ is touched ONLY by that leg and by the cpu_baseline leg.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "ae-wavenet_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

os.environ.setdefault("NCCL_DEBUG", "WARN")
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's version banner / warnings go to stderr: stdout carries ONE JSON line

import torch  # noqa: E402

ARCH_BASIC = dict(filter_sz=2, n_lc_out=128, lc_upsample_strides=[5, 4, 4, 4], lc_upsample_filt_sizes=[25, 16, 16, 16],
                  n_res=368, n_dil=256, n_skp=256, n_post=256, n_quant=256, n_blocks=2, n_block_layers=10,
                  n_global_embed=10, n_speakers=40, bias=True, n_lc_in=64)
METRIC = "audio samples/s through WaveNet fwd+bwd"
UNIT = "samples/s"


class HP(dict):
    __getattr__ = dict.__getitem__


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        d = json.load(open(path))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d.get("bf16_tflops_sustained", d["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1400.0, source="fallback")


def synth_batch(B, wav_len, lc_len, n_lc_in, n_speakers, seed):
    """Synthetic 16 kHz window: mu-law codes of a two-tone + noise signal (util.mu_encode_np, util.py:62-67)."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(wav_len).float()
    f1 = 0.01 + 0.02 * torch.rand(B, 1, generator=g)
    f2 = 0.05 + 0.10 * torch.rand(B, 1, generator=g)
    x = 0.3 * torch.sin(f1 * t) + 0.2 * torch.sin(f2 * t + 1.0) + 0.05 * torch.randn(B, wav_len, generator=g)
    x = x.clamp(-1, 1)
    mu = 255.0
    amp = torch.sign(x) * torch.log1p(mu * x.abs()) / torch.log1p(torch.tensor(mu))
    wav = ((amp + 1) * 0.5 * mu + 0.5).to(torch.int32).float()
    lc = torch.randn(B, n_lc_in, lc_len, generator=g)
    spk = torch.randint(0, n_speakers, (B,), generator=g)
    jit = torch.arange(lc_len).unsqueeze(0).repeat(B, 1)
    return wav, lc, spk, jit


def build_decoder(W, WaveNet, vc, arch=None):
    """Stand-alone decoder built exactly like MfccInverter._init_geometry (mfcc_inverter.py:38-65)."""
    hp = HP(arch or ARCH_BASIC)
    parent = vc.VirtualConv(filter_info=1, stride=320, parent=None, name="LC-grid")
    wn = WaveNet(hp, parent_vc=parent)
    end_gr = vc.GridRange((0, 10 ** 7), (0, W), 1)
    vc.compute_inputs(wn.vc["end_grcc"], end_gr)
    beg = wn.vc["beg_grcc"]
    geo = dict(wav_len=parent.in_len(), lc_len=parent.child.in_len(), dec_in_len=beg.in_len(),
               trim_dec_out=[end_gr.sub[0] - parent.input_gr.sub[0], end_gr.sub[1] - parent.input_gr.sub[0]])
    wn.trim_ups_out = torch.tensor([0, beg.in_len()], dtype=torch.long)
    wn.post_init(W)
    return wn, geo


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                          "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                         text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = max([int(float(r[2])) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()] or [0])
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=mx or None, reasons=sorted(reasons),
                    samples=len(sm))


def grcc_layer_fwd_bytes(B, R, D, S, C, T_in, d, W, train):
    """read x, read cond slice, write sig, RMW skip-sum, weights once (+ saved tanh/sigmoid/z if training)."""
    T_out = T_in - d
    n_weights = 2 * (D * R * 2) + 2 * D + 2 * D * C + S * D + R * D          # 607 744 for arch.basic
    b = 4 * B * (R * T_in + C * T_out + R * T_out + 2 * S * W) + 4 * n_weights
    if train:
        b += 4 * B * 2 * D * T_out          # + the saved tanh and sigmoid (z is their product, not counted)
    return b


def grcc_layer_fwd_flops(B, R, D, S, C, T_in, d, W, K=2):
    T_out = T_in - d
    return 2.0 * B * T_out * D * (2 * K * R + 2 * C) + 2.0 * B * T_out * R * D + 2.0 * B * W * S * D


# ------------------------------------------------------------------------------------------------- reference arm / CPU
def port_step(B, W, steps, warmup, seed=2507, device="cpu"):
    """The reference's own implementation of the path (oracle port of wavenet.py:323-364 + RecLoss + autograd): the same
    ATen calls the reference modules make.  device="cpu": the CPU baseline / reference arm; device="cuda": the GPU LIBRARY
    baseline (eager PyTorch through cuDNN on the same B200, SURVEY.md 2b / BASELINE.md 3)."""
    from oracle import torch_oracle as orc
    from aewn import geometry as vc
    import aewn
    torch.manual_seed(seed)
    with torch.device("cpu"):
        wn, geo = build_decoder(W, aewn.WaveNet, vc)
    sd = {k: (v.to(device).requires_grad_(True) if v.dtype == torch.float32 and k != "cond.eye" else v.to(device))
          for k, v in wn.state_dict().items()}
    ogeo = dict(trim_ups_out=wn.trim_ups_out.tolist(), wav_cond_offset=list(wn.wav_cond_offset),
                leads=[l.leads.tolist() for l in wn.conv_layers], n_win_batch=W, trim_dec_out=geo["trim_dec_out"])
    wav, lc, spk, jit = [t.to(device) for t in synth_batch(B, geo["wav_len"], geo["lc_len"], 64, 40, 1234)]
    cuda = str(device).startswith("cuda")
    times = []
    for i in range(warmup + steps):
        if cuda:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        loss, _ = orc.decoder_loss(sd, ARCH_BASIC, ogeo, wav, lc, spk, jit)
        loss.backward()
        for v in sd.values():
            if getattr(v, "grad", None) is not None:
                v.grad = None
        if cuda:
            torch.cuda.synchronize()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    times.sort()
    return B * W / times[len(times) // 2], times


def cpu_port_step(B, W, steps, warmup, seed=2507):
    return port_step(B, W, steps, warmup, seed, "cpu")


def gpu_library_baseline(B, W, dev, steps=3, warmup=1):
    """Eager PyTorch / cuDNN on the same GPU, same configuration as the headline: once with the framework's default conv
    precision (TF32) and once with IEEE fp32 convolutions.  The reference ships no CUDA kernels of its own, so this is the
    only pre-existing Blackwell code path for its hot path."""
    out = {}
    prev = torch.backends.cudnn.allow_tf32
    try:
        for name, tf32 in (("tf32", True), ("ieee", False)):
            torch.backends.cudnn.allow_tf32 = tf32
            try:
                sps, times = port_step(B, W, steps, warmup, device=dev)
                out[name] = dict(value=sps, unit=UNIT, ms_per_step=1e3 * times[len(times) // 2], steps=len(times))
            except Exception as e:   # noqa: BLE001 -- e.g. out of memory: report, do not fail the bench
                out[name] = dict(unavailable=f"{type(e).__name__}: {str(e)[:160]}")
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    out["what"] = (f"oracle port of the reference modules (same ATen ops: conv1d, conv_transpose1d, one_hot, tanh, sigmoid, "
                   f"log_softmax + autograd), eager on this GPU, batch {B} x window {W}, fwd + RecLoss + bwd")
    return out


def pick_cpu_threads():
    """All the host threads the port can USE: oneDNN conv on a 2 x 2048 window stops scaling (and regresses badly) long
    before 128 threads, so probe a few counts with one step each and keep the fastest."""
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64) if c <= ncpu}) or [ncpu]
    best, best_sps = cands[0], 0.0
    for c in cands:
        torch.set_num_threads(c)
        sps, _ = cpu_port_step(1, 512, 1, 1)
        if sps > best_sps:
            best, best_sps = c, sps
    torch.set_num_threads(best)
    return best


def host_mem_gb():
    try:
        import psutil
        return psutil.virtual_memory().available / 2 ** 30
    except Exception:   # noqa: BLE001
        return 0.0


def run_reference(args, rank):
    """`--impl reference`: the reference's own CPU implementation of the path (the oracle port: the reference tree does not
    exist on the bench box) on the host cores.  Runs the HEADLINE configuration (cfg2: batch 8 x window 16384, ~10-20 s per
    step, ~45 GB of autograd state) when the host has the memory for it, 1 warm-up + min(K, 3) steps; otherwise -- and
    always as a second key -- a bounded sample of the same network (batch 2 x window 2048)."""
    if rank != 0:
        return
    pick_cpu_threads()
    full = host_mem_gb() >= 96 and (os.cpu_count() or 1) >= 16 and os.environ.get("AEWN_REF_SMALL", "0") != "1"
    sB, sW = 2, 2048
    s_sps, s_times = cpu_port_step(sB, sW, max(1, min(args.steps, 3)), 1)
    sample = dict(value=s_sps, unit=UNIT, global_batch=sB, window=sW, ms_per_step=1e3 * s_times[len(s_times) // 2])
    if full:
        B, W = 8, 16384
        sps, times = cpu_port_step(B, W, max(1, min(args.steps, 3)), 1)
    else:
        B, W, sps, times = sB, sW, s_sps, s_times
    out = dict(metric=METRIC, value=sps, unit=UNIT, n_gpus=args.gpus, steps=len(times), warmup=1,
               ms_per_step=1e3 * times[len(times) // 2], higher_is_better=True, scaling="weak", vs_baseline=None,
               dtype="f32", data="synthetic", impl="reference",
               config=dict(workload=f"cfg2: WaveNet decoder par/arch.basic.json train step (fwd+RecLoss+bwd), batch {B}, "
                                    f"window {W}, CPU", global_batch=B, window=W, same_config_as_headline=full),
               cpu_baseline=dict(value=sps, unit=UNIT, cores=torch.get_num_threads(), kind="port",
                                 sample=(f"the headline configuration itself (batch {B} x window {W}), median of "
                                         f"{len(times)} steps after 1 warm-up" if full else
                                         f"same decoder, batch {B} x window {W} (cfg2 is 8 x 16384; this host has "
                                         f"{host_mem_gb():.0f} GB free / {os.cpu_count()} cores), median step")),
               bounded_sample=sample,
               e2e=dict(value=sps, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------- our arm
def bench_workload(args, workload, rank, local_rank, world, dev, light=False):
    """One workload on this rank's GPU; returns the JSON dict on rank 0 (None elsewhere).  light=True: timing only (no
    per-launch instrumentation, roofline, baselines) -- used for the extra cfg4 / cfg5 lines of a multi-GPU run."""
    import torch.distributed as dist
    import aewn
    from aewn import geometry as vc, ops, _lib
    from aewn.dist import FlatGradSync

    cfg3, cfg5 = workload == "cfg3", workload == "cfg5"
    W = args.window if not (cfg5 and args.window == 16384) else 65536
    B = (args.batch if workload == args.workload else 0) or (16 if cfg3 else 2 if cfg5 else 8)
    arch = dict(ARCH_BASIC, n_blocks=3, n_res=512) if cfg5 else ARCH_BASIC
    torch.manual_seed(2507)                      # identical replicas
    if cfg3:
        from aewn.autoencoder import AutoEncoder
        ae = AutoEncoder(HP(dict(ARCH_BASIC, n_lc_in=32)), 39, 768, "vqvae-ema", 32, 0.25, 0.99, 4096, True)
        ae.init_geometry(W)
        ae = ae.to(dev).train()
        wn = ae.decoder
        geo = dict(wav_len=ae.dec_in_len, lc_len=ae.embed_len, dec_in_len=ae.dec_in_len)
        model = ae
        sync = FlatGradSync(ae.parameters(), vqema=ae.bottleneck, fused_accumulate=True)
        wav_h, _, spk_h, jit_h = synth_batch(B, ae.dec_in_len, ae.embed_len, 32, 40, 1234 + rank)
        lc_h = torch.randn(B, 39, ae.enc_in_mel_len, generator=torch.Generator().manual_seed(99 + rank))   # mel input
        wav_h, lc_h, spk_h, jit_h = [t.pin_memory() for t in (wav_h, lc_h, spk_h, jit_h)]

        def fwd_bwd(wav, mels, spk, jit):
            sync.zero_grad()
            pred, target, com, rec = ae.run(mels, wav, spk, jit)
            loss = com + rec
            loss.backward()
            return loss

        def step(wav, mels, spk, jit):
            loss = fwd_bwd(wav, mels, spk, jit)
            sync.sync()                          # ONE all-reduce: grads | z_sum | n_sum | metrics, then the EMA update
            opt.step()
            return loss
        graph_cfg3 = not args.eager_cfg3
        if graph_cfg3:
            ae.bottleneck.static_diagnostics = True
    else:
        wn, geo = build_decoder(W, aewn.WaveNet, vc, arch)
        wn = wn.to(dev).train()
        model = wn
        loss_fn = aewn.RecLoss()
        sync = FlatGradSync(wn.parameters(), fused_accumulate=True)
        wav_h, lc_h, spk_h, jit_h = [t.pin_memory() for t in
                                     synth_batch(B, geo["wav_len"], geo["lc_len"], 64, 40, 1234 + rank)]
        t0w, t1w = geo["trim_dec_out"]

        def fwd_bwd(wav, lc, spk, jit):
            sync.zero_grad()
            quant = wn(wav, lc, spk, jit)
            loss = loss_fn(quant[..., :-1], wav[:, t0w:t1w][..., 1:])
            loss.backward()
            return loss

        def step(wav, lc, spk, jit):
            loss = fwd_bwd(wav, lc, spk, jit)
            sync.sync()
            opt.step()
            return loss
        graph_cfg3 = False
    # checkpoint.py:49, par/train.basic.json:6; capturable: the step counter lives on the device (CUDA-graph replay)
    opt = torch.optim.Adam(model.parameters(), lr=2e-5, fused=True, capturable=True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    dwav, dlc, dspk, djit = [t.to(dev) for t in (wav_h, lc_h, spk_h, jit_h)]
    for _ in range(args.warmup):
        loss = step(dwav, dlc, dspk, djit)
    ops.check_device_errors()

    # The public training-step API (aewn.train.GraphedStep): the step replayed as ONE CUDA graph (one GPU), or zero-grad +
    # forward + loss + backward as a graph followed by the eager NCCL all-reduce and Adam (N GPUs: collectives stay out
    # of the capture).  The VQ-VAE step is captured with VQEMA.static_diagnostics (unique() replaced by a static-shape
    # count; --eager-cfg3 keeps the reference's diagnostics and runs eagerly).  Falls back to eager steps if capture
    # fails, and says so in `config`.
    eager_step, graph_note = step, "eager"
    launches_per_step = None
    if not args.no_graph and (not cfg3 or graph_cfg3):
        try:
            from aewn.train import GraphedStep
            loss = None                    # drop the eager warm-up's autograd graph: its AccumulateGrad nodes are bound
            if cfg3:                       # to the default stream and would invalidate the capture; same for the module
                ae.encoding_bn = None      # attributes that hold graph tensors
                ae.bottleneck.ze = ae.bottleneck.min_dist = None
                ae.objective.metrics = {}
            torch.cuda.synchronize()
            l0 = _lib.launch_count()
            gstep = GraphedStep(step if world == 1 else fwd_bwd, [dwav, dlc, dspk, djit], warmup=1,
                                capture_on_warmup_stream=cfg3)
            launches_per_step = (_lib.launch_count() - l0) // 2      # 1 warm-up step + 1 captured step
            if world == 1:
                step = gstep
            else:
                def step(wav, lc, spk, jit):
                    loss = gstep(wav, lc, spk, jit)
                    sync.sync()
                    opt.step()
                    return loss
            dwav, dlc, dspk, djit = gstep.static_in
            graph_note = "cuda_graph" if world == 1 else "cuda_graph(fwd+bwd) + eager all-reduce + Adam"
        except Exception as e:   # noqa: BLE001 -- report and measure eagerly
            graph_note = f"eager (graph capture failed: {type(e).__name__}: {str(e)[:120]})"
            step = eager_step
            torch.cuda.synchronize()

    # ---- timed region 1: device-resident inputs, CUDA events around the K steps (no per-launch instrumentation)
    sampler = ClockSampler(local_rank)
    barrier()
    launches0 = _lib.launch_count()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    # a START/END range (process-wide, unlike push/pop which is per thread: backward runs on autograd's thread) lets ncu
    # restrict a capture to the timed steps:  ncu --nvtx --nvtx-include "aewn_timed" ...
    nvtx_id = torch.cuda.nvtx.range_start("aewn_timed")
    for _ in range(args.steps):
        loss = step(dwav, dlc, dspk, djit)
    torch.cuda.nvtx.range_end(nvtx_id)           # (ncu filters on the LAUNCH being inside the range: no sync needed)
    e1.record()
    barrier()
    clocks = sampler.stop()
    launches = _lib.launch_count() - launches0
    if launches_per_step is not None and graph_note.startswith("cuda_graph"):
        launches = launches_per_step * args.steps          # replayed from the graph: no host-side launch calls to count
    ms = e0.elapsed_time(e1) / args.steps
    final_loss = float(loss.detach())
    ops.check_device_errors()

    prof_ms, times, itimes = None, {}, {}
    if not light:
        # ---- timed region 1b: the same K steps with a CUDA-event pair around EVERY kernel launch of ours (on the
        # launching stream): per-kernel durations for the roofline / per-class shares.  Kept out of region 1 because ~250
        # extra events per step cost ~2 % of the step.
        prof = ops.LaunchProfiler()
        ops.set_profiler(prof)
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(args.steps):
            loss = eager_step(dwav, dlc, dspk, djit)           # eager: a graph replay cannot carry per-launch events
        p1.record()
        barrier()
        ops.set_profiler(None)
        prof_ms = p0.elapsed_time(p1) / args.steps
        times = prof.times_ms()
        # ---- region 1c: forward only, nothing kept for a backward pass (torch.no_grad(): the layers skip the tanh /
        # sigmoid / z writes) -- the inference byte count of SURVEY.md 8d / the per-layer target of BASELINE.md 5
        if not cfg3:
            iprof = ops.LaunchProfiler()
            with torch.no_grad():
                for _ in range(2):
                    wn(dwav, dlc, dspk, djit)
                ops.set_profiler(iprof)
                for _ in range(args.steps):
                    wn(dwav, dlc, dspk, djit)
                torch.cuda.synchronize()
                ops.set_profiler(None)
            itimes = iprof.times_ms()

    # ---- timed region 2: end to end through the public module API with HOST (pinned) inputs and a D2H loss read
    barrier()
    t_start = time.perf_counter()
    for _ in range(args.steps):
        if graph_note.startswith("cuda_graph"):
            loss = step(wav_h, lc_h, spk_h, jit_h)    # H2D copies from pinned memory straight into the graph's inputs
        else:
            loss = step(*[t.to(dev, non_blocking=True) for t in (wav_h, lc_h, spk_h, jit_h)])
        _ = loss.item()                           # D2H read of the step's result
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t_start) / args.steps

    if world > 1:
        t = torch.tensor([ms, e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(t[0]), float(t[1])
        lt = torch.tensor([float(launches)], device=dev)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt[0])

    out = None
    if rank == 0:
        pk = peaks()
        R, D, S, C = arch["n_res"], 256, 256, 138
        fused = bool(getattr(next(iter(ops._plans.values()), None), "fused", False)) if ops._plans else False
        out = dict(
            metric=METRIC, value=world * B * W / (ms * 1e-3), unit=UNIT, n_gpus=world, steps=args.steps,
            warmup=args.warmup, ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None,
            dtype=("fp16 tensor-core operands (10-bit mantissa, like TF32) in the forward layers and, scaled by a per-step "
                   "power of two, in the stack's data and weight gradients; tf32 operands elsewhere; fp32 accumulation, "
                   "residual stream, storage and optimizer" if fused and ops.dgrad16_mode() == "2" else
                   "fp16 tensor-core operands in the forward layers (10-bit mantissa, like TF32), tf32 operands elsewhere; "
                   "fp32 accumulation, residual stream, storage and optimizer" if fused else "tf32"), data="synthetic",
            config=dict(workload=("cfg3/cfg4: full VQ-VAE-EMA autoencoder par/arch.vqvae-ema.json train step (Encoder -> "
                                  f"VQEMA -> WaveNet), batch {B}/GPU, window {W}" if cfg3 else
                                  f"cfg5: deep decoder stress (30 dilation layers, 512 residual channels) train step, batch "
                                  f"{B}/GPU, window {W}" if cfg5 else
                                  f"cfg2: WaveNet decoder par/arch.basic.json train step, batch {B}/GPU, window {W}"),
                        step=("H2D(e2e only)+fwd+VQEMALoss+RecLoss+bwd+allreduce(grads|z_sum|n_sum)+EMA+Adam" if cfg3 else
                              "H2D(e2e only)+fwd+RecLoss+bwd+allreduce+Adam"), global_batch=world * B, window=W,
                        engine_mode=ops.ENGINE_MODE, fused_layer_forward=fused, step_execution=graph_note,
                        dec_in_len=geo["dec_in_len"], parallelism=f"dp{world}",
                        l2="activations per step ~13 GB >> 126 MB L2 (no flush needed)", seed=2507,
                        final_loss=final_loss),
            clocks=clocks,
            e2e=dict(value=world * B * W / (e2e_ms * 1e-3), unit=UNIT, ms_per_step=e2e_ms,
                     h2d_bytes_per_step=int(sum(t.numel() * t.element_size() for t in (wav_h, lc_h, spk_h, jit_h))),
                     d2h_bytes_per_step=4),
            gpu_launches=launches,
        )
        if not light:
            geomS = wn.stack_geometry(geo["dec_in_len"])

            def layer_sums(tms, train):
                tot_b = tot_f = tot_ms = 0.0
                per = []
                T_in = geomS.T0
                for l, d in enumerate(geomS.dils):
                    g1, g2 = tms.get(f"fwd_gemm1.{l}", []), tms.get(f"fwd_gemm2.{l}", [])
                    gl = tms.get(f"fwd_layer.{l}", [])          # fused layer kernel: one launch per layer
                    if gl or (g1 and g2):
                        l_ms = sum(gl) / len(gl) if gl else sum(g1) / len(g1) + sum(g2) / len(g2)
                        l_b = grcc_layer_fwd_bytes(B, R, D, S, C, T_in, d, W, train=train)
                        tot_ms += l_ms
                        tot_b += l_b
                        tot_f += grcc_layer_fwd_flops(B, R, D, S, C, T_in, d, W)
                        per.append(round(l_b / (l_ms * 1e-3) / 1e9, 1))
                    T_in -= d
                return tot_b, tot_f, tot_ms, per

            fwd_bytes, fwd_flops, fwd_ms, per_layer_gbs = layer_sums(times, True)
            inf_bytes, _, inf_ms, inf_gbs = layer_sums(itimes, False)
            gemm_ms = sum(sum(v) for v in times.values()) / args.steps
            per_class = {}
            for tag, v in times.items():
                per_class[tag.split(".")[0]] = per_class.get(tag.split(".")[0], 0.0) + sum(v) / args.steps
            hbm_ach = fwd_bytes / (fwd_ms * 1e-3) / 1e9 if fwd_ms else None
            inf_ach = inf_bytes / (inf_ms * 1e-3) / 1e9 if inf_ms else None
            tf_ach = fwd_flops / (fwd_ms * 1e-3) / 1e12 if fwd_ms else None
            traffic = None
            tpath = os.path.join(ROOT, "profiles", "traffic.json")
            if os.path.isfile(tpath):
                tj = json.load(open(tpath))
                traffic = tj.get("grcc_fwd_fused_dram_bytes_per_launch" if fused else "grcc_layer_fwd_dram_bytes_per_launch")
            nL = max(1, geomS.L)
            out["roofline"] = dict(
                bound="hbm", achieved=hbm_ach, peak=pk["hbm_gbs"], unit="GB/s",
                frac=(hbm_ach / pk["hbm_gbs"]) if hbm_ach else None, traffic=traffic,
                kernel=("grcc_fwd_kernel: one fused launch per GRCC layer (train mode: + saved tanh, sigmoid); algorithmic "
                        "bytes per SURVEY.md 8d (fp32 tensors: read x, cond; write x_next; RMW skip; + 2 D saved "
                        "activations), summed over the layers" if fused else
                        "tgemm_kernel: GRCC layer forward (2 launches/layer: conv+gate, res+skip), algorithmic bytes per "
                        "SURVEY.md 8d incl. saved tanh/sigmoid, summed over the layers"),
                algorithmic_bytes_per_launch=fwd_bytes / nL if fwd_bytes else None,
                # the same launch on the bytes it actually moves (ncu dram__bytes of one layer, profiles/traffic.json)
                frac_on_measured_traffic=(traffic / (fwd_ms / nL * 1e-3) / 1e9 / pk["hbm_gbs"]) if (traffic and fwd_ms) else None,
                ms_per_layer_fwd=fwd_ms / nL, per_layer_gbs=per_layer_gbs,
                frac_inference=(inf_ach / pk["hbm_gbs"]) if inf_ach else None, achieved_inference=inf_ach,
                ms_per_layer_inference=inf_ms / nL if inf_ms else None, per_layer_gbs_inference=inf_gbs,
                inference_note="forward only under torch.no_grad(): no saved activations; BASELINE.md 5's per-layer target "
                               "(0.199 ms on 783.9 MB for layer 0) sits at the TF32 tensor floor (~0.19 ms under the power "
                               "cap); the fp16-operand kernel's tensor floor is half of that",
                peak_source=pk["source"])
            out["roofline_tensor"] = dict(
                bound="tensor", achieved=tf_ach, peak=pk["bf16_tflops"], unit="TFLOP/s",
                frac=(tf_ach / pk["bf16_tflops"]) if tf_ach else None,
                note=("fp16 operands: the measured bf16 cuBLAS rate is the like-for-like denominator" if fused else
                      "TF32 operands: nominal peak = half of the bf16 rate, i.e. frac_of_tf32_peak = 2 x frac"),
                frac_of_tf32_peak=None if fused or not tf_ach else 2 * tf_ach / pk["bf16_tflops"])
            out["kernel_share"] = dict(tcgen05_ms_per_step=gemm_ms, step_ms=ms, instrumented_step_ms=prof_ms,
                                       outside_engines_frac=max(0.0, 1.0 - gemm_ms / prof_ms) if prof_ms else None,
                                       per_class_ms=per_class)
    # release this workload's workspaces before the next one
    del model, opt, sync, step, eager_step
    ops._plans.clear()
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3", "cfg5"],
                    help="cfg2 (default, the headline): WaveNet decoder train step; cfg3: full VQ-VAE-EMA autoencoder "
                         "train step (Encoder -> VQEMA -> WaveNet, par/arch.vqvae-ema.json), batch 16/GPU; with "
                         "--gpus N this is cfg4 (grads + EMA statistics in ONE all-reduce); cfg5: deep decoder stress "
                         "(30 dilation layers, 512 residual channels, window 65536, batch 2/GPU)")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: cfg2 8, cfg3 16)")
    ap.add_argument("--window", type=int, default=16384)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the eager-PyTorch/cuDNN baseline on this GPU")
    ap.add_argument("--no-extra", action="store_true",
                    help="multi-GPU runs of the default workload also time cfg4 and cfg5 (BASELINE.json's multi-GPU "
                         "configurations) and report them under `extra`; this switch skips that")
    ap.add_argument("--no-graph", action="store_true", help="run the timed steps eagerly instead of replaying a CUDA graph")
    ap.add_argument("--eager-cfg3", action="store_true",
                    help="cfg3 only: keep the reference's data-dependent diagnostics (unique()) and run the step eagerly; "
                         "default: VQEMA.static_diagnostics and the step replayed as a CUDA graph")
    ap.add_argument("--graph-cfg3", action="store_true", help="(default now; kept for older session scripts)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the aewn hot path has no CPU fallback; use --impl reference)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    out = bench_workload(args, args.workload, rank, local_rank, world, dev)
    if world > 1 and args.workload == "cfg2" and not args.no_extra:
        # BASELINE.json's multi-GPU configurations, under the same launch: cfg4 = the VQ-VAE-EMA step data-parallel with
        # [grads | z_sum | n_sum] in one all-reduce (chassis.py:168-171,187-190; vqema_bn.py:172-195), cfg5 = the deep stack
        extra = {}
        for name, wl in (("cfg4", "cfg3"), ("cfg5", "cfg5")):
            try:
                r = bench_workload(args, wl, rank, local_rank, world, dev, light=True)
                if rank == 0:
                    extra[name] = {k: r[k] for k in ("value", "unit", "ms_per_step", "n_gpus", "e2e", "gpu_launches", "clocks")}
                    extra[name]["workload"] = r["config"]["workload"]
                    extra[name]["step_execution"] = r["config"]["step_execution"]
            except Exception as e:   # noqa: BLE001
                if rank == 0:
                    extra[name] = dict(unavailable=f"{type(e).__name__}: {str(e)[:160]}")
        if rank == 0:
            out["extra"] = extra
    if rank == 0:
        cfg2 = args.workload == "cfg2"
        if world == 1 and cfg2 and not args.no_gpu_baseline:
            B, W = (args.batch or 8), args.window
            out["gpu_library_baseline"] = gpu_library_baseline(B, W, dev)
            tf = out["gpu_library_baseline"].get("tf32", {}).get("value")
            if tf:
                out["gpu_library_baseline"]["speedup_vs_tf32"] = out["value"] / tf
        if not args.no_cpu_baseline and world == 1 and cfg2:
            pick_cpu_threads()
            sps, ctimes = cpu_port_step(2, 2048, 2, 1)
            out["cpu_baseline"] = dict(value=sps, unit=UNIT, cores=torch.get_num_threads(), kind="port",
                                       sample="same decoder (oracle port of wavenet.py), batch 2 x window 2048, "
                                              "median of 2 steps after 1 warm-up (--impl reference runs the full "
                                              "configuration)")
        else:
            out["cpu_baseline"] = None
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
