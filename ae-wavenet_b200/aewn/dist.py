"""Data-parallel step plumbing: ONE all-reduce per step over a single flat fp32 buffer
[ all parameter gradients | VQ-EMA z_sum | VQ-EMA n_sum | metric scalars ].

Mirrors what the reference gets from torch_xla on TPU (chassis.py:168-169 xm.optimizer_step = all-reduce(sum) of the
gradients scaled by 1/world; chassis.py:187-190 xm.all_reduce of [loss, tprb] scaled by 1/count), on NCCL over
NVLink 5 / NVSwitch (backend "nccl"; "gloo" in the CPU tests).  Gradients and metrics are averaged; the EMA code
statistics are TOTALS over all replicas, after which every rank applies the same EMA update
(vqema_bn.py:190-195) -- identical to the single-process result up to fp32 summation order."""
import torch
import torch.distributed as dist


class FlatGradSync:
    def __init__(self, params, vqema=None, n_metrics=2, process_group=None, fused_accumulate=False):
        self.params = [p for p in params if p.requires_grad]
        self.vqema = vqema
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        dev = self.params[0].device
        self.n_grad = sum(p.numel() for p in self.params)
        self.n_z = vqema.k * vqema.d if vqema is not None else 0
        self.n_n = vqema.k if vqema is not None else 0
        self.n_metrics = n_metrics
        self.flat = torch.zeros(self.n_grad + self.n_z + self.n_n + n_metrics, device=dev)
        # gradients live INSIDE the flat buffer: backward writes/accumulates straight into it, no gather copy
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        if vqema is not None:
            vqema.defer_ema = True
        # fused_accumulate=True: the decoder's backward adds its weight gradients into these persistent .grad buffers
        # with ONE launch instead of one clone + one add per parameter (ops.ACCUMULATE_INTO_GRAD).  Only for steps that
        # run exactly one loss.backward() per forward: a caller that ALSO takes torch.autograd.grad(...) through the
        # decoder (mfcc_inverter.py:103 before chassis.py:157) would have the weight gradients added twice, because a
        # custom Function cannot tell which of its input gradients a particular backward call asks for.
        if fused_accumulate:
            from . import ops
            ops.ACCUMULATE_INTO_GRAD = True

    def zero_grad(self):
        self.flat[:self.n_grad].zero_()

    def sync(self, metrics=None):
        """Call after backward.  Returns the averaged metric scalars (tensor of n_metrics)."""
        o = self.n_grad
        if self.vqema is not None:
            self.flat[o:o + self.n_z] = self.vqema.z_sum.reshape(-1)
            self.flat[o + self.n_z:o + self.n_z + self.n_n] = self.vqema.n_sum
        m0 = o + self.n_z + self.n_n
        if metrics is not None:
            self.flat[m0:m0 + self.n_metrics] = torch.stack([torch.as_tensor(m, device=self.flat.device).float().reshape(())
                                                             for m in metrics])
        if self.world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat[:self.n_grad].mul_(1.0 / self.world)
            self.flat[m0:].mul_(1.0 / self.world)
        if self.vqema is not None:
            z_tot = self.flat[o:o + self.n_z].view(self.vqema.k, self.vqema.d)
            n_tot = self.flat[o + self.n_z:o + self.n_z + self.n_n]
            self.vqema.apply_ema(z_tot, n_tot)
        return self.flat[m0:m0 + self.n_metrics]
