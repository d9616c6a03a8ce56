"""Data-parallel step plumbing: ONE all-reduce per step over a single flat fp32 buffer
[ all parameter gradients | VQ-EMA z_sum | VQ-EMA n_sum | metric scalars ].

Mirrors what the reference gets from torch_xla on TPU (chassis.py:168-169 xm.optimizer_step = all-reduce(sum) of the
gradients scaled by 1/world; chassis.py:187-190 xm.all_reduce of [loss, tprb] scaled by 1/count), on NCCL over
NVLink 5 / NVSwitch (backend "nccl"; "gloo" in the CPU tests).  Gradients and metrics are averaged; the EMA code
statistics are TOTALS over all replicas, after which every rank applies the same EMA update
(vqema_bn.py:190-195) -- identical to the single-process result up to fp32 summation order."""
import weakref

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
        self._views = []
        off = 0
        for p in self.params:
            self._views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self._bind(copy=False)
        if vqema is not None:
            vqema.defer_ema = True
        # fused_accumulate=True: the decoder's backward adds its weight gradients into these persistent .grad buffers
        # with ONE launch instead of one clone + one add per parameter.  The opt-in is recorded on THESE parameters (a
        # weak reference to this object: it lapses when the sync object goes away) and is honoured per backward call
        # only when that call accumulates into the parameters (ops.fused_accumulate_applies), so a caller that also
        # takes torch.autograd.grad(...) through the decoder (mfcc_inverter.py:103 before chassis.py:157) stays correct.
        if fused_accumulate:
            from . import ops
            ops.mark_fused_accumulate(self.params, weakref.ref(self))

    def _bind(self, copy=True):
        """(Re-)alias every p.grad to its slice of the flat buffer.  optimizer.zero_grad() defaults to set_to_none=True
        (the reference's loop calls it, chassis.py:151) and any `p.grad = ...` re-assignment detaches a gradient from the
        buffer; the all-reduce would then carry stale zeros while the optimizer steps on local gradients.  A detached
        gradient's values are copied in (copy=True) before the alias is restored."""
        rebound = 0
        for p, v in zip(self.params, self._views):
            g = p.grad
            if g is not None and g.data_ptr() == v.data_ptr() and g.shape == v.shape:
                continue
            if g is not None and copy:
                v.copy_(g)
            elif g is None and copy:
                v.zero_()
            p.grad = v
            rebound += 1
        return rebound

    def zero_grad(self):
        self._bind(copy=False)
        self.flat[:self.n_grad].zero_()

    def sync(self, metrics=None):
        """Call after backward.  Returns the averaged metric scalars (tensor of n_metrics)."""
        self._bind(copy=True)          # gradients that were detached from the flat buffer since zero_grad() come back in
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
