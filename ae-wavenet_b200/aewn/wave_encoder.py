"""Drop-in replacements for the reference's wave_encoder.py (ConvReLURes :8-50, Encoder :53-103) on the kernel path.

Each layer is one tcgen05 time-major GEMM over the k tap-shifted views of its input with the bias, ReLU, residual add
and zero-activation count fused into the epilogue (ops.tap_conv).  The backward pass uses the out-of-place form
act = relu(pre) + x[...] (the reference's in-place add breaks autograd)."""
import torch
from torch import nn

from . import ops
from .compat import vconv, xavier_init
from .wavenet import _require_cuda


class ConvReLURes(nn.Module):
    def __init__(self, n_in_chan, n_out_chan, filter_sz, stride=1, do_res=True, parent_vc=None, name=None):
        super().__init__()
        self.n_in = n_in_chan
        self.n_out = n_out_chan
        self.conv = nn.Conv1d(n_in_chan, n_out_chan, filter_sz, stride, padding=0, bias=True)
        self.relu = nn.ReLU()
        self.name = name
        self.stride = stride
        self.vc = vconv.VirtualConv(filter_info=filter_sz, stride=stride, parent=parent_vc, name=name)
        self.do_res = do_res
        if self.do_res:
            if stride != 1:
                import sys
                print("Stride must be 1 for residually connected convolution", file=sys.stderr)
                raise ValueError
            l_off, r_off = vconv.output_offsets(self.vc, self.vc)
            self.register_buffer("residual_offsets", torch.tensor([l_off, r_off]))
            self._res_lw = int(l_off)
            if n_in_chan != n_out_chan:
                raise ValueError("residual connection needs n_in_chan == n_out_chan")
        xavier_init(self.conv)

    def forward(self, x):
        _require_cuda(x)
        count = torch.zeros(1, dtype=torch.int64, device=x.device)
        act = ops.tap_conv(x, self.conv.weight, self.conv.bias, stride=self.stride, mode=2 if self.do_res else 1,
                           res_lw=self._res_lw if self.do_res else 0, zero_count=count)
        # wave_encoder.py:46 -- fraction of exact zeros AFTER the residual add; stays on the device (no sync)
        self.frac_zero_act = count[0].double() / act.nelement()
        return act


class Encoder(nn.Module):
    def __init__(self, n_in, n_out, parent_vc):
        super().__init__()
        stack_in_chan = [n_in] + [n_out] * 8
        stack_filter_sz = [3, 3, 4, 3, 3, 1, 1, 1, 1]
        stack_strides = [1, 1, 2, 1, 1, 1, 1, 1, 1]
        stack_residual = [False, True, False, True, True, True, True, True, True]
        self.net = nn.Sequential()
        self.vc = dict()
        for i, (in_chan, filt_sz, stride, do_res) in enumerate(zip(stack_in_chan, stack_filter_sz, stack_strides,
                                                                  stack_residual)):
            name = "CRR_{}(filter_sz={}, stride={}, do_res={})".format(i, filt_sz, stride, do_res)
            mod = ConvReLURes(in_chan, n_out, filt_sz, stride, do_res, parent_vc, name)
            self.net.add_module(str(i), mod)
            parent_vc = mod.vc
        self.vc["beg"] = self.net[0].vc
        self.vc["end"] = self.net[-1].vc

    def set_parent_vc(self, parent_vc):
        self.vc["beg"].parent = parent_vc
        parent_vc.child = self.vc["beg"]

    def update_metrics(self):
        self.metrics = {"enc_az_{}".format(i): mod.frac_zero_act for i, mod in enumerate(self.net)}

    def forward(self, mels):
        out = self.net(mels)
        self.update_metrics()
        return out
