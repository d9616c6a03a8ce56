"""Training-step plumbing on CUDA streams and graphs.

`GraphedStep` captures one whole optimisation step -- gradient zeroing, forward, loss, backward, gradient all-reduce,
optimizer -- into a CUDA graph and replays it.  A WaveNet train step is ~500 kernel launches (ours through ctypes, the
conditioning front-end through cuDNN, the optimizer's multi-tensor kernels); run eagerly, the stretches of small
kernels (front-end backward, optimizer, gradient bookkeeping) leave the GPU waiting for the Python interpreter.  The
reference's analogue is the XLA step graph it gets on TPU (chassis.py:168-169); on GPU it runs eagerly.

The captured callable must be sync-free and shape-static: no .item(), no data-dependent shapes (VQEMA.forward calls
unique() for a diagnostic, so the VQ-VAE step is not capturable as is), inputs are copied into fixed device buffers.
All aewn kernels qualify: they take their stream from torch.cuda.current_stream() at call time, never allocate and
never synchronise (include/aewn.h)."""
import torch


class GraphedStep:
    def __init__(self, step_fn, example_inputs, warmup=3, capture_on_warmup_stream=False):
        """step_fn(*tensors) -> loss tensor (0-dim).  example_inputs: device tensors fixing shapes / dtypes.
        The caller must not hold results of earlier eager calls of step_fn (a live autograd graph keeps AccumulateGrad
        nodes bound to the stream they were created on, which invalidates the capture)."""
        self.step_fn = step_fn
        self.static_in = [t.clone() for t in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                      # warm-up off the default stream (PyTorch's capture recipe):
            for _ in range(warmup):                        # builds plans, autotunes cuDNN, creates optimizer state
                step_fn(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        # capture_on_warmup_stream: capture on the SAME side stream as the warm-up, so that autograd's AccumulateGrad nodes
        # created during the warm-up and kept alive by module attributes that hold graph tensors (VQEMA.ze, .min_dist,
        # AutoEncoder.encoding_bn) match the capturing stream; by default torch.cuda.graph picks a stream of its own
        with (torch.cuda.graph(self.graph, stream=side) if capture_on_warmup_stream else torch.cuda.graph(self.graph)):
            self.static_loss = step_fn(*self.static_in)
        torch.cuda.synchronize()

    def __call__(self, *inputs):
        """inputs: host (pinned) or device tensors of the captured shapes.  Returns the (static) loss tensor."""
        for dst, src in zip(self.static_in, inputs):
            if src is not dst:
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_loss
