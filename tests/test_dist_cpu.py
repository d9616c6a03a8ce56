"""world_size-2 gloo test of the flat-buffer gradient / EMA-statistics all-reduce (host logic only; no CUDA)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakeVQ:
    """stands in for VQEMA on CPU: the sync logic only touches k, d, z_sum, n_sum, defer_ema, apply_ema"""

    def __init__(self, k, d):
        self.k, self.d = k, d
        self.z_sum, self.n_sum = torch.zeros(k, d), torch.zeros(k)
        self.numer, self.denom = torch.ones(k, d), torch.ones(k)
        self.defer_ema = False

    def apply_ema(self, z, n):
        self.numer = 0.99 * self.numer + 0.01 * z
        self.denom = 0.99 * self.denom + 0.01 * n


def worker(rank, world, port, out):
    sys.path.insert(0, os.path.join(ROOT, "ae-wavenet_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from aewn.dist import FlatGradSync
    torch.manual_seed(0)
    lin = torch.nn.Linear(4, 3)
    vq = FakeVQ(5, 2)
    sync = FlatGradSync(lin.parameters(), vqema=vq, n_metrics=2)
    assert vq.defer_ema
    x = torch.full((2, 4), float(rank + 1))
    sync.zero_grad()
    loss = lin(x).sum()
    loss.backward()                      # accumulates INTO the flat buffer views
    vq.z_sum[:] = rank + 1
    vq.n_sum[:] = 10 * (rank + 1)
    m = sync.sync(metrics=[loss.detach(), torch.tensor(float(rank))])
    torch.save(dict(wgrad=lin.weight.grad.clone(), numer=vq.numer.clone(), denom=vq.denom.clone(), m=m.clone()),
               os.path.join(out, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_flat_allreduce_world2(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    # d(sum(Wx+b))/dW = sum over batch of x -> rank r: 2*(r+1) per entry; averaged over 2 ranks = 3
    assert torch.allclose(r0["wgrad"], torch.full((3, 4), 3.0)) and torch.equal(r0["wgrad"], r1["wgrad"])
    # EMA statistics are TOTALS: z 1+2 = 3, n 10+20 = 30
    assert torch.allclose(r0["numer"], torch.full((5, 2), 0.99 + 0.03)) and torch.equal(r0["numer"], r1["numer"])
    assert torch.allclose(r0["denom"], torch.full((5,), 0.99 + 0.30))
    assert torch.allclose(r0["m"][1], torch.tensor(0.5)) and torch.equal(r0["m"], r1["m"])


def test_flat_sync_rebinds_detached_gradients():
    """optimizer.zero_grad(set_to_none=True) -- the reference loop's call (chassis.py:151) -- or a p.grad re-assignment
    detaches gradients from the flat all-reduce buffer; sync() / zero_grad() must copy them back in and re-alias, or the
    all-reduce would carry stale zeros while the optimizer steps on local gradients."""
    sys.path.insert(0, os.path.join(ROOT, "ae-wavenet_b200"))
    from aewn.dist import FlatGradSync
    torch.manual_seed(0)
    lin = torch.nn.Linear(4, 3)
    sync = FlatGradSync(lin.parameters())
    opt = torch.optim.SGD(lin.parameters(), lr=0.1)
    opt.zero_grad()                                   # set_to_none=True: p.grad is None now
    assert lin.weight.grad is None
    lin(torch.ones(2, 4)).sum().backward()            # autograd allocates fresh, detached .grad tensors
    assert lin.weight.grad.data_ptr() != sync._views[0].data_ptr()
    sync.sync()
    assert lin.weight.grad.data_ptr() == sync._views[0].data_ptr()
    assert torch.allclose(sync.flat[:12].view(3, 4), torch.full((3, 4), 2.0))
    assert torch.allclose(sync.flat[12:15], torch.full((3,), 2.0))
    sync.zero_grad()
    assert float(lin.weight.grad.abs().max()) == 0.0 and lin.weight.grad.data_ptr() == sync._views[0].data_ptr()
    lin.weight.grad = torch.ones(3, 4)                # explicit re-assignment
    sync.sync()
    assert lin.weight.grad.data_ptr() == sync._views[0].data_ptr() and float(sync.flat[:12].min()) == 1.0


def test_fused_accumulate_mark_is_per_parameter_set_and_lapses_with_the_sync_object():
    sys.path.insert(0, os.path.join(ROOT, "ae-wavenet_b200"))
    import gc
    from aewn import ops
    from aewn.dist import FlatGradSync
    a, b = torch.nn.Linear(2, 2), torch.nn.Linear(2, 2)
    sync = FlatGradSync(a.parameters(), fused_accumulate=True)
    assert all(getattr(p, "_aewn_fused_owner", None) is not None for p in a.parameters())
    assert all(getattr(p, "_aewn_fused_owner", None) is None for p in b.parameters())
    assert not hasattr(ops, "ACCUMULATE_INTO_GRAD")                  # no process-global switch
    # outside a backward pass the engine query raises -> the fused path does not apply
    assert ops.fused_accumulate_applies(list(a.parameters())) is False
    del sync
    gc.collect()
    assert all(p._aewn_fused_owner() is None for p in a.parameters())
