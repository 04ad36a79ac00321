"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: tile sharding and the variable-length all-gather."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from treelearn_b200.dist import allgather_rows, allreduce_grads, grad_buckets, shard_indices


def test_shard_indices_is_a_balanced_partition():
    w = [10, 1, 7, 7, 3, 9, 2, 8]
    parts = [shard_indices(w, r, 3) for r in range(3)]
    assert sorted(sum(parts, [])) == list(range(len(w)))
    loads = [sum(w[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= max(w)
    assert shard_indices(w, 0, 1) == list(range(len(w)))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    n = 3 + 4 * rank                      # ragged: rank 0 -> 3 rows, rank 1 -> 7 rows
    t = torch.arange(n * 2, dtype=torch.float32).reshape(n, 2) + 100 * rank
    out = allgather_rows(t)
    empty = allgather_rows(torch.zeros((0, 2)) if rank == 0 else t)     # an empty contribution
    q.put((rank, out.tolist(), empty.shape[0]))
    dist.destroy_process_group()


def test_allgather_rows_gloo_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = (torch.arange(6, dtype=torch.float32).reshape(3, 2).tolist()
              + (torch.arange(14, dtype=torch.float32).reshape(7, 2) + 100).tolist())
    for rank, out, n_empty in res:
        assert out == expect            # identical on both ranks, rank order preserved
        assert n_empty == 7


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from treelearn_b200 import TreeLearn
    torch.manual_seed(0)
    net = TreeLearn(channels=8, num_blocks=3, fixed_modules=['semantic_linear'])
    for i, p in enumerate(net.parameters()):
        if p.requires_grad and i % 5 != 0:                       # leave some grads None: treated as zeros
            p.grad = torch.full_like(p, float(rank + 1) * (1 + i % 3))
    allreduce_grads(net)
    ok = True
    for i, p in enumerate(net.parameters()):
        if not p.requires_grad:
            ok &= p.grad is None
        elif i % 5 != 0:
            ok &= bool(torch.allclose(p.grad, torch.full_like(p, 1.5 * (1 + i % 3))))    # mean of ranks {1,2} x factor
        else:
            ok &= bool((p.grad == 0).all())
    q.put((rank, ok))
    dist.destroy_process_group()


def test_allreduce_grads_gloo_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]


def test_grad_buckets_cover_trainable_parameters_once():
    from treelearn_b200 import TreeLearn
    net = TreeLearn(channels=8, num_blocks=4, fixed_modules=['offset_linear'])
    buckets = grad_buckets(net)
    flat = [id(p) for b in buckets for p in b]
    want = [id(p) for p in net.parameters() if p.requires_grad]
    assert sorted(flat) == sorted(want) and len(set(flat)) == len(flat)
    assert len(buckets) == 5                                   # 4 U-Net levels + the non-U-Net bucket


class _Toy(torch.nn.Module):
    """Parameter names shaped like the U-Net's (`unet.` + one `.u.` hop per level) so grad_buckets splits them by level."""

    def __init__(self):
        super().__init__()
        lin = lambda: torch.nn.Linear(4, 4)   # noqa: E731
        self.input_conv = lin()
        self.unet = torch.nn.Module()
        self.unet.blocks = lin()
        self.unet.u = torch.nn.Module()
        self.unet.u.blocks = lin()
        self.unet.u.unused = lin()            # never used in forward: its gradient stays None (= zeros in the bucket)
        self.unet.blocks_tail = lin()

    def forward(self, x):
        x = self.unet.blocks(self.input_conv(x))
        return self.unet.blocks_tail(x + self.unet.u.blocks(x)).sum()


def _overlap_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from treelearn_b200.dist import OverlappedGradReducer
    torch.manual_seed(0)
    a, b = _Toy(), _Toy()
    b.load_state_dict(a.state_dict())
    x = torch.randn(5, 4, generator=torch.Generator().manual_seed(10 + rank))
    red = OverlappedGradReducer(a)
    ok = True
    for it in range(2):                       # two steps: the reducer resets itself
        for m in (a, b):
            for p in m.parameters():
                p.grad = None
        a(x).backward()
        launched = red.launched_in_backward
        red.finish()
        b(x).backward()
        allreduce_grads(b)
        # the level-1 bucket holds the unused layer, so it can only be launched by finish(); the other two complete in backward
        ok &= launched == 2
        for (n, p), (_, r) in zip(a.named_parameters(), b.named_parameters()):
            ok &= p.grad is not None and bool(torch.allclose(p.grad, r.grad, rtol=0, atol=1e-7))
            if 'unused' in n:
                ok &= bool((p.grad == 0).all())
    red.remove()
    q.put((rank, ok))
    dist.destroy_process_group()


def test_overlapped_grad_reducer_matches_allreduce_grads_gloo_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_overlap_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]
