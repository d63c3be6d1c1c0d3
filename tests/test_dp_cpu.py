"""world_size-2 gloo test of the data-parallel bucketing logic (GradReducer) on CPU tensors."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _FakeEngine:
    """Minimal stand-in exposing what GradReducer uses: flat_grad, segment_bounds, grad_segment_hook."""

    def __init__(self, sizes):
        self.segment_bounds, off = [], 0
        for s in sizes:
            self.segment_bounds.append((off, off + s))
            off += s
        self.flat = torch.zeros(off)
        self.flat_grad = torch.zeros(off)
        self.grad_segment_hook = None


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vitb200.dp import GradReducer
    eng = _FakeEngine([1000, 5000, 300, 7000, 64])
    red = GradReducer(bucket_bytes=4 * 5500)
    red.attach(eng)
    flushed = []
    orig = red._flush

    def spy():
        if red._bucket_start is not None:
            flushed.append((red._bucket_start, red._bucket_end))
        orig()

    red._flush = spy
    for step in range(2):
        red.begin_step()
        for i, (a, b) in enumerate(eng.segment_bounds):           # backward produces segments in order
            eng.flat_grad[a:b] = float(rank + 1) * (i + 1) + step
            eng.grad_segment_hook(i)
        red.finish_step()
        expect = torch.cat([torch.full((b - a,), sum(float(r + 1) * (i + 1) + step for r in range(world)))
                            for i, (a, b) in enumerate(eng.segment_bounds)])
        assert torch.equal(eng.flat_grad, expect), (rank, step)
    # overlap=False (launch-bound models): no per-segment callbacks, one all-reduce of the whole buffer after backward
    eng2 = _FakeEngine([100, 28])
    red2 = GradReducer(overlap=False)
    red2.attach(eng2)
    assert eng2.grad_segment_hook is None
    red2.begin_step()
    eng2.flat_grad[:] = float(rank + 1)
    red2.finish_step()
    assert torch.equal(eng2.flat_grad, torch.full((128,), float(sum(r + 1 for r in range(world)))))
    q.put((rank, flushed))
    dist.destroy_process_group()


def test_bucketed_allreduce_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # buckets are contiguous slices in production order; identical on both ranks (deterministic bucket order)
    assert res[0][1] == res[1][1]
    b = res[0][1][:3]
    assert b == [(0, 6000), (6000, 13300), (13300, 13364)]
