"""Data-parallel gradient reduction: bucketed NCCL all-reduce over NVLink/NVSwitch, overlapped with backward.

The reference trains on a single device (base.py:51) — DP is a requirement of BASELINE.json, and
its comparable library baseline is torch DistributedDataParallel.  Here the engine's flat gradient buffer is laid
out in gradient-production order (heads, blocks L-1..0, embedding), so a bucket is a contiguous slice: no flatten
/ unflatten copies.  As soon as the backward pass finishes a segment the engine calls ``_on_segment``; when the
open bucket is large enough an event is recorded on the compute stream and ``all_reduce(SUM)`` of that slice is
enqueued on a side stream.  The loss gradient is pre-scaled by 1/world, so SUM yields the mean gradient.
"""
import torch
import torch.distributed as dist


class GradReducer:
    def __init__(self, process_group=None, bucket_bytes=48 << 20, overlap=True):
        """overlap=True: buckets are all-reduced on a side stream while backward is still running (the engine calls back per finished
        segment; the step is then a single stream of eager launches or ONE captured graph).  overlap=False: one all-reduce of the whole
        gradient buffer after backward — for small, launch-bound models (the DETR encoder at 4 images per GPU), whose forward / backward
        are replayed from their own CUDA graphs and must not be interleaved with Python callbacks."""
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised before creating a GradReducer")
        self.pg = process_group
        self.world_size = dist.get_world_size(process_group)
        self.bucket_elems = bucket_bytes // 4
        self.overlap = overlap
        self.engine = None
        self.comm_stream = None
        self._works = []
        self._bucket_start = None
        self._bucket_end = None
        self._last_segment = None

    def attach(self, engine):
        self.engine = engine
        engine.grad_segment_hook = self._on_segment if self.overlap else None
        self._last_segment = len(engine.segment_bounds) - 1
        self._synced = False
        self.sync_parameters(strict=False)

    def sync_parameters(self, strict=True):
        """Rank 0's parameters become everybody's (what DistributedDataParallel does at construction): without it ranks built from
        different seeds / checkpoints would silently train different replicas.  Deferred to the first step when the model is not on
        its device yet."""
        eng = self.engine
        if hasattr(eng, "ensure_bound"):
            if not eng._order[0][1].is_cuda and not strict:
                return
            eng.ensure_bound()
        src = dist.get_global_rank(self.pg, 0) if self.pg is not None else 0
        dist.broadcast(eng.flat, src=src, group=self.pg)
        eng.bf16_fresh = False      # the bf16 shadow must be re-cast from the broadcast values
        self._synced = True

    def begin_step(self):
        if not self._synced:
            self.sync_parameters()
        self._works = []
        self._bucket_start = None
        if self.comm_stream is None and self.engine.flat.is_cuda:
            self.comm_stream = torch.cuda.Stream(device=self.engine.flat.device, priority=-1)

    def _flush(self):
        if self._bucket_start is None:
            return
        eng = self.engine
        sl = eng.flat_grad[self._bucket_start:self._bucket_end]
        if sl.is_cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ev)
                self._works.append(dist.all_reduce(sl, op=dist.ReduceOp.SUM, group=self.pg, async_op=True))
        else:  # gloo / CPU tests of the bucketing logic
            self._works.append(dist.all_reduce(sl, op=dist.ReduceOp.SUM, group=self.pg, async_op=True))
        self._bucket_start = None

    def _on_segment(self, idx):
        start, end = self.engine.segment_bounds[idx]
        if self._bucket_start is None:
            self._bucket_start = start
        self._bucket_end = end
        if self._bucket_end - self._bucket_start >= self.bucket_elems or idx == self._last_segment:
            self._flush()

    def finish_step(self):
        if not self.overlap:     # everything at once, on the compute stream
            dist.all_reduce(self.engine.flat_grad, op=dist.ReduceOp.SUM, group=self.pg)
            return
        self._flush()
        for w in self._works:
            w.wait()
        self._works = []
