"""Input pipeline step in front of the hot path: pinned-memory, asynchronous host->device staging.

The reference iterates a synchronous, un-pinned DataLoader and moves every batch with a blocking ``.to(device)``
(utils/load_data.py:33-35, base.py:51).  ``DevicePrefetcher`` wraps any iterable of ``(images, labels)`` host batches:
batch i+1 is copied into one of two device buffers on a side stream (through a pinned staging buffer when the source
is pageable) while batch i is being consumed, and the consumer's stream only waits on the copy's event — the same
double-buffering bench.py uses for its end-to-end leg.  The fp32 -> bf16 cast the reference would need under autocast
is not done here: the patch-embedding kernel reads fp32 pixels and casts on the fly (vb_patchify).
"""
import torch


class DevicePrefetcher:
    def __init__(self, loader, device="cuda", depth=2):
        self.loader, self.device, self.depth = loader, torch.device(device), max(2, int(depth))
        self.stream = torch.cuda.Stream(device=self.device)
        self._slots = None   # per slot: [pinned images, pinned labels, device images, device labels, ready event, consumed event, used]

    def __len__(self):
        return len(self.loader)

    def _alloc(self, images, labels):
        self._slots = []
        for _ in range(self.depth):
            self._slots.append([torch.empty(images.shape, dtype=images.dtype).pin_memory(), torch.empty(labels.shape, dtype=labels.dtype).pin_memory(),
                                torch.empty(images.shape, dtype=images.dtype, device=self.device),
                                torch.empty(labels.shape, dtype=labels.dtype, device=self.device), torch.cuda.Event(), torch.cuda.Event(), False])

    def _stage(self, slot, images, labels):
        pin_i, pin_l, dev_i, dev_l, ready, consumed, used = slot
        if images.shape != dev_i.shape or labels.shape != dev_l.shape or images.dtype != dev_i.dtype:   # ragged last batch
            with torch.cuda.stream(self.stream):
                out = (images.to(self.device, non_blocking=True), labels.to(self.device, non_blocking=True))
                ev = torch.cuda.Event()
                ev.record(self.stream)
            # allocated on the side stream, consumed on the caller's: the caching allocator must not recycle them underneath it
            cur = torch.cuda.current_stream(self.device)
            out[0].record_stream(cur)
            out[1].record_stream(cur)
            return out[0], out[1], ev
        # The previous H2D copy out of this slot's pinned buffers is asynchronous and is itself held back on the GPU until the
        # consumer released the device buffer (wait_event(consumed) below), so with a sync-free trainer the host can be several
        # batches ahead: wait until that copy has really left the pinned memory before overwriting it.
        if used:
            ready.synchronize()
        slot[6] = True
        src_i = images if images.is_pinned() else pin_i.copy_(images)
        src_l = labels if labels.is_pinned() else pin_l.copy_(labels)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(consumed)          # the step that last used this slot has been enqueued and finished reading it
            dev_i.copy_(src_i, non_blocking=True)
            dev_l.copy_(src_l, non_blocking=True)
            ready.record(self.stream)
        return dev_i, dev_l, ready

    def __iter__(self):
        it = iter(self.loader)
        try:
            nxt = next(it)
        except StopIteration:
            return
        if self._slots is None:
            self._alloc(nxt[0], nxt[1])
        cur_stream = torch.cuda.current_stream(self.device)
        for s in self._slots:
            s[5].record(cur_stream)
        i = 0
        staged = self._stage(self._slots[0], nxt[0], nxt[1])
        while staged is not None:
            slot = self._slots[i % self.depth]
            try:
                nxt = next(it)
                nxt_staged = self._stage(self._slots[(i + 1) % self.depth], nxt[0], nxt[1])
            except StopIteration:
                nxt_staged = None
            dev_i, dev_l, ready = staged
            cur_stream = torch.cuda.current_stream(self.device)
            cur_stream.wait_event(ready)
            yield dev_i, dev_l
            slot[5].record(torch.cuda.current_stream(self.device))   # consumer work on this batch has been enqueued
            staged = nxt_staged
            i += 1
