"""Sync-free training step for the drop-in models: fused cross-entropy, in-engine backward, optional data-parallel
gradient all-reduce overlapped with backward, fused flat Adam (+ bf16 parameter refresh), replayed from one CUDA graph.

This is the B200-native equivalent of the reference's hot loop body (base.py:51-57 == vanilla_vit.py:233-239:
``zero_grad -> model(images) -> CrossEntropyLoss -> backward -> Adam.step``) without its two ``.item()`` host syncs
per step (base.py:59-62): the loss stays on the device and is only read when the caller asks for it.
"""
import os

import torch

from . import ops


class FusedAdam:
    """torch.optim.Adam semantics (lr, betas, eps, weight_decay as L2) over the engine's flat buffers: one kernel."""

    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.engine = model._get_engine()
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.exp_avg = None
        self.exp_avg_sq = None
        self.step_counter = None   # int32 device tensor: the step number lives on the GPU (graph-replayable)

    def _ensure_state(self):
        eng = self.engine
        frozen = [k for k, p in eng._order if not p.requires_grad]
        if frozen:
            raise NotImplementedError(f"vitb200 FusedAdam updates the whole flat parameter buffer; {len(frozen)} parameters are frozen "
                                      "(requires_grad=False) — use the autograd path with a torch.optim optimizer for partial fine-tuning")
        eng.ensure_bound()
        if self.exp_avg is None or self.exp_avg.device != eng.flat.device or self.exp_avg.numel() != eng.total:
            self.exp_avg = torch.zeros_like(eng.flat)
            self.exp_avg_sq = torch.zeros_like(eng.flat)
            self.step_counter = torch.zeros(1, device=eng.flat.device, dtype=torch.int32)

    def zero_grad(self, set_to_none=False):
        eng = self.engine
        eng.ensure_bound()
        eng.prepare_grads()
        eng.flat_grad.zero_()

    def step(self, grad_scale=1.0):
        eng = self.engine
        self._ensure_state()
        ops.adam_step(eng.flat, eng.flat_grad, self.exp_avg, self.exp_avg_sq, eng.flat_bf16, lr=self.lr, beta1=self.betas[0],
                      beta2=self.betas[1], eps=self.eps, weight_decay=self.weight_decay, step=0, grad_scale=grad_scale,
                      step_counter=self.step_counter)
        eng.bf16_fresh = True


class Trainer:
    """``loss = trainer.step(images, labels)``; images/labels may live on the host (pinned) or on the device.

    The first step for a given batch shape runs eagerly (allocations, kernel attribute setup), the second is captured
    into a CUDA graph, later steps copy the batch into the graph's static input buffers and replay it.  Call
    ``invalidate()`` after changing parameters from outside (load_state_dict, manual edits) so the bf16 shadow is
    refreshed."""

    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, reducer=None, use_cuda_graph=True,
                 distillation=None):
        """distillation: ``dict(teacher=module, type='hard'|'soft', alpha=..., tau=...)`` — the arguments of the reference's
        ``DistillationLoss`` (utils/distillation_loss.py:21-28; deit.py:33-35) for a two-head (DeiT-style) model: the teacher runs
        under ``no_grad`` on the same images ahead of the graph replay, the loss and both logits gradients come from one kernel."""
        self.model = model
        self.distillation = distillation
        if distillation is not None:
            assert distillation["type"] in ("soft", "hard"), "distillation type 'none' is the plain cross-entropy trainer"
            assert getattr(model._get_engine(), "two_heads", False), "distillation needs a model with a distillation head"
        self._teacher_logits = None
        self.engine = model._get_engine()
        self.opt = FusedAdam(model, lr, betas, eps, weight_decay)
        self.reducer = reducer
        # Data parallel: the NCCL all-reduces on the side stream are captured into the same CUDA graph as the kernels (round 2, 2 GPUs:
        # 31.50 vs 31.86 ms/step with eager launches).  VITB200_DP_GRAPH=0 launches the data-parallel step eagerly.  A process that
        # captured collectives must drop its graphs before destroying the process group (Trainer.close()).
        self.use_cuda_graph = use_cuda_graph and (reducer is None or os.environ.get("VITB200_DP_GRAPH", "1") != "0")
        self._loss = None
        self._correct = None
        self._graphs = {}   # batch shape -> (graph, static_images, static_labels)
        self._eager_done = set()
        if reducer is not None:
            reducer.attach(self.engine)

    def close(self):
        """Drops the captured graphs (they hold NCCL kernels of the reducer's communicator): call before destroy_process_group()."""
        import gc
        torch.cuda.synchronize()
        self._graphs.clear()
        gc.collect()
        torch.cuda.synchronize()

    def invalidate(self):
        self.engine.bf16_fresh = False
        self._graphs.clear()
        self._eager_done.clear()

    @property
    def correct_count(self):
        """int32 device tensor: number of argmax-correct samples in the last step (no host sync)."""
        return self._correct

    def _step_body(self, images, labels):
        eng = self.engine
        n0 = ops.LAUNCHES["n"]
        self.opt.zero_grad()
        self._loss.zero_()
        self._correct.zero_()
        m = self.model
        rates = (getattr(m, "dropout", getattr(m, "drop_rate", 0.0)), getattr(m, "attention_dropout", getattr(m, "attn_drop_rate", 0.0)))
        eng.p_drop, eng.p_attn = (float(rates[0]), float(rates[1])) if m.training else (0.0, 0.0)
        outs, ws = eng.forward(images, training=True, want="logits")
        B = images.shape[0]
        world = self.reducer.world_size if self.reducer is not None else 1
        if self.distillation is not None:
            dl = self.distillation
            ops.distill_loss(outs[0], outs[1], self._teacher_logits, labels, self._loss, kind=dl["type"], alpha=dl["alpha"],
                             tau=dl.get("tau", 1.0), dlogits_bf16=ws["dlogits"][0][:, :eng.C], dlogits_kd_bf16=ws["dlogits"][1][:, :eng.C],
                             grad_scale=1.0 / world, correct_accum=self._correct)
        else:
            ops.cross_entropy(outs[0], labels, self._loss, weight=1.0 / B, dlogits_bf16=ws["dlogits"][0][:, :eng.C],
                              grad_scale=1.0 / world, correct_accum=self._correct)
        if self.reducer is not None:
            self.reducer.begin_step()
        eng.backward(ws, [None] * len(outs), want="logits", dlogits_ready=True)
        if self.reducer is not None:
            self.reducer.finish_step()
        self.opt.step()
        self.launches_per_step = ops.LAUNCHES["n"] - n0   # libvitb200 kernels per step (graph replays launch the same set)

    def step(self, images, labels):
        eng = self.engine
        eng.ensure_bound()
        dev = eng.flat.device
        self.opt._ensure_state()
        if self._loss is None or self._loss.device != dev:
            self._loss = torch.zeros(1, device=dev, dtype=torch.float32)
            self._correct = torch.zeros(1, device=dev, dtype=torch.int32)
        key = (tuple(images.shape), str(images.dtype))
        if self.distillation is not None:
            # the teacher is a foreign nn.Module (deit.py:32: a CNN): it runs eagerly, outside the graph, into a static buffer
            with torch.no_grad():
                t = self.distillation["teacher"](images if images.is_cuda else images.to(dev, non_blocking=True)).float()
            if self._teacher_logits is None or self._teacher_logits.shape != t.shape:
                self._teacher_logits = torch.empty_like(t)
                self._graphs.clear()
            self._teacher_logits.copy_(t)
        if not self.use_cuda_graph:
            if not images.is_cuda:
                images = images.to(dev, non_blocking=True)
            if not labels.is_cuda:
                labels = labels.to(dev, non_blocking=True)
            self._step_body(images.float() if images.dtype != torch.float32 else images, labels)
            return self._loss
        entry = self._graphs.get(key)
        if entry is None:
            s_img = torch.empty(images.shape, device=dev, dtype=torch.float32)
            s_lab = torch.empty(labels.shape, device=dev, dtype=torch.int64)
            s_img.copy_(images, non_blocking=True)
            s_lab.copy_(labels, non_blocking=True)
            if key not in self._eager_done:
                self._eager_done.add(key)          # first step: eager (lazy allocations, cudaFuncSetAttribute, NCCL init)
                self._step_body(s_img, s_lab)
                return self._loss
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                self._step_body(s_img, s_lab)
            entry = (graph, s_img, s_lab)
            self._graphs[key] = entry
            graph.replay()
            return self._loss
        graph, s_img, s_lab = entry
        s_img.copy_(images, non_blocking=True)
        s_lab.copy_(labels, non_blocking=True)
        graph.replay()
        return self._loss
