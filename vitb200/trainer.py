"""Sync-free training step for the drop-in models: fused cross-entropy, in-engine backward, optional data-parallel
gradient all-reduce overlapped with backward, fused flat Adam (+ bf16 parameter refresh).

This is the B200-native equivalent of the reference's hot loop body (base.py:51-57 == vanilla_vit.py:233-239:
``zero_grad -> model(images) -> CrossEntropyLoss -> backward -> Adam.step``) without its two ``.item()`` host syncs
per step (base.py:59-62): the loss stays on the device and is only read when the caller asks for it.
"""
import torch

from . import ops


class FusedAdam:
    """torch.optim.Adam semantics (lr, betas, eps, weight_decay as L2) over the engine's flat buffers: one kernel."""

    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.engine = model._get_engine()
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.step_count = 0
        self.exp_avg = None
        self.exp_avg_sq = None

    def zero_grad(self, set_to_none=False):
        eng = self.engine
        eng.ensure_bound()
        eng.prepare_grads()
        eng.flat_grad.zero_()

    def step(self, grad_scale=1.0):
        eng = self.engine
        if self.exp_avg is None or self.exp_avg.device != eng.flat.device or self.exp_avg.numel() != eng.total:
            self.exp_avg = torch.zeros_like(eng.flat)
            self.exp_avg_sq = torch.zeros_like(eng.flat)
        self.step_count += 1
        ops.adam_step(eng.flat, eng.flat_grad, self.exp_avg, self.exp_avg_sq, eng.flat_bf16, lr=self.lr, beta1=self.betas[0],
                      beta2=self.betas[1], eps=self.eps, weight_decay=self.weight_decay, step=self.step_count, grad_scale=grad_scale)
        eng.bf16_fresh = True


class Trainer:
    """``loss = trainer.step(images, labels)``; images/labels may live on the host (pinned) or on the device."""

    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, reducer=None):
        self.model = model
        self.engine = model._get_engine()
        self.opt = FusedAdam(model, lr, betas, eps, weight_decay)
        self.reducer = reducer
        self._loss = None
        self._correct = None
        if reducer is not None:
            reducer.attach(self.engine)

    def step(self, images, labels):
        eng = self.engine
        if not images.is_cuda:
            images = images.cuda(non_blocking=True)
        if not labels.is_cuda:
            labels = labels.cuda(non_blocking=True)
        eng.ensure_bound()
        if self._loss is None or self._loss.device != eng.flat.device:
            self._loss = torch.zeros(1, device=eng.flat.device, dtype=torch.float32)
            self._correct = torch.zeros(1, device=eng.flat.device, dtype=torch.int32)
        self.opt.zero_grad()
        self._loss.zero_()
        outs, ws = eng.forward(images, training=True, want="logits")
        B = images.shape[0]
        world = self.reducer.world_size if self.reducer is not None else 1
        ops.cross_entropy(outs[0], labels, self._loss, weight=1.0 / B, dlogits_bf16=ws["dlogits"][0][:, :eng.C],
                          grad_scale=1.0 / world, correct_accum=self._correct)
        if self.reducer is not None:
            self.reducer.begin_step()
        eng.backward(ws, [None] * len(outs), want="logits", dlogits_ready=True)
        if self.reducer is not None:
            self.reducer.finish_step()
        self.opt.step()
        return self._loss
