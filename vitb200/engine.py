"""Forward/backward engine for the pre-norm ViT / DeiT encoder stack on one B200.

The engine owns (a) a flat fp32 parameter buffer and a flat fp32 gradient buffer that the module's
``nn.Parameter``s are views of (laid out in gradient-production order so that data-parallel buckets are
contiguous slices), (b) a bf16 shadow of the parameters for the tensor-core operands, and (c) per-batch-size
activation workspaces.  It sequences the C-ABI kernels of libvitb200.so on the current CUDA stream; there is no
PyTorch arithmetic on the path (torch is used for allocation, streams and the autograd hook only).

Reference dataflow being reproduced: ViT.forward_features / Encoder.forward / EncoderBlock.forward
(vanilla_vit.py:186-207, 102-106, 73-83) and their autograd; SURVEY.md Appendix C lists what is saved.
"""
import torch

from . import ops

import os
_INFER_SPLIT_K = os.environ.get("VITB200_INFER_SPLIT_K", "0") == "1"
# In-projection bias gradient (column sums of dq | dk | dv).  "id" (default): the key part is exactly zero (softmax ignores a constant
# key offset) and, without attention dropout, the value part equals the column sum of dO (dV = P^T dO and every row of P sums to
# one), so two bandwidth-bound passes over L2-hot [M, D] matrices (dq, dO) replace work on the attention kernel's critical path
# (the in-kernel sums cost 60 us of a 240 us launch).  "1": in-kernel sums; "0": three column-sum passes over dq, dk, dv.
_QKV_COLSUM_MODE = os.environ.get("VITB200_FUSED_QKV_COLSUM", "gemm")
# "gemm" (default): the in-projection and fc1 bias gradients are column sums of the weight-gradient GEMM's own A operand (dY), taken
#         from its shared-memory operand stages (VbGemmDesc::a_colsum) — no separate pass over dY;
# "id":   q-bias = colsum(dQ) kernel, v-bias through colsum(dV) = colsum(dO) = colsum(d) W_proj, k-bias = 0;
# "1":    column sums inside the attention backward kernel;  "0": one colsum kernel over dqkv.
_WGRAD_COLSUM = _QKV_COLSUM_MODE == "gemm"
LAYER_ROLES = ("ln1_w", "ln1_b", "qkv_w", "qkv_b", "proj_w", "proj_b", "ln2_w", "ln2_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b")
_ALIGN = 64  # elements; keeps every parameter 256-byte (fp32) / 128-byte (bf16) aligned for TMA and float4 access


def _round_up(x, m):
    return (x + m - 1) // m * m


def pick_split_k(tiles, k_blocks, sms):
    """Split-K factor for a wgrad GEMM (work units are tiles x splits, run in waves over `sms` CTA pairs): minimise
    waves x (k-blocks per unit x 0.40 us + 2.1 us for the unit's fp32 reduce-add epilogue) — the two constants are a fit to
    tools/wgrad_split_sweep.py on a B200 — keeping >= 8 k-blocks per unit.  (Filling the last wave to 99 % with 30 splits, which the
    round-1 rule chose for the ViT-B in-projection, costs 11 epilogues per CTA pair instead of 3: 141 us against 124 us.)"""
    best, best_t = 1, float("inf")
    for s in range(1, 33):
        if s > 1 and k_blocks // s < 8:
            break
        t = -(-tiles * s // sms) * (-(-k_blocks // s) * 0.40 + 2.1)
        if t < best_t * 0.995:
            best, best_t = s, t
    return best


def dropout_start_seed():
    """First value of an engine's dropout-mask counter: a function of torch.manual_seed() and the data-parallel rank, so that
    neither every run nor every rank draws the same masks (the counter is bumped once per training forward)."""
    rank = torch.distributed.get_rank() if torch.distributed.is_available() and torch.distributed.is_initialized() else 0
    return ((torch.initial_seed() * 0x9E3779B97F4A7C15 + (rank + 1) * 0x632BE59BD9B4E019) >> 33) & 0x3FFFFFFF


class WorkspaceLease:
    """Marks a workspace as holding the saved activations of a forward whose backward is still pending.  While the lease is alive
    ``workspace()`` hands out another buffer set for the same shape, so a second forward (evaluation inside a training loop, two
    forwards before one backward) cannot overwrite them — the semantics of PyTorch's saved tensors.  Released at the end of the
    backward, or when the autograd node is freed without one."""

    def __init__(self, ws):
        self.ws = ws
        ws["leased"] = True

    def release(self):
        if self.ws is not None:
            self.ws["leased"] = False
            self.ws = None

    __del__ = release


def getstate_without_engine(module):
    """``__getstate__`` of the drop-in modules: the engine (flat buffers' bookkeeping, workspaces, captured CUDA graphs) is a cache that
    is rebuilt lazily, so pickling (torch.save(model)) and copying never see it."""
    d = module.__dict__.copy()
    if "_engine" in d:
        d["_engine"] = None
    d.pop("_solo", None)
    return d


class FlatParams:
    """Flat fp32 parameter / gradient buffers (+ bf16 shadow) that a module's nn.Parameters are views of.

    Subclasses fill ``self._order`` (list of (key, Parameter) in gradient-production order) and call
    ``_layout(segment_sizes)``."""

    def _layout(self, seg_sizes):
        self.offsets = {}
        off, seg_start, count, seg_i = 0, 0, 0, 0
        seg_bounds = []
        self._by_key = {}
        for key, p in self._order:
            self.offsets[key] = off
            self._by_key[key] = p
            off += _round_up(p.numel(), _ALIGN)
            count += 1
            if count == seg_sizes[seg_i]:
                seg_bounds.append((seg_start, off))
                seg_start, count, seg_i = off, 0, seg_i + 1
        self.total = off
        self.segment_bounds = seg_bounds
        self.flat = self.flat_bf16 = self.flat_grad = None
        self._ws = {}
        self._sms = None
        self.grad_segment_hook = None  # callable(segment_index) invoked as soon as a segment's gradients are complete

    def _params_are_views(self):
        if self.flat is None:
            return False
        base = self.flat.data_ptr()
        for key, p in self._order:
            if p.data_ptr() != base + 4 * self.offsets[key] or p.device != self.flat.device:
                return False
        return True

    def bind(self, device):
        """(Re)creates the flat buffers on `device` and re-points every parameter's storage at its slice."""
        flat = torch.zeros(self.total, device=device, dtype=torch.float32)
        for key, p in self._order:
            o = self.offsets[key]
            with torch.no_grad():
                flat[o:o + p.numel()].view(p.shape).copy_(p.data.to(device=device, dtype=torch.float32))
                p.data = flat[o:o + p.numel()].view(p.shape)
        self.flat = flat
        self.flat_bf16 = torch.zeros(self.total, device=device, dtype=torch.bfloat16)
        self.flat_grad = torch.zeros(self.total, device=device, dtype=torch.float32)
        for key, p in self._order:
            p.grad = None
        self._ws.clear()
        self.__dict__.setdefault("_infer_graphs", {}).clear()   # captured graphs hold the old buffers' addresses
        self.__dict__.setdefault("_train_graphs", {}).clear()
        self.bf16_fresh = False
        self._sms = torch.cuda.get_device_properties(device).multi_processor_count

    def _free_workspace(self, key):
        """The first buffer set of this shape that is not leased to a pending backward (None if there is none yet)."""
        for ws in self._ws.setdefault(key, []):
            if not ws.get("leased"):
                return ws
        return None

    def check_input(self, t, name="input"):
        """Inputs must live on the engine's CUDA device: the kernels take raw device pointers (a host tensor would be an illegal access)."""
        if t is not None and (not t.is_cuda or t.device != self.flat.device):
            raise RuntimeError(f"vitb200: {name} is on {t.device} but the model is on {self.flat.device}; move it there first "
                               "(there is no CPU fallback)")
        if torch.cuda.current_device() != self.flat.device.index:
            raise RuntimeError(f"vitb200: the model lives on {self.flat.device} but the current CUDA device is cuda:{torch.cuda.current_device()}; "
                               "call torch.cuda.set_device(...) first (one process per GPU: the kernels launch on the current device)")

    def ensure_bound(self):
        p0 = self._order[0][1]
        if not p0.is_cuda:
            raise RuntimeError("vitb200 runs on a CUDA (sm_100a) device only; move the module to cuda first. "
                               "There is no CPU fallback.")
        if not self._params_are_views():
            self.bind(p0.device)

    def w(self, key):
        """bf16 shadow of a parameter, as a 2-D matrix."""
        p = self._by_key[key]
        o = self.offsets[key]
        v = self.flat_bf16[o:o + p.numel()]
        return v.view(p.shape[0], -1) if p.dim() >= 2 else v

    def f(self, key):
        p = self._by_key[key]
        o = self.offsets[key]
        return self.flat[o:o + p.numel()]

    def gview(self, key):
        p = self._by_key[key]
        o = self.offsets[key]
        v = self.flat_grad[o:o + p.numel()]
        return v.view(p.shape[0], -1) if p.dim() >= 2 else v

    # ------------------------------------------------------------------ small-batch inference ---------------------
    # At batch 1-32 an eval forward is launch-bound (ViT-L/16: ~175 kernels, each behind a ctypes call that encodes TMA descriptors;
    # 3.4 ms at batch 1 where the kernels need < 1 ms), so it is replayed from a CUDA graph: first call per shape eager (lazy allocations,
    # cudaFuncSetAttribute), second call captured, later calls copy the batch into the static input and replay.  The fp32 -> bf16
    # parameter cast stays inside the graph, so parameter edits between calls are always seen.  VITB200_INFER_GRAPH=0 disables it.
    INFER_GRAPH_MAX_BATCH = int(os.environ.get("VITB200_INFER_GRAPH_MAX_BATCH", "32"))

    def forward_inference(self, x, *, want):
        """forward(x, training=False, want=want) for a no-grad caller; CUDA-graph replay for small batches."""
        self.ensure_bound()
        self.check_input(x, "the input batch")
        if os.environ.get("VITB200_INFER_GRAPH", "1") == "0" or x.shape[0] > self.INFER_GRAPH_MAX_BATCH or not x.is_cuda \
                or torch.cuda.is_current_stream_capturing():
            return self.forward(x, training=False, want=want)[0]
        graphs = self.__dict__.setdefault("_infer_graphs", {})
        key = (tuple(x.shape), want)
        ent = graphs.get(key)
        if ent is None:
            graphs[key] = "warm"
            return self.forward(x, training=False, want=want)[0]
        if ent == "warm":
            static_in = torch.empty(x.shape, device=x.device, dtype=torch.float32)
            static_in.copy_(x)
            torch.cuda.synchronize(x.device)
            graph = torch.cuda.CUDAGraph()
            self.bf16_fresh = False      # the captured forward must contain the fp32 -> bf16 parameter cast
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                outs, _ = self.forward(static_in, training=False, want=want)
            ent = graphs[key] = (graph, static_in, outs)
        graph, static_in, outs = ent
        static_in.copy_(x)
        graph.replay()
        return outs

    # ------------------------------------------------------------------ launch-bound training through autograd -------
    # The reference's own loops (base.py:51-57) drive the model through autograd: forward and backward are two calls with the loss in
    # between, so the Trainer's whole-step graph does not apply.  For small problems (the repo's CIFAR configs: 65 tokens x 256 dims)
    # those calls are launch-bound, so each is replayed from its own CUDA graph: first call per shape eager, second captured, later
    # calls copy the batch / the output gradients into static buffers and replay.  Off when a data-parallel reducer is attached (its
    # collectives are issued from Python between kernels).  VITB200_AUTOGRAD_GRAPH=0 / 1 forces it off / on for every size.
    # Captures use capture_error_mode="thread_local": a DataLoader pin-memory thread allocating while the autograd thread captures
    # must not be failed by CUDA's global capture mode.
    AUTOGRAD_GRAPH_MAX_TOKENS = int(os.environ.get("VITB200_AUTOGRAD_GRAPH_MAX_TOKENS", "20000"))

    def _autograd_graph_ok(self, tokens):
        mode = os.environ.get("VITB200_AUTOGRAD_GRAPH", "auto")
        if mode == "0" or self.grad_segment_hook is not None or torch.cuda.is_current_stream_capturing():
            return False
        return mode == "1" or tokens <= self.AUTOGRAD_GRAPH_MAX_TOKENS

    def graphed(self, key, tokens, inputs, fn, before_capture=None, busy=None):
        """``fn(*inputs)`` — eagerly the first time a (key, input shapes) combination is seen, captured into a CUDA graph the second
        time, replayed afterwards with the inputs copied into the graph's static buffers.  ``inputs`` are tensors or None; the result of
        ``fn`` must only reference persistent workspace buffers.  Falls back to the eager call when graphs do not apply, or when
        ``busy(result)`` says the buffers the captured graph writes are leased to a pending backward."""
        if not self._autograd_graph_ok(tokens) or any(t is not None and not t.is_cuda for t in inputs):
            return fn(*inputs)
        graphs = self.__dict__.setdefault("_train_graphs", {})
        full_key = (key, tuple(None if t is None else (tuple(t.shape), t.dtype) for t in inputs))
        ent = graphs.get(full_key)
        if ent is None:
            graphs[full_key] = "warm"
            return fn(*inputs)
        if ent == "warm":
            static = [None if t is None else torch.empty(t.shape, device=t.device, dtype=t.dtype) for t in inputs]
            for st, t in zip(static, inputs):
                if t is not None:
                    st.copy_(t)
            torch.cuda.synchronize(self.flat.device)
            if before_capture is not None:
                before_capture()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                res = fn(*static)
            ent = graphs[full_key] = (graph, static, res)
        graph, static, res = ent
        if busy is not None and busy(res):
            return fn(*inputs)
        for st, t in zip(static, inputs):
            if t is not None:
                st.copy_(t)
        graph.replay()
        return res

    def forward_train(self, x, *, want):
        """forward(x, training=True, want=want) for the autograd node; returns (outputs, ws)."""
        self.ensure_bound()
        self.check_input(x, "the input batch")
        if not (x.is_cuda and self._autograd_graph_ok(x.shape[0] * self.S)):
            return self.forward(x, training=True, want=want)
        graphs = self.__dict__.setdefault("_train_graphs", {})
        key = ("fwd", tuple(x.shape), want, float(self.p_drop), float(self.p_attn))
        ent = graphs.get(key)
        if ent is None:
            graphs[key] = "warm"
            return self.forward(x, training=True, want=want)
        if ent == "warm":
            static_in = torch.empty(x.shape, device=x.device, dtype=torch.float32)
            static_in.copy_(x)
            torch.cuda.synchronize(x.device)
            graph = torch.cuda.CUDAGraph()
            self.bf16_fresh = False      # the captured forward must contain the fp32 -> bf16 parameter cast
            if self._free_workspace((x.shape[0], True)) is None:      # every buffer set is leased: stay eager, capture another time
                return self.forward(x, training=True, want=want)
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                outs, ws = self.forward(static_in, training=True, want=want)
            ent = graphs[key] = (graph, static_in, outs, ws)
        graph, static_in, outs, ws = ent
        if ws.get("leased"):             # its activations still wait for a backward: run eagerly on another buffer set
            return self.forward(x, training=True, want=want)
        static_in.copy_(x)
        graph.replay()
        self.bf16_fresh = False          # the replay cast the parameters itself; a pending "fresh" mark must not outlive it
        return outs, ws

    def backward_train(self, ws, grads, *, want, input_grad=False):
        """backward(ws, grads, want=want) for the autograd node (graph replay under the same conditions as forward_train)."""
        if not self._autograd_graph_ok(ws["M"]) or any(g is not None and not g.is_cuda for g in grads):
            return self.backward(ws, grads, want=want, input_grad=input_grad)
        graphs = self.__dict__.setdefault("_train_graphs", {})
        key = ("bwd", id(ws), want, tuple(None if g is None else tuple(g.shape) for g in grads), ws.get("p_drop", 0.0), ws.get("p_attn", 0.0),
               input_grad)
        ent = graphs.get(key)
        if ent is None:
            graphs[key] = "warm"
            return self.backward(ws, grads, want=want, input_grad=input_grad)
        self.prepare_grads()             # p.grad bookkeeping (and the zeroing it may need) stays outside the graph
        if ent == "warm":
            static_g = [None if g is None else torch.empty(g.shape, device=g.device, dtype=torch.float32) for g in grads]
            for sg, g in zip(static_g, grads):
                if g is not None:
                    sg.copy_(g)
            torch.cuda.synchronize(self.flat.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                res = self.backward(ws, static_g, want=want, input_grad=input_grad)
            ent = graphs[key] = (graph, static_g, res)
        graph, static_g, res = ent
        for sg, g in zip(static_g, grads):
            if g is not None:
                sg.copy_(g)
        graph.replay()
        return res

    def refresh_bf16(self):
        # a fused optimizer step (trainer.FusedAdam) leaves the shadow up to date and sets bf16_fresh
        if getattr(self, "bf16_fresh", False):
            self.bf16_fresh = False
            return
        ops.cast_bf16(self.flat, self.flat_bf16)

    def prepare_grads(self):
        """Makes p.grad of every TRAINABLE parameter a view of the flat gradient buffer (zeroing it if grads were None).  Frozen
        parameters (requires_grad=False: fine-tuning a head on a frozen encoder) keep p.grad = None as under autograd, so an optimizer
        built over model.parameters() leaves them alone; the kernels still write their slices of the flat buffer, which nobody reads."""
        base = self.flat_grad.data_ptr()
        need_zero = False
        foreign = []
        trainable = [(key, p) for key, p in self._order if p.requires_grad]
        for key, p in trainable:
            if p.grad is None:
                need_zero = True
            elif p.grad.data_ptr() != base + 4 * self.offsets[key]:
                foreign.append((key, p, p.grad))
        if need_zero or foreign:
            all_none = all(p.grad is None for _, p in trainable)
            if all_none:
                self.flat_grad.zero_()
            else:
                # keep accumulated values of grads that already are views, zero the slices of the others
                for key, p in trainable:
                    if p.grad is None:
                        o = self.offsets[key]
                        self.flat_grad[o:o + p.numel()].zero_()
            for key, p, gr in foreign:
                o = self.offsets[key]
                self.flat_grad[o:o + p.numel()].view(p.shape).copy_(gr)
            for key, p in trainable:
                o = self.offsets[key]
                p.grad = self.flat_grad[o:o + p.numel()].view(p.shape)

    def _wgrad(self, dy, x, key, rows=None, bias_key=None, bias_grad=None):
        """dW[key] (or its first `rows` rows / a row slice) += dy^T x  with dy [M, N_out], x [M, K_in] (token-major bf16);
        with ``bias_key`` (or ``bias_grad``, an fp32 [N_out] slice of the flat gradient) also d bias += column sums of dy — inside the
        same GEMM when _WGRAD_COLSUM, else one more kernel."""
        dW = self.gview(key)
        if bias_key is not None:
            bias_grad = self.gview(bias_key)
        if rows is not None:
            dW = dW[rows[0]:rows[1]]
        n_out, k_in = dW.shape
        # work units are 256x256 tiles owned by CTA pairs (cta_group::2): sms // 2 pairs run concurrently
        tiles = (-(-(-(-n_out // 128)) // 2)) * (-(-k_in // 256))
        split = pick_split_k(tiles, -(-dy.shape[0] // 64), max(1, self._sms // 2))
        fused = bias_grad is not None and _WGRAD_COLSUM
        ops.gemm(dy, x, dW, a_major=1, b_major=1, epilogue=ops.EPI_ACCUM, split_k=split, a_colsum=bias_grad if fused else None)
        if bias_grad is not None and not fused:
            ops.colsum_bf16(dy, bias_grad)

    def _seg_done(self, idx):
        if self.grad_segment_hook is not None:
            self.grad_segment_hook(idx)


class VitEngine(FlatParams):
    """Sequences the kernels for `n_prefix` token rows (1 = cls, 2 = cls + dist) + patches through L pre-norm blocks."""

    def __init__(self, *, image_size, patch_size, hidden_dim, num_heads, mlp_dim, num_layers, num_classes, n_prefix, eps,
                 globals_, layers, seq_length=None, block_mode=False):
        """globals_: role -> Parameter for cls, [dist], pos, conv_w, conv_b, lnf_w, lnf_b, head_w, head_b, [headd_w, headd_b];
        layers: list of dicts role -> Parameter (LAYER_ROLES).
        block_mode (with seq_length): the bare block stack, tokens in -> tokens out, no position embedding and no final norm — a
        stand-alone EncoderBlock (vanilla_vit.py:73-83); outputs / gradients are requested with want="block"."""
        if hidden_dim % 64 != 0 or hidden_dim // num_heads != 64 or hidden_dim > 1024:
            # every config of the reference has head_dim 64: ViT 256/4, DeiT 192/3, 384/6, 768/12 (utils/args.py:6-15,43-61)
            raise NotImplementedError(f"vitb200 kernels need head_dim == 64 and hidden_dim <= 1024 (got hidden_dim {hidden_dim}, "
                                      f"{num_heads} heads)")
        # tokens mode (seq_length given): the stack is fed [B, S, D] tokens instead of images — a stand-alone Encoder
        # (vanilla_vit.py:88-106): input + pos_embedding, dropout, L blocks, final LayerNorm; no patch embedding, no heads.
        self.tokens_mode = seq_length is not None
        self.block_mode = bool(block_mode)
        assert self.tokens_mode or not self.block_mode
        if self.tokens_mode:
            image_size, patch_size = 4, 4
        assert patch_size % 4 == 0 and image_size % patch_size == 0
        self.image_size, self.p = image_size, patch_size
        self.D, self.H, self.F, self.L, self.C = hidden_dim, num_heads, mlp_dim, num_layers, num_classes
        self.n_prefix, self.eps = n_prefix, eps
        self.P = (image_size // patch_size) ** 2
        self.S = self.P + n_prefix
        if self.tokens_mode:
            self.S, self.P, self.n_prefix = int(seq_length), int(seq_length), 0
        self.Kp = 3 * patch_size * patch_size
        self.Kp_ld = _round_up(self.Kp, 8)
        self.C_ld = _round_up(num_classes, 8)
        self.g = globals_
        self.layers = layers
        self.two_heads = "headd_w" in globals_
        # dropout: rates used by the next training forward; set by the owning module from its constructor arguments
        self.p_drop, self.p_attn = 0.0, 0.0
        self._drop_counter = None
        # gradient-production order: heads + final norm, blocks L-1..0, embedding
        seg0 = ["head_w", "head_b"] + (["headd_w", "headd_b"] if self.two_heads else []) + ["lnf_w", "lnf_b"]
        # Conditional positional encodings: globals "cpe_w"/"cpe_b" = the depthwise conv applied to the token stream
        # right after the class-token cat (cpe_vit.py:143,197; cpvt.py:144,199); per-layer "peg_w"/"peg_b" = CPVT's PEG at the end of
        # every block (cpvt.py:80,93-96); "pos" is optional then (CPVT's Encoder has no learned position embedding: cpvt.py:99-113).
        self.has_cpe = "cpe_w" in globals_
        self.has_peg = bool(layers) and "peg_w" in layers[0]
        self.has_pos = "pos" in globals_
        assert not (self.has_cpe or self.has_peg) or (n_prefix == 1 and not self.tokens_mode), "CPE / PEG: class-token ViT only"
        assert self.has_pos or self.has_cpe or self.block_mode, "a model without a position embedding needs a CPE"
        emb = (["pos"] if self.has_pos else []) + (["cpe_w", "cpe_b"] if self.has_cpe else []) + ["cls"] \
            + (["dist"] if n_prefix == 2 else []) + ["conv_w", "conv_b"]
        if self.tokens_mode:
            seg0, emb = ["lnf_w", "lnf_b"], ["pos"]
        if self.block_mode:
            seg0, emb = [], []
        roles = (("peg_w", "peg_b") if self.has_peg else ()) + LAYER_ROLES
        self._order = [(("g", r), globals_[r]) for r in seg0]
        for li in range(num_layers - 1, -1, -1):
            self._order += [((li, r), layers[li][r]) for r in roles]
        self._order += [(("g", r), globals_[r]) for r in emb]
        self._layout(([len(seg0)] if seg0 else []) + [len(roles)] * num_layers + ([len(emb)] if emb else []))

    # ------------------------------------------------------------------ dropout -----------------------------------
    EMBED_SITE = 4000

    @staticmethod
    def drop_site(layer, site):
        """stream id of a dropout site: site 0 = attention out-proj output (vanilla_vit.py:78), 1 = after GELU (mlp.2, :38),
        2 = MLP output (mlp.4, :42), 3 = attention probabilities (nn.MultiheadAttention(dropout=), :67); EMBED_SITE = Encoder.dropout (:104)."""
        return layer * 8 + site

    def _begin_dropout(self, ws):
        """New masks for this forward: bump the device-side counter and snapshot it into the workspace (graph-capturable)."""
        if self._drop_counter is None or self._drop_counter.device != self.flat.device:
            self._drop_counter = torch.full((1,), dropout_start_seed(), device=self.flat.device, dtype=torch.int32)
        self._drop_counter.add_(1)
        if "drop_seed" not in ws:
            ws["drop_seed"] = torch.zeros(1, device=self.flat.device, dtype=torch.int32)
        ws["drop_seed"].copy_(self._drop_counter)

    # ------------------------------------------------------------------ workspaces --------------------------------
    def workspace(self, B, training):
        key = (B, training)
        ws = self._free_workspace(key)
        if ws is not None:
            return ws
        dev = self.flat.device
        M, D, Fd, S, H = B * self.S, self.D, self.F, self.S, self.H
        bf, f32 = torch.bfloat16, torch.float32
        e = lambda *shape, dtype=bf: torch.empty(*shape, device=dev, dtype=dtype)
        n_sets = self.L if training else 1
        ws = {"B": B, "M": M}
        if not self.tokens_mode:
            ws["patches"] = torch.zeros(B, self.P, self.Kp_ld, device=dev, dtype=bf)
        ws["x"] = [e(B, S, D, dtype=f32) for _ in range(self.L + 1 if training else 2)]
        ws["layer"] = []
        for _ in range(n_sets):
            ws["layer"].append({
                "h1": e(M, D), "qkv": e(M, 3 * D), "o": e(M, D), "lse": e(B, H, S, dtype=f32), "x1": e(B, S, D, dtype=f32),
                "h2": e(M, D), "a": e(M, Fd) if training else None, "g": e(M, Fd),
                "mean1": e(M, dtype=f32), "rstd1": e(M, dtype=f32), "mean2": e(M, dtype=f32), "rstd2": e(M, dtype=f32),
            })
        if self.has_cpe:
            ws["xt"] = e(B, S, D, dtype=f32)       # tokens before the CPE (its wgrad needs them)
        if self.has_peg:
            for lb in ws["layer"]:
                lb["x2"] = e(B, S, D, dtype=f32)   # x1 + mlp(ln_2(x1)): the PEG's input
        if training and (self.has_cpe or self.has_peg):
            ws["d_res"] = e(B, S, D, dtype=f32)    # conv^T of the stream gradient
        ws["y_all"] = None  # allocated on demand (forward_features)
        ws["y_tok_f32"] = e(B * max(self.n_prefix, 1), D, dtype=f32)
        ws["y_tok"] = e(B * max(self.n_prefix, 1), D)
        ws["meanf"] = e(M, dtype=f32)
        ws["rstdf"] = e(M, dtype=f32)
        ws["logits"] = [torch.zeros(B, self.C_ld, device=dev, dtype=f32) for _ in range(2 if self.two_heads else 1)]
        if training:
            ws["d"] = e(B, S, D, dtype=f32)        # fp32 gradient of the residual stream (updated in place)
            ws["d_bf16"] = e(M, D)                 # its bf16 copy (GEMM operand)
            ws["dh"] = e(M, D)                     # gradient w.r.t. a LayerNorm output / attention output
            ws["da"] = e(M, Fd)
            ws["dqkv"] = e(M, 3 * D)
            ws["delta"] = e(B, H, S, dtype=f32)
            ws["dlogits"] = [torch.zeros(B, self.C_ld, device=dev, dtype=bf) for _ in range(2 if self.two_heads else 1)]
            ws["dy_tok"] = e(B * max(self.n_prefix, 1), D)
            ws["dh_cls"] = e(B, D)
            ws["cs_tmp"] = e(D, dtype=f32)         # this backward's out-proj bias gradient (feeds the value-bias gradient)
            ws["stat_cls"] = [e(B, dtype=f32), e(B, dtype=f32)]
            ws["possum"] = e(S, D, dtype=f32)
            ws["dxp"] = e(B * self.P, D) if not self.tokens_mode else None
        self._ws[key].append(ws)
        return ws

    # ------------------------------------------------------------------ forward -----------------------------------
    def _embed(self, ws, images):
        B = ws["B"]
        if self.block_mode:    # a bare block: the caller's tokens are the block input
            ws["x"][0].view(ws["M"], self.D).copy_(images.view(ws["M"], self.D))
            return
        if self.tokens_mode:   # images = caller-supplied tokens [B, S, D]
            x02 = ws["x"][0].view(ws["M"], self.D)
            ops.add_rows_bcast(images.view(ws["M"], self.D), self.f(("g", "pos")).view(self.S, self.D), x02)
            if ws.get("p_drop", 0.0) > 0:
                ops.dropout_f32(x02, ws["p_drop"], ws["drop_seed"], self.EMBED_SITE, dst=x02)
            return
        pat = ws["patches"]
        if self.Kp_ld == self.Kp:
            ops.patchify(images, pat, self.p)
        else:  # never hit by the reference configs (3*p*p is a multiple of 8 for p % 4 == 0)
            raise RuntimeError("patch matrix pitch must be a multiple of 8")
        x0 = ws["x"][0]
        if self.has_cpe:
            # cat(cls, conv_proj(patches)) -> CPE -> (+ pos_embedding, cpe_vit.py:112): the pos add rides on the depthwise-conv kernel
            xt = ws["xt"]
            ops.gemm(pat, self.w(("g", "conv_w")), xt, bias=self.f(("g", "conv_b")), c_row_offset=self.n_prefix)
            ops.token_rows(xt, self.f(("g", "cls")), None, None, self.n_prefix)
            ops.dwconv_fwd(xt, self.f(("g", "cpe_w")), self.f(("g", "cpe_b")), x0, n_prefix=self.n_prefix,
                           pos=self.f(("g", "pos")) if self.has_pos else None)
        else:
            pos = self.f(("g", "pos")).view(1, self.S, self.D)
            ops.gemm(pat, self.w(("g", "conv_w")), x0, epilogue=ops.EPI_RESIDUAL, bias=self.f(("g", "conv_b")), aux=pos,
                     c_row_offset=self.n_prefix, aux_broadcast=True)
            ops.token_rows(x0, self.f(("g", "cls")), self.f(("g", "dist")) if self.n_prefix == 2 else None, pos.view(self.S, self.D),
                           self.n_prefix)
        if ws.get("p_drop", 0.0) > 0:
            x02 = x0.view(ws["M"], self.D)
            ops.dropout_f32(x02, ws["p_drop"], ws["drop_seed"], self.EMBED_SITE, dst=x02)

    def _block_fwd(self, li, x_in, x_out, buf, ws, training):
        B, S, D, H = ws["B"], self.S, self.D, self.H
        M = ws["M"]
        xin2, x12, xout2 = x_in.view(M, D), buf["x1"].view(M, D), x_out.view(M, D)
        ops.layernorm_fwd(xin2, self.f((li, "ln1_w")), self.f((li, "ln1_b")), self.eps, y_bf16=buf["h1"],
                          mean=buf["mean1"] if training else None, rstd=buf["rstd1"] if training else None)
        ops.gemm(buf["h1"], self.w((li, "qkv_w")), buf["qkv"], bias=self.f((li, "qkv_b")))
        qkv = buf["qkv"]
        pd, pa = ws.get("p_drop", 0.0), ws.get("p_attn", 0.0)
        ops.attention_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], buf["o"], buf["lse"] if training else None, B=B, H=H, S=S,
                          tok_stride=1, batch_stride=S, dropout=(pa, ws["drop_seed"], self.drop_site(li, 3)) if pa > 0 else None)
        # Small-batch inference: a residual GEMM over <= 4 row tiles occupies a handful of CTA pairs and walks K serially (fc2 of
        # ViT-L: 4 pairs x 64 k-blocks), so it is split along K: out = residual + bias first, then split-K partial products are
        # reduce-added into it (fp32, TMA reduce).
        # Opt-in (VITB200_INFER_SPLIT_K=1): the reduce order of the partial sums is not fixed, so outputs differ in the last fp32 bit from
        # call to call; the default keeps evaluation bit-reproducible.
        split_small = _INFER_SPLIT_K and (not training) and pd == 0 and M <= 1024
        if split_small:
            self._residual_gemm_split_k(buf["o"], (li, "proj_w"), (li, "proj_b"), xin2, x12)
        elif pd > 0:   # x1 = dropout(out_proj(o)) + x_in   (vanilla_vit.py:77-79)
            ops.gemm(buf["o"], self.w((li, "proj_w")), ws["tmp32"], bias=self.f((li, "proj_b")))
            ops.dropout_f32(ws["tmp32"], pd, ws["drop_seed"], self.drop_site(li, 0), aux=xin2, dst=x12)
        else:
            ops.gemm(buf["o"], self.w((li, "proj_w")), x12, epilogue=ops.EPI_RESIDUAL, bias=self.f((li, "proj_b")), aux=xin2)
        ops.layernorm_fwd(x12, self.f((li, "ln2_w")), self.f((li, "ln2_b")), self.eps, y_bf16=buf["h2"],
                          mean=buf["mean2"] if training else None, rstd=buf["rstd2"] if training else None)
        ops.gemm(buf["h2"], self.w((li, "fc1_w")), buf["a"] if training else None, C2=buf["g"], epilogue=ops.EPI_GELU,
                 bias=self.f((li, "fc1_b")))
        x2 = buf["x2"].view(M, D) if self.has_peg else xout2    # CPVT: x2 = x1 + y feeds the PEG; otherwise it is the block output
        if split_small:
            self._residual_gemm_split_k(buf["g"], (li, "fc2_w"), (li, "fc2_b"), x12, x2)
        elif pd > 0:   # mlp.2: the same mask scales gelu(x) and the saved gelu'(x), so the backward epilogue stays a single multiply
            ops.dropout_bf16_pair(buf["g"], buf["a"], pd, ws["drop_seed"], self.drop_site(li, 1))
            ops.gemm(buf["g"], self.w((li, "fc2_w")), ws["tmp32"], bias=self.f((li, "fc2_b")))
            ops.dropout_f32(ws["tmp32"], pd, ws["drop_seed"], self.drop_site(li, 2), aux=x12, dst=x2)   # mlp.4 + residual (:83)
        else:
            ops.gemm(buf["g"], self.w((li, "fc2_w")), x2, epilogue=ops.EPI_RESIDUAL, bias=self.f((li, "fc2_b")), aux=x12)
        if self.has_peg:   # cpvt.py:93-96: x = x + y; x = peg(x); return x + y   (y = x2 - x1 inside the kernel)
            ops.dwconv_fwd(buf["x2"], self.f((li, "peg_w")), self.f((li, "peg_b")), x_out, n_prefix=self.n_prefix, sub=buf["x1"])

    def _residual_gemm_split_k(self, A, wkey, bkey, residual, out):
        """out = residual + bias + A W^T with the contraction split over the idle CTA pairs (VB_EPI_ACCUM into the pre-filled output)."""
        W = self.w(wkey)
        tiles = -(-A.shape[0] // 256) * -(-W.shape[0] // 256)
        k_blocks = -(-W.shape[1] // 64)
        split = max(1, min(k_blocks // 4, max(1, self._sms // 2) // tiles))
        ops.add_rows_bcast(residual, self.f(bkey).view(1, -1), out)
        ops.gemm(A, W, out, epilogue=ops.EPI_ACCUM, split_k=split)

    def forward(self, images, *, training, want):
        """want: 'logits' (head(s) on the prefix token rows) or 'features' ([B,S,D] fp32 after the final norm).
        Returns (outputs, ws) where outputs is a list of fp32 tensors (views into the workspace)."""
        self.ensure_bound()
        self.check_input(images, "the input batch")
        if images.dtype != torch.float32 or not images.is_contiguous():
            images = images.contiguous().float()
        B = images.shape[0]
        ws = self.workspace(B, training)
        ws["p_drop"] = float(self.p_drop) if training else 0.0
        ws["p_attn"] = float(self.p_attn) if training else 0.0
        if ws["p_drop"] > 0 or ws["p_attn"] > 0:
            self._begin_dropout(ws)
            if ws["p_drop"] > 0 and "tmp32" not in ws:
                ws["tmp32"] = torch.empty(ws["M"], self.D, device=self.flat.device, dtype=torch.float32)
        else:
            ws.setdefault("drop_seed", None)
        self.refresh_bf16()
        self._embed(ws, images)
        xs = ws["x"]
        for li in range(self.L):
            buf = ws["layer"][li if training else 0]
            x_in = xs[li] if training else xs[li & 1]
            x_out = xs[li + 1] if training else xs[(li + 1) & 1]
            self._block_fwd(li, x_in, x_out, buf, ws, training)
        x_last = xs[self.L] if training else xs[self.L & 1]
        ws["x_last"] = x_last
        S, D, M = self.S, self.D, ws["M"]
        if want == "block":
            return [x_last], ws
        if want == "features":
            if ws["y_all"] is None:
                ws["y_all"] = torch.empty(B, S, D, device=x_last.device, dtype=torch.float32)
            ops.layernorm_fwd(x_last.view(M, D), self.f(("g", "lnf_w")), self.f(("g", "lnf_b")), self.eps, y_f32=ws["y_all"].view(M, D),
                              mean=ws["meanf"] if training else None, rstd=ws["rstdf"] if training else None)
            return [ws["y_all"]], ws
        # logits: the final LayerNorm is only needed on the prefix token rows (ViT.forward takes x[:, 0]: vanilla_vit.py:212)
        outs = []
        for t in range(self.n_prefix):
            rows = x_last[:, t, :]                      # [B, D] view with pitch S*D
            sl = slice(t * B, (t + 1) * B)
            ops.layernorm_fwd(rows, self.f(("g", "lnf_w")), self.f(("g", "lnf_b")), self.eps, y_bf16=ws["y_tok"][sl],
                              mean=ws["meanf"][sl] if training else None, rstd=ws["rstdf"][sl] if training else None)
            hw, hb = ("head_w", "head_b") if t == 0 else ("headd_w", "headd_b")
            logits = ws["logits"][t][:, :self.C]
            ops.gemm(ws["y_tok"][sl], self.w(("g", hw)), logits, bias=self.f(("g", hb)))
            outs.append(logits)
        return outs, ws

    # ------------------------------------------------------------------ backward ----------------------------------
    def backward(self, ws, grads, *, want, dlogits_ready=False, input_grad=False):
        """grads: list matching forward outputs (fp32). Accumulates into the flat gradient buffer.
        dlogits_ready: ws["dlogits"][t] already holds the bf16 logits gradients (fused cross-entropy path).
        input_grad: also return d loss / d images ([B,3,H,W] fp32, workspace-owned) — what autograd gives the reference's conv_proj
        when the batch requires grad (saliency maps, adversarial examples)."""
        self.prepare_grads()
        B, S, D, M, L = ws["B"], self.S, self.D, ws["M"], self.L
        d, d_bf, dh = ws["d"], ws["d_bf16"], ws["dh"]
        d2 = d.view(M, D)
        x_last = ws["x"][L]
        last_b2 = (L - 1, "fc2_b")
        pd, pa, seed = ws.get("p_drop", 0.0), ws.get("p_attn", 0.0), ws.get("drop_seed")

        def masked_operand(layer, site, bias_key):
            # hidden dropout in backward: the bf16 GEMM operand of the dropped linear output is keep * d / (1 - p) (mask regenerated
            # from the forward's seed); its column sum is that layer's bias gradient
            ops.dropout_f32(d2, pd, seed, self.drop_site(layer, site), dst_bf16=d_bf)
            ops.colsum_bf16(d_bf, self.gview(bias_key))
        peg = self.has_peg
        fuse_tail = pd == 0 and not peg   # the LayerNorm backward also emits the bf16 operand + fc2 bias gradient of the block behind it
        if want == "block":
            # no final norm behind the last block: its output gradient is the stream gradient; the bf16 operand and the fc2 bias
            # gradient the final LayerNorm backward would have emitted are made here
            gy = grads[0]
            if gy.dtype != torch.float32 or not gy.is_contiguous():
                gy = gy.contiguous().float()
            d2.copy_(gy.view(M, D))
            if fuse_tail:
                ops.cast_bf16(d2.view(-1), d_bf.view(-1))
                ops.colsum_bf16(d_bf, self.gview(last_b2))
        elif want == "features":
            gy = grads[0]
            if gy.dtype != torch.float32 or not gy.is_contiguous():
                gy = gy.contiguous().float()
            ops.layernorm_bwd(gy.view(M, D), x_last.view(M, D), ws["meanf"], ws["rstdf"], self.f(("g", "lnf_w")), dx=d2,
                              dx_bf16=d_bf if fuse_tail else None, dgamma=self.gview(("g", "lnf_w")), dbeta=self.gview(("g", "lnf_b")),
                              dx_colsum=self.gview(last_b2) if fuse_tail else None)
        else:
            d.zero_()
            d_bf.zero_()
            for t in range(self.n_prefix):
                if grads[t] is None and not dlogits_ready:
                    continue
                sl = slice(t * B, (t + 1) * B)
                hw, hb = ("head_w", "head_b") if t == 0 else ("headd_w", "headd_b")
                dl = ws["dlogits"][t]
                if not dlogits_ready:
                    # fp32 -> bf16 cast of the logits gradient into the padded operand buffer (padding stays zero)
                    tmp = ws["logits"][t]
                    tmp[:, :self.C].copy_(grads[t])
                    ops.cast_bf16(tmp.view(-1), dl.view(-1))
                dlv = dl[:, :self.C]
                ops.gemm(dlv, ws["y_tok"][sl], self.gview(("g", hw)), a_major=1, b_major=1, epilogue=ops.EPI_ACCUM)
                ops.colsum_bf16(dl, self._padded_bias_grad(hb))
                ops.gemm(dlv, self.w(("g", hw)), ws["dy_tok"][sl], b_major=1)
                self._fold_bias_grad(hb)
                xr = x_last[:, t, :]
                d_rows = d[:, t, :]
                db_rows = d_bf.view(B, S, D)[:, t, :]
                ops.layernorm_bwd(ws["dy_tok"][sl], xr, ws["meanf"][sl], ws["rstdf"][sl], self.f(("g", "lnf_w")), dx=d_rows,
                                  dx_bf16=db_rows if fuse_tail else None, dgamma=self.gview(("g", "lnf_w")), dbeta=self.gview(("g", "lnf_b")),
                                  dx_colsum=self.gview(last_b2) if fuse_tail else None)
        if pd > 0 and not peg:
            masked_operand(L - 1, 2, last_b2)
        self._seg_done(0)
        for li in range(L - 1, -1, -1):
            buf = ws["layer"][li]
            x_in = ws["x"][li].view(M, D)
            d_res = d2
            if peg:
                # cpvt.py:93-96 backward: out = peg(x2) + y, x2 = x1 + y  =>  d x2 = conv^T(d out), d y = d out + d x2, d x1 (residual) = d x2
                d_res = ws["d_res"].view(M, D)
                ops.dwconv_bwd_weight(d, buf["x2"], self.gview((li, "peg_w")).view(-1), self.gview((li, "peg_b")), n_prefix=self.n_prefix)
                if pd > 0:   # y = dropout(fc2 output): the GEMM operand is keep * (d out + d x2) / (1 - p)
                    ops.dwconv_bwd_data(d, self.f((li, "peg_w")), n_prefix=self.n_prefix, dx=ws["d_res"], sum_f32=ws["tmp32"])
                    ops.dropout_f32(ws["tmp32"], pd, seed, self.drop_site(li, 2), dst_bf16=d_bf)
                else:
                    ops.dwconv_bwd_data(d, self.f((li, "peg_w")), n_prefix=self.n_prefix, dx=ws["d_res"], sum_bf16=d_bf)
                ops.colsum_bf16(d_bf, self.gview((li, "fc2_b")))
            if li == L - 1 and want == "logits" and self.n_prefix == 1 and pd == 0 and not peg:
                # ViT.forward only consumes x[:, 0] (vanilla_vit.py:212): the gradient entering the last block is non-zero
                # on the class-token rows only, so its MLP / out-proj backward runs on B rows instead of B*S.  The rows are
                # addressed in place through strided views (row pitch S * width); nothing is gathered.
                Fd = self.F
                d_c = d_bf.view(B, S, D)[:, 0, :]
                da_c, dh_c = ws["da"][:B], ws["dh_cls"]
                self._wgrad(d_c, buf["g"].view(B, S, Fd)[:, 0, :], (li, "fc2_w"))
                ops.gemm(d_c, self.w((li, "fc2_w")), da_c, b_major=1, epilogue=ops.EPI_DGELU, aux=buf["a"].view(B, S, Fd)[:, 0, :])
                self._wgrad(da_c, buf["h2"].view(B, S, D)[:, 0, :], (li, "fc1_w"), bias_key=(li, "fc1_b"))
                ops.gemm(da_c, self.w((li, "fc1_w")), dh_c, b_major=1)
                ws["stat_cls"][0].copy_(buf["mean2"].view(B, S)[:, 0])
                ws["stat_cls"][1].copy_(buf["rstd2"].view(B, S)[:, 0])
                ops.layernorm_bwd(dh_c, buf["x1"][:, 0, :], ws["stat_cls"][0], ws["stat_cls"][1], self.f((li, "ln2_w")), dres=d[:, 0, :],
                                  dx=d[:, 0, :], dx_bf16=d_c, dgamma=self.gview((li, "ln2_w")), dbeta=self.gview((li, "ln2_b")),
                                  dx_colsum=self.gview((li, "proj_b")))
                self._wgrad(d_c, buf["o"].view(B, S, D)[:, 0, :], (li, "proj_w"))
                dh.zero_()
                ops.gemm(d_c, self.w((li, "proj_w")), dh.view(B, S, D)[:, 0, :], b_major=1)
            else:
                # ---- MLP ----
                self._wgrad(d_bf, buf["g"], (li, "fc2_w"))
                ops.gemm(d_bf, self.w((li, "fc2_w")), ws["da"], b_major=1, epilogue=ops.EPI_DGELU, aux=buf["a"])
                self._wgrad(ws["da"], buf["h2"], (li, "fc1_w"), bias_key=(li, "fc1_b"))
                ops.gemm(ws["da"], self.w((li, "fc1_w")), dh, b_major=1)
                # the out-proj bias gradient of THIS backward goes to a scratch first when it also yields the value-bias gradient
                vb_chain = pd == 0 and pa == 0 and _QKV_COLSUM_MODE == "id"
                if vb_chain:
                    ws["cs_tmp"].zero_()
                ops.layernorm_bwd(dh, buf["x1"].view(M, D), buf["mean2"], buf["rstd2"], self.f((li, "ln2_w")), dres=d_res, dx=d2,
                                  dx_bf16=d_bf if pd == 0 else None, dgamma=self.gview((li, "ln2_w")), dbeta=self.gview((li, "ln2_b")),
                                  dx_colsum=(ws["cs_tmp"] if vb_chain else self.gview((li, "proj_b"))) if pd == 0 else None)
                if vb_chain:   # d proj_b += cs;  d v-bias += cs W_proj  (= colsum(dO) = colsum(dV): DESIGN.md §3.2)
                    ops.vecmat_accum(ws["cs_tmp"], self.f((li, "proj_w")).view(D, D), self.gview((li, "qkv_b"))[2 * D:],
                                     x_accum=self.gview((li, "proj_b")))
                    ws["vb_done"] = True
                if pd > 0:
                    masked_operand(li, 0, (li, "proj_b"))
                # ---- attention ----
                self._wgrad(d_bf, buf["o"], (li, "proj_w"))
                ops.gemm(d_bf, self.w((li, "proj_w")), dh, b_major=1)
            qkv, dqkv = buf["qkv"], ws["dqkv"]
            cs_mode = _QKV_COLSUM_MODE if (pa == 0 or _QKV_COLSUM_MODE != "id") else "1"   # the identity needs un-dropped P rows
            qkv_bg = self.gview((li, "qkv_b"))
            if cs_mode == "id" and not ws.pop("vb_done", False):
                ops.colsum_bf16(dh, qkv_bg[2 * D:])          # d v-bias = colsum(dV) = colsum(dO): dO was just written by the out-proj dgrad
            ops.attention_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], buf["o"], buf["lse"], dh, dqkv[:, :D], dqkv[:, D:2 * D],
                              dqkv[:, 2 * D:], ws["delta"], B=B, H=self.H, S=S, tok_stride=1, batch_stride=S,
                              dropout=(pa, seed, self.drop_site(li, 3)) if pa > 0 else None,
                              dqkv_colsum=qkv_bg if cs_mode == "1" else None)
            if cs_mode == "id":
                ops.colsum_bf16(dqkv[:, :D], qkv_bg[:D])     # d q-bias (the k-bias gradient is exactly zero)
            self._wgrad(dqkv, buf["h1"], (li, "qkv_w"), bias_key=(li, "qkv_b") if cs_mode == "gemm" else None)
            if cs_mode == "0":
                ops.colsum_bf16(dqkv, qkv_bg)
            ops.gemm(dqkv, self.w((li, "qkv_w")), dh, b_major=1)
            prev_b2 = self.gview((li - 1, "fc2_b")) if li > 0 else None
            ops.layernorm_bwd(dh, x_in, buf["mean1"], buf["rstd1"], self.f((li, "ln1_w")), dres=d2, dx=d2,
                              dx_bf16=d_bf if (li > 0 and fuse_tail) else None, dgamma=self.gview((li, "ln1_w")), dbeta=self.gview((li, "ln1_b")),
                              dx_colsum=prev_b2 if fuse_tail else None)
            if pd > 0 and li > 0 and not peg:
                masked_operand(li - 1, 2, (li - 1, "fc2_b"))
            self._seg_done(L - li)
        if self.block_mode:    # d is the gradient w.r.t. the caller's tokens
            return d
        # ---- embedding ----
        if pd > 0:   # Encoder.dropout (vanilla_vit.py:104): gradient of x + pos is keep * d / (1 - p)
            ops.dropout_f32(d2, pd, seed, self.EMBED_SITE, dst=d2)
        if self.tokens_mode:   # d pos = sum_b d; d is the gradient w.r.t. the caller's tokens
            ops.embed_bwd(d, ws["possum"], None, self.gview(("g", "pos")).view(-1), None, None, None, 0)
            self._seg_done(L + 1)
            return d
        if self.has_cpe:
            # x0 = CPE(cat(cls, conv_proj)) (+ pos): d pos = sum_b d; CPE weight gradients from (d, saved tokens); then the class-token /
            # conv-bias / conv-weight gradients from conv^T(d)
            if self.has_pos:
                ops.embed_bwd(d, ws["possum"], None, self.gview(("g", "pos")).view(-1), None, None, None, self.n_prefix)
            ops.dwconv_bwd_weight(d, ws["xt"], self.gview(("g", "cpe_w")).view(-1), self.gview(("g", "cpe_b")), n_prefix=self.n_prefix)
            ops.dwconv_bwd_data(d, self.f(("g", "cpe_w")), n_prefix=self.n_prefix, dx=ws["d_res"])
            ops.embed_bwd(ws["d_res"], ws["possum"], ws["dxp"], None, self.gview(("g", "cls")).view(-1), None, self.gview(("g", "conv_b")),
                          self.n_prefix)
        else:
            ops.embed_bwd(d, ws["possum"], ws["dxp"], self.gview(("g", "pos")).view(-1), self.gview(("g", "cls")).view(-1),
                          self.gview(("g", "dist")).view(-1) if self.n_prefix == 2 else None, self.gview(("g", "conv_b")), self.n_prefix)
        pat = ws["patches"].view(B * self.P, self.Kp_ld)[:, :self.Kp]
        self._wgrad(ws["dxp"], pat, ("g", "conv_w"))
        self._seg_done(L + 1)
        if input_grad:     # d patches = d tokens . conv_w (dgrad GEMM), scattered back to the image layout
            if "dimg" not in ws:
                ws["dpat"] = torch.empty(B * self.P, self.Kp_ld, device=d.device, dtype=torch.bfloat16)
                ws["dimg"] = torch.empty(B, 3, self.image_size, self.image_size, device=d.device, dtype=torch.float32)
            ops.gemm(ws["dxp"], self.w(("g", "conv_w")), ws["dpat"], b_major=1)
            ops.unpatchify(ws["dpat"], ws["dimg"], self.p)
            return ws["dimg"]
        return None

    # The head bias has num_classes entries, but the column-sum kernel works on the 8-padded logits-gradient
    # buffer; sum into a padded scratch and fold the valid part into the gradient.
    def _padded_bias_grad(self, hb):
        if not hasattr(self, "_hb_scratch") or self._hb_scratch.device != self.flat.device:
            self._hb_scratch = torch.zeros(self.C_ld, device=self.flat.device, dtype=torch.float32)
        self._hb_scratch.zero_()
        return self._hb_scratch

    def _fold_bias_grad(self, hb):
        self.gview(("g", hb)).add_(self._hb_scratch[:self.C])
