"""One FastSpeech2 training step (forward + loss + backward [+ gradient all-reduce]) as a replayable
CUDA graph with static device buffers -- the B200-native stand-in for what pytorch_lightning does
around `BaselineSystem.training_step` (lightning/systems/language/FastSpeech2.py:53-90, main.py:34-40).

The graph is keyed by the padded batch shape (B, Ts, Tm): lengths, durations and all tensor contents
are device data, so one capture serves every batch of that shape; dropout masks change per replay
through a device-side step counter (ops.advance_rng).
"""
import torch

from .. import ops
from .dp import GradBuckets

_DEFER = __import__("os").environ.get("FS2_NO_DEFER") is None  # A/B switch for tools

# batch tuple slots that are device tensors (lightning/collates/utils.py:70-85)
_TENSOR_SLOTS = (2, 3, 4, 6, 7, 9, 10, 11, 12)


class TrainStep:
    def __init__(self, model, loss_fn, example_batch, use_graph=True, buckets=None, device=None, optimizer=None,
                 embedding_model=None, model_kwargs=None, wcache=None):
        """optimizer: optional runtime.FusedAdam built on `buckets`; its clip + Adam + LR-schedule launches then
        become part of the captured step.

        embedding_model: optional phoneme-embedding module (MultilingualEmbedding, ...).  Batch slot 3 then holds
        the int64 phoneme ids `[B, Ts]` and `emb_texts = embedding_model(batch[3])` is computed INSIDE the step, as
        in the reference (lightning/systems/language/FastSpeech2.py:53-55); its parameters are part of the gradient
        buckets, so they are all-reduced, clipped and updated together with the model's
        (`nn.ModuleList([model, embedding_model])`, FastSpeech2.py:47-48).  Without it slot 3 holds already
        embedded phonemes `[B, Ts, d]` and no gradient flows past them.  `buckets`, when passed, must have been
        built over the parameters of both modules.

        model_kwargs: extra keyword arguments of the model call (e.g. average_spk_emb=True for FSCL task steps)."""
        self.model, self.loss_fn = model, loss_fn
        self.embedding_model = embedding_model
        self.model_kwargs = dict(model_kwargs or {})
        self.optimizer = optimizer
        self.device = device or next(model.parameters()).device
        params = list(model.parameters())
        if embedding_model is not None:
            params += [p for p in embedding_model.parameters()]
        self.buckets = buckets or GradBuckets(params, device=self.device)
        if embedding_model is not None:
            missing = [p for p in embedding_model.parameters() if p.requires_grad and
                       getattr(p, "main_grad", None) is None]
            if missing:
                raise ValueError("TrainStep: `buckets` does not cover the embedding model's parameters")
        ops.set_grad_listener(self.buckets.notify)
        # all bf16 operand copies of the weights are refreshed by one launch at the start of every step
        # (`wcache`: share one set of copies between several captured steps, runtime/cache.py)
        self.wcache = wcache if wcache is not None else ops.WeightCache(model)
        # The tensors of a batch live in ONE flat buffer per role (256-byte aligned slots; `static[i]` / `stage[i]` /
        # `host[i]` are views): the graph's static inputs, the device staging copy of the next batch, the pinned host
        # copy -- so a batch moves with one copy per hop instead of one per tensor.
        self.static = list(example_batch)
        offs, total = {}, 0
        for i in _TENSOR_SLOTS:
            t = example_batch[i]
            offs[i] = total
            total += (t.numel() * t.element_size() + 255) // 256 * 256

        def views(flat):
            out = {}
            for i in _TENSOR_SLOTS:
                t = example_batch[i]
                n = t.numel() * t.element_size()
                out[i] = flat[offs[i]:offs[i] + n].view(t.dtype).view(t.shape)
            return out

        self._static_flat = torch.zeros(total, dtype=torch.uint8, device=self.device)
        self._stage_flat = torch.zeros(total, dtype=torch.uint8, device=self.device)
        self._host_flat = torch.zeros(total, dtype=torch.uint8).pin_memory()
        sv, self.stage, self.host = views(self._static_flat), views(self._stage_flat), views(self._host_flat)
        for i in _TENSOR_SLOTS:
            self.host[i].copy_(example_batch[i])
            sv[i].copy_(example_batch[i])
            self.static[i] = sv[i]
        self.losses = torch.zeros(6, dtype=torch.float32, device=self.device)
        self.losses_host = torch.empty(6, dtype=torch.float32, pin_memory=True)
        self._lag_host = [torch.zeros(6, dtype=torch.float32).pin_memory() for _ in range(2)]
        self._lag_event = [None, None]
        self._lag_n = 0
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in self.host.values())
        self.d2h_bytes = self.losses_host.numel() * 4
        # double-buffered input path: the next batch is copied host -> device on a copy stream into the staging
        # buffer while the current step runs; run() then moves it into the graph's static buffer (one device to
        # device copy, ~10 us for 34 MB)
        self.copy_stream = torch.cuda.Stream(self.device)
        self._staged = False
        self._stage_free = None
        self.graph = None
        self.use_graph = use_graph
        if use_graph:
            self._capture()

    # ---------------------------------------------------------------------------------------------
    def _body(self):
        b = self.static
        self.wcache.refresh()
        ops.set_weight_cache(self.wcache)  # valid for this step only: the copies were just refreshed
        try:
            self.buckets.zero()
            emb = b[3] if self.embedding_model is None else self.embedding_model(b[3])
            out = self.model(b[2], emb, *b[4:12], lang_args=b[12], **self.model_kwargs)
            losses = self.loss_fn(tuple(b[:12]), out)
            ops.DEFER_JOIN = _DEFER  # weight gradients run free on the side stream until the end of the backward
            try:
                losses[0].backward()
            finally:
                ops.DEFER_JOIN = False
                ops.final_join()
        finally:
            ops.set_weight_cache(None)
        self.buckets.finish()
        if self.optimizer is not None:
            self.optimizer.step()
        self.losses.copy_(torch.stack([l.detach() for l in losses]))
        self.buckets.reduce_scalars(self.losses)
        ops.advance_rng()

    def _capture(self):
        s = torch.cuda.Stream(self.device)
        s.wait_stream(torch.cuda.current_stream())
        import os
        import sys
        dbg = (lambda m: print("[step] " + m, file=sys.stderr, flush=True)) if os.environ.get("FS2_DEBUG") \
            else (lambda m: None)
        with torch.cuda.stream(s):
            # warm-up outside capture (allocator, lazy inits, NCCL communicators).  The warm-up bodies are REAL steps
            # on the example batch (Adam updates, BatchNorm running statistics, dropout counter): everything they
            # mutate is snapshotted and put back, so constructing a TrainStep leaves the training state untouched.
            snap = self._snapshot()
            for i in range(2):
                self._body()
                torch.cuda.synchronize()
                dbg("warm-up body %d done" % i)
            self._restore(snap)
            del snap
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._body()
        dbg("captured")

    def _mutable_state(self):
        """Every tensor a step body writes besides the gradient buffers and the loss read-back."""
        ts = []
        mods = [self.model] + ([self.embedding_model] if self.embedding_model is not None else [])
        for m in mods:
            ts += [b for b in m.buffers()]
            if self.optimizer is not None:
                ts += [p.data for p in m.parameters()]
        if self.optimizer is not None:
            o = self.optimizer
            ts += [o.flat_param, o.exp_avg, o.exp_avg_sq, o.step_dev, o.gnorm_sq, o.grad_norm]
        ts.append(ops._Rng.tensor(self.device))  # dropout step counter (created on first use)
        return ts

    def _snapshot(self):
        return [(t, t.clone()) for t in self._mutable_state()]

    def _restore(self, snap):
        for t, c in snap:
            t.copy_(c)

    def close(self):
        """Drop the captured graph (and with it the captured NCCL kernels) -- call before
        torch.distributed.destroy_process_group()."""
        torch.cuda.synchronize(self.device)
        self.graph = None
        ops.set_grad_listener(None)

    # ---------------------------------------------------------------------------------------------
    def load_batch(self, batch=None):
        """Host -> device copy of one batch (pinned staging buffers, same padded shape)."""
        if batch is not None:
            for i in _TENSOR_SLOTS:
                self.host[i].copy_(batch[i])
        self._static_flat.copy_(self._host_flat, non_blocking=True)

    def prefetch_batch(self, batch=None, pinned=None):
        """Asynchronous host -> device copy of the NEXT batch on a copy stream into device staging buffers: it
        overlaps the step that is currently running; the following run() consumes it.  `batch`: a 13-tuple of
        (pageable) host tensors, first copied into this object's pinned buffers; `pinned`: {slot: pinned host
        tensor} copied from directly (a data loader with pin_memory=True)."""
        src = self.host
        if pinned is not None:
            src = pinned
        elif batch is not None:
            self.copy_stream.synchronize()  # the DMA of the previous prefetch no longer reads self.host
            for i in _TENSOR_SLOTS:
                self.host[i].copy_(batch[i])
        # only the device-to-device moves at the head of the previous run() read the staging buffers: wait for
        # those (not for the whole step), so that this copy overlaps the step that is running now
        if self._stage_free is not None:
            self.copy_stream.wait_event(self._stage_free)
        with torch.cuda.stream(self.copy_stream):
            if src is self.host:
                self._stage_flat.copy_(self._host_flat, non_blocking=True)
            else:
                for i in _TENSOR_SLOTS:
                    self.stage[i].copy_(src[i], non_blocking=True)
        self._staged = True

    def run(self):
        """Device-resident step: inputs are whatever load_batch() / prefetch_batch() last provided."""
        if self._staged:
            cur = torch.cuda.current_stream()
            cur.wait_stream(self.copy_stream)
            self._static_flat.copy_(self._stage_flat, non_blocking=True)
            self._stage_free = torch.cuda.Event()
            self._stage_free.record(cur)
            self._staged = False
        if self.graph is not None:
            self.graph.replay()
        else:
            self._body()

    def read_losses(self):
        self.losses_host.copy_(self.losses, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.losses_host

    def read_losses_lagged(self):
        """Asynchronous logging: enqueue the 24-byte device -> host read of THIS step's losses and return the host
        copy of the PREVIOUS step's (None after the first call).  The host never waits for the step it has just
        launched, so the next step is enqueued while this one runs and the GPU does not idle between steps
        (pytorch_lightning logs the same way: `self.log` keeps the tensor and syncs at the logging interval)."""
        k = self._lag_n & 1
        self._lag_host[k].copy_(self.losses, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self._lag_event[k] = ev
        self._lag_n += 1
        prev = self._lag_event[k ^ 1]
        if prev is None:
            return None
        prev.synchronize()
        return self._lag_host[k ^ 1]

    def step_e2e(self, batch=None):
        """What a user calls: H2D of the batch, the step, D2H of the six losses."""
        self.load_batch(batch)
        self.run()
        return self.read_losses()
