"""Data-parallel gradient plumbing: one flat fp32 gradient buffer, bucketed in reverse execution
order, each bucket all-reduced (NCCL over NVLink / NVSwitch on the box, gloo in the CPU tests) on a
side stream as soon as the backward pass has produced its last gradient.

The reference gets this from pytorch_lightning's `strategy="ddp"` (main.py:34-40): gradients are
averaged over ranks, losses are normalised per rank (mean of means), BatchNorm statistics and dropout
streams stay per rank.  Here the weight-gradient kernels write straight into the flat buffer
(`param.main_grad` views, see ops.grad_target), so a bucket is ready for the wire without a copy.
"""
import torch
import torch.distributed as dist


class GradBuckets:
    def __init__(self, params, bucket_bytes=32 << 20, process_group=None, device=None):
        self.all_params = list(params)  # given order, frozen parameters included (torch.optim state indices)
        self.params = [p for p in self.all_params if p.requires_grad]
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        device = device or self.params[0].device
        # reverse registration order ~ the order in which backward produces gradients
        order = list(reversed(self.params))
        # every gradient starts on a 16-byte boundary (vector reductions / TMA-friendly), so sizes are
        # rounded up to 4 floats; the few padding elements stay zero and ride along in the all-reduce
        total = sum((p.numel() + 3) // 4 * 4 for p in order)
        self.flat = torch.zeros(total, dtype=torch.float32, device=device)
        self.buckets = []  # (start, end) element ranges of self.flat
        self._bucket_of = {}
        self._pending0 = []  # parameters per bucket (sizes only; membership is in self._members)
        off, bstart, bcount = 0, 0, 0
        for p in order:
            n = p.numel()
            if p.dim() == 3 and p.shape[2] > 1 and p.shape[1] % 4 == 0 and device.type == "cuda":
                # Conv1d weight [Co, Ci, k]: the gradient lives in the weight-gradient GEMM's natural order
                # [Co][k][Ci] (unit-stride vector atomics, no scratch + re-layout pass); `main_grad` / `.grad`
                # are the permuted view with the parameter's logical shape, so optimizers see matching elements
                Co, Ci, k = p.shape
                view = self.flat[off:off + n].view(Co, k, Ci).permute(0, 2, 1)
            else:
                view = self.flat[off:off + n].view(p.shape)
            p.main_grad = view
            p.grad = view  # autograd-produced gradients (embeddings) accumulate in place as well
            self._bucket_of[id(p)] = len(self.buckets)
            bcount += 1
            off += (n + 3) // 4 * 4
            if (off - bstart) * 4 >= bucket_bytes:
                self.buckets.append((bstart, off))
                self._pending0.append(bcount)
                bstart, bcount = off, 0
        if off > bstart:
            self.buckets.append((bstart, off))
            self._pending0.append(bcount)
        self._members = [set() for _ in self.buckets]  # parameter ids per bucket
        for pid, b in self._bucket_of.items():
            self._members[b].add(pid)
        self._pending = [set(m) for m in self._members]
        self._launched = [False] * len(self.buckets)
        self._late = set()  # buckets that received a gradient announcement after they went on the wire
        self.comm_stream = torch.cuda.Stream(device) if device.type == "cuda" else None
        self.overlap = True
        self.sync_enabled = True  # False: gradients stay local (inner-loop steps of the few-shot adaptation)
        # NCCL averages inside the collective (no separate divide pass over the 138 MB); gloo has no AVG
        self._avg = False
        if self.world > 1:
            try:
                self._avg = dist.get_backend(process_group) == "nccl"
            except Exception:
                self._avg = False

    def rehome_parameters(self):
        """Move every parameter into ONE flat fp32 buffer with the same element offsets (and the same [Co][k][Ci]
        order for Conv1d weights) as its gradient in `self.flat`, so that optimizers and the few-shot inner loop
        update all of them with one launch.  Idempotent; returns the flat parameter buffer."""
        fp = getattr(self, "flat_param", None)
        if fp is not None:
            return fp
        g = self.flat
        fp = torch.zeros_like(g)
        for p in self.params:
            off = p.main_grad.data_ptr() - g.data_ptr()
            assert off % 4 == 0
            flat = fp[off // 4: off // 4 + p.numel()]
            if p.main_grad.is_contiguous():
                view = flat.view(p.shape)
            else:  # Conv1d weight whose gradient is kept in [Co][k][Ci] order: same element order for the value
                Co, Ci, k = p.shape
                assert p.main_grad.stride() == (k * Ci, 1, Ci)
                view = flat.view(Co, k, Ci).permute(0, 2, 1)
            view.copy_(p.data)
            p.data = view
        self.flat_param = fp
        return fp

    # -- per-step protocol ---------------------------------------------------------------------------
    def zero(self):
        self._main_stream = torch.cuda.current_stream() if self.flat.is_cuda else None
        self.flat.zero_()
        self._pending = [set(m) for m in self._members]
        self._launched = [False] * len(self.buckets)
        self._late.clear()

    def _all_reduce(self, b):
        s, e = self.buckets[b]
        chunk = self.flat[s:e]
        if self.world > 1:
            if self._avg:
                dist.all_reduce(chunk, op=dist.ReduceOp.AVG, group=self.group)
            else:
                dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group)
                chunk.div_(self.world)
        self._launched[b] = True

    def _launch(self, b):
        if self.comm_stream is not None:
            cur = torch.cuda.current_stream()
            self.comm_stream.wait_stream(cur)
            main = getattr(self, "_main_stream", None)
            if main is not None and main != cur:  # announced from a branch stream: the bucket also holds gradients
                self.comm_stream.wait_stream(main)  # produced on the stream that runs the step
            from .. import ops  # weight gradients are produced on the side stream (ops.fork_side)
            for bs in ops.branch_streams(cur.device):  # backward of the variance predictors runs on branch streams
                self.comm_stream.wait_stream(bs)
            side = ops.side_stream_of(cur.device)
            if side is not None:
                self.comm_stream.wait_stream(side)
            with torch.cuda.stream(self.comm_stream):
                self._all_reduce(b)
        else:
            self._all_reduce(b)

    def notify(self, params):
        """Called by the backward functions when the gradients of `params` are final."""
        if self.world == 1 or not self.overlap or not self.sync_enabled:
            return
        for p in params:
            b = self._bucket_of.get(id(p))
            if b is None:
                continue
            if self._launched[b]:
                # a parameter whose backward Function runs more than once per step (module called twice, shared
                # weights, few-shot inner loops) was announced again AFTER its bucket was reduced: the late
                # contribution would be added on top of an already averaged bucket.  Refuse loudly.
                self._late.add(b)
                continue
            self._pending[b].discard(id(p))  # once per distinct parameter, however often it is announced
            if not self._pending[b]:
                self._launch(b)

    def finish(self):
        """Reduce whatever has not been sent yet and join the side stream."""
        if self._late:
            raise RuntimeError("GradBuckets: gradients of bucket(s) %s were announced again after the bucket had been "
                               "all-reduced (a module ran twice in one backward pass); set buckets.overlap = False "
                               "for such steps so that every bucket is reduced in finish()" % sorted(self._late))
        if self.world > 1 and self.sync_enabled:
            for b in range(len(self.buckets)):
                if not self._launched[b]:
                    self._launch(b)
            if self.comm_stream is not None:
                torch.cuda.current_stream().wait_stream(self.comm_stream)

    def reduce_scalars(self, t):
        """Mean over ranks of a small tensor (the six logged losses: FastSpeech2.py:89 sync_dist=True)."""
        if self.world > 1:
            if self._avg:
                dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group)
            else:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
                t.div_(self.world)
        return t
