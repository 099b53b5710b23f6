"""Shape-bucketed cache of captured training steps.

`TrainStep` holds ONE CUDA graph for ONE padded batch shape (B, Ts, Tm); a real data loader produces a new
(max_src_len, max_mel_len) almost every batch, and every capture costs seconds.  `TrainStepCache` pads each incoming
batch up to the next bucket boundary -- by default Ts to a multiple of 32 phonemes and Tm to a multiple of 128 frames,
the row-tile size of the kernels, so the ragged tile schedule skips the extra rows at no cost -- and keeps one captured
step per bucket (LRU, `max_graphs`).  All steps share the gradient buckets, the optimizer and one WeightCache.

What bucketing changes: exactly what the reference itself would compute if its collate had padded the same utterances
to (Ts_bucket, Tm_bucket) and passed those as max_src_len / max_mel_len (lightning/collates/utils.py:70-85) -- the
FFT blocks and the losses only see valid rows, while PostNet BatchNorm statistics and the variance predictors' conv
halo at the batch's longest utterance see padding rows instead of the sequence end, as they do in the reference for
every utterance shorter than the batch maximum.  `bucket=(1, 1)` keeps the exact shapes (one graph per distinct shape).
"""
from collections import OrderedDict

import torch

from .. import ops
from .dp import GradBuckets
from .step import _TENSOR_SLOTS, TrainStep


def _roundup(x, m):
    return (int(x) + m - 1) // m * m


def pad_batch(batch, Ts, Tm):
    """Zero-pad a 13-tuple batch (host or device tensors) to (Ts, Tm); max_src_len / max_mel_len become Ts / Tm."""
    b = list(batch)

    def pad(t, dim, size):
        if t is None or t.shape[dim] == size:
            return t
        shape = list(t.shape)
        shape[dim] = size - t.shape[dim]
        return torch.cat([t, t.new_zeros(shape)], dim=dim)

    b[3] = pad(b[3], 1, Ts)
    for i in (9, 10, 11):  # pitch / energy are per phoneme (or per frame for frame-level features), durations per phoneme
        t = b[i]
        b[i] = pad(t, 1, Ts if t.shape[1] == batch[3].shape[1] else Tm)
    b[6] = pad(b[6], 1, Tm)
    b[5], b[8] = Ts, Tm
    return tuple(b)


class TrainStepCache:
    def __init__(self, model, loss_fn, bucket=(32, 128), max_graphs=8, buckets=None, optimizer=None,
                 embedding_model=None, model_kwargs=None, device=None):
        self.model, self.loss_fn = model, loss_fn
        self.bucket = (max(int(bucket[0]), 1), max(int(bucket[1]), 1))
        self.max_graphs = int(max_graphs)
        self.device = device or next(model.parameters()).device
        params = list(model.parameters()) + (list(embedding_model.parameters()) if embedding_model is not None else [])
        self.buckets = buckets or GradBuckets(params, device=self.device)
        self.optimizer, self.embedding_model, self.model_kwargs = optimizer, embedding_model, model_kwargs
        self.wcache = ops.WeightCache(model)
        self.steps = OrderedDict()  # (B, Ts_bucket, Tm_bucket) -> TrainStep
        self.captures = 0

    def key_for(self, batch):
        return (batch[3].shape[0], _roundup(batch[5], self.bucket[0]), _roundup(batch[8], self.bucket[1]))

    def step_for(self, batch):
        """(captured step for this batch's bucket, the batch padded to it)"""
        key = self.key_for(batch)
        padded = pad_batch(batch, key[1], key[2])
        st = self.steps.get(key)
        if st is None:
            if len(self.steps) >= self.max_graphs:
                _, old = self.steps.popitem(last=False)  # least recently used
                old.close()
            st = TrainStep(self.model, self.loss_fn, padded, use_graph=True, buckets=self.buckets, device=self.device,
                           optimizer=self.optimizer, embedding_model=self.embedding_model,
                           model_kwargs=self.model_kwargs, wcache=self.wcache)
            self.steps[key] = st
            self.captures += 1
        else:
            self.steps.move_to_end(key)
        ops.set_grad_listener(self.buckets.notify)
        return st, padded

    def run(self, batch):
        """One training step on `batch` (any shape): returns the six losses (pinned host tensor)."""
        st, padded = self.step_for(batch)
        return st.step_e2e(padded)
