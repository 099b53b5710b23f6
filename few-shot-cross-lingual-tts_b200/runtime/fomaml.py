"""First-order few-shot adaptation step (BASELINE.json configs[2]; SURVEY.md 8d "C3"): for one task,

    theta_0 = theta
    for i in range(k):  theta_{i+1} = theta_i - inner_lr * grad L(support mini-batch i; theta_i)      (plain SGD)
    outer gradient   = grad L(query batch; theta_k)                       (first order: no second derivatives)
    theta is restored; the outer gradient sits in the flat bucket buffer, averaged over ranks (one task per GPU,
    lightning/systems/adaptor.py:21), ready for the fused clip + Adam step.

The live reference never adapts (config/algorithm/language/fscl-orig.yaml:40-43 `adaptation_steps: 0`, the MAML
object of TransEmbOrig.py:279 is constructed but unused), so this loop is the north_star extension that SURVEY.md
specifies: k = 5 steps (dev_maml.yaml:44) at lr 1e-3 (fscl-orig.yaml:38) on support mini-batches of 8 drawn from the
task's shots.  Every forward / backward is the same sm_100a path as the ordinary step; the SGD update is ONE axpy
launch over the flat parameter buffer (fs2_axpy_f32), the restore one flat device copy.  Inner steps never touch the
NCCL buckets (GradBuckets.sync_enabled = False); only the query gradient is all-reduced.
"""
import torch

from .. import _cabi, ops
from .dp import GradBuckets


class FirstOrderTaskStep:
    def __init__(self, model, loss_fn, buckets=None, inner_lr=1e-3, model_kwargs=None, embedding_model=None,
                 device=None):
        self.model, self.loss_fn = model, loss_fn
        self.embedding_model = embedding_model
        self.model_kwargs = dict(model_kwargs or {})
        self.inner_lr = float(inner_lr)
        self.device = device or next(model.parameters()).device
        params = list(model.parameters())
        if embedding_model is not None:
            params += list(embedding_model.parameters())
        self.buckets = buckets or GradBuckets(params, device=self.device)
        self.flat_param = self.buckets.rehome_parameters()
        self.theta0 = torch.empty_like(self.flat_param)
        self.losses = torch.zeros(6, dtype=torch.float32, device=self.device)

    def _dev(self, batch):
        return tuple(x.to(self.device, non_blocking=True) if torch.is_tensor(x) else x for x in batch)

    def _fwd_bwd(self, b):
        self.buckets.zero()
        emb = b[3] if self.embedding_model is None else self.embedding_model(b[3])
        out = self.model(b[2], emb, *b[4:12], lang_args=b[12], **self.model_kwargs)
        losses = self.loss_fn(tuple(b[:12]), out)
        losses[0].backward()
        return losses

    def run(self, support_batches, query_batch):
        """support_batches: k batch tuples (one inner SGD step each); query_batch: the task's query tuple.
        Returns the six query losses (device tensor, rank-averaged); the outer gradient is in buckets.flat."""
        L = _cabi.lib()
        bk = self.buckets
        ops.set_grad_listener(bk.notify)
        n = self.flat_param.numel()
        self.theta0.copy_(self.flat_param)
        bk.sync_enabled = False
        try:
            for sb in support_batches:
                self._fwd_bwd(self._dev(sb))
                st = torch.cuda.current_stream().cuda_stream
                _cabi.check(L.fs2_axpy_f32(self.flat_param.data_ptr(), bk.flat.data_ptr(), -self.inner_lr, n, st),
                            "axpy(inner SGD)")
        finally:
            bk.sync_enabled = True
        losses = self._fwd_bwd(self._dev(query_batch))  # at the adapted weights; buckets all-reduce as usual
        bk.finish()
        self.flat_param.copy_(self.theta0)  # first-order: the outer update applies the query gradient to theta
        self.losses.copy_(torch.stack([l.detach() for l in losses]))
        bk.reduce_scalars(self.losses)
        ops.advance_rng()
        return self.losses
