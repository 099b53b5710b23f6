"""TEST INFRASTRUCTURE ONLY (imported by tests/ and nothing else): fp32 restatement of the reference's
optimizer step for the FastSpeech2 path.

  * global-norm gradient clipping -- pytorch_lightning `gradient_clip_val` (reference main.py:104-110,
    config/train/baseline.yaml `grad_clip_thresh: 1.0`), i.e. torch.nn.utils.clip_grad_norm_(params, max_norm, 2):
    coef = max_norm / (total_norm + 1e-6), applied only when < 1;
  * torch.optim.Adam with the reference's hyper-parameters (lightning/optimizer.py:5-16): L2 weight decay added
    to the gradient, bias-corrected first / second moments, eps added after the sqrt;
  * the LambdaLR factor of lightning/scheduler.py:21-41 (`sqrt_schedule`) and :44-62 (`const_schedule`): the k-th
    optimizer step (k = 0, 1, ...) runs at lr0 * factor(k) with current_step = k + 1.

Pinned by tests/golden/optim.pt, which oracle/make_golden_optim.py produces by running torch.optim.Adam +
clip_grad_norm_ + the reference's own scheduler functions imported from /root/reference.
"""
import math

import torch


def lr_factor(k, sched, warmup, anneal_steps, anneal_rate):
    """lightning/scheduler.py:26-38 (sqrt) and :49-60 (const); k = number of scheduler steps taken."""
    if sched == "none":
        return 1.0
    cur = k + 1
    if warmup > 0:
        if cur <= warmup:
            f = cur / warmup
        else:
            f = math.sqrt(warmup / cur) if sched == "sqrt" else 1.0
    else:
        f = 1.0
    for s in anneal_steps:
        if cur > s:
            f = f * anneal_rate
    return f


def clip_coef(grads, max_norm):
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).float()
    coef = max_norm / (total + 1e-6)
    return torch.clamp(coef, max=1.0), total


def adam_update(p, g, m, v, t, lr, b1, b2, eps, wd):
    """One torch.optim.Adam update (t = 1-based step), in place on p, m, v."""
    if wd != 0:
        g = g + wd * p
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1 = 1 - b1 ** t
    bc2 = 1 - b2 ** t
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))


def run(params, grads_per_step, cfg):
    """params: list of fp32 tensors (modified in place); grads_per_step: list (steps) of lists of tensors.
    cfg: dict(lr, betas, eps, weight_decay, max_norm, sched, warmup, anneal_steps, anneal_rate).
    Returns (lrs, total_norms)."""
    ms = [torch.zeros_like(p) for p in params]
    vs = [torch.zeros_like(p) for p in params]
    lrs, norms = [], []
    for k, grads in enumerate(grads_per_step):
        grads = [g.clone() for g in grads]
        if cfg["max_norm"] > 0:
            coef, total = clip_coef(grads, cfg["max_norm"])
            grads = [g * coef for g in grads]
            norms.append(float(total))
        lr = cfg["lr"] * lr_factor(k, cfg["sched"], cfg["warmup"], cfg["anneal_steps"], cfg["anneal_rate"])
        lrs.append(lr)
        for p, g, m, v in zip(params, grads, ms, vs):
            adam_update(p, g, m, v, k + 1, lr, cfg["betas"][0], cfg["betas"][1], cfg["eps"], cfg["weight_decay"])
    return lrs, norms
