"""Shared helpers of the parity tests."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=False)


def disable_dropout(model):
    """Parity mode: dropout = identity, BatchNorm stays in batch-stat mode."""
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if hasattr(m, "p_dropout"):
            m.p_dropout = 0.0
        if m.__class__.__name__ == "VariancePredictor":
            m.dropout = 0.0
    return model


def rel_err(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def cosine(a, b):
    a, b = a.detach().float().cpu().flatten(), b.detach().float().cpu().flatten()
    return (torch.dot(a, b) / (a.norm() * b.norm() + 1e-20)).item()


def cuda_batch(batch):
    return tuple(x.cuda() if torch.is_tensor(x) else x for x in batch)
