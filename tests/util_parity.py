"""Shared helpers of the parity tests."""
import os
import re
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


COS_MIN = 0.999
# Tensors allowed below COS_MIN (measured r2: 0.9940-0.9990 there, every other tensor >= 0.9990).  Reason: the variance
# predictors are [Conv1d -> ReLU -> LayerNorm] x2 regressing targets that are uncorrelated with the input at random
# init, so their gradient is an incoherent sum over phonemes; a pre-activation that sits within bf16 rounding of the
# ReLU kink flips its 0/1 derivative, which changes that element's contribution by 100 %, and an incoherent sum does
# not average this away.  The fp32 reference itself shows it: rounding only the predictor INPUT to bf16 drops these
# cosines to 0.9975-0.9988 (tests/test_oracle.py::test_predictor_gradients_are_intrinsically_sensitive_to_bf16_input).
COS_EXCEPTIONS = [
    (re.compile(r"variance_adaptor\.(pitch|energy|duration)_predictor\.conv_layer\.(conv1d_[12]\.conv|layer_norm_1)\."),
     0.99),
    (re.compile(r"variance_adaptor\.(pitch|energy)_embedding\.weight"), 0.99),
]


def cos_floor(name):
    for pat, floor in COS_EXCEPTIONS:
        if pat.match(name):
            return floor
    return COS_MIN


# Norm-ratio tolerance: 5 %, except the variance predictors' output-side biases (10 %): their gradient is
# sum_i 2 (pred_i - target_i) / N over the phonemes -- with zero-mean normalised targets a cancelling sum, so a 1 %
# error of the predictions is a several-% error of the sum while the direction (cosine 1.0000) is untouched.
RATIO_EXCEPTIONS = [
    (re.compile(r"variance_adaptor\.(pitch|energy|duration)_predictor\.(conv_layer\.layer_norm_2\.bias|linear_layer\.bias)"),
     0.10),
]


def ratio_tol(name):
    for pat, tol in RATIO_EXCEPTIONS:
        if pat.match(name):
            return tol
    return 0.05
