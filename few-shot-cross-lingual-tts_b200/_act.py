"""Activation-dtype and mask plumbing shared by the boundary modules (host logic only)."""
import torch

BF16 = torch.bfloat16


def to_act(x):
    """Reference callers hand fp32 tensors; the kernels run on bf16. Returns (bf16 tensor, caller dtype)."""
    if not x.is_cuda:
        raise RuntimeError(
            "fs2-b200 modules run on sm_100a CUDA kernels only; got a %s tensor (there is no CPU path)"
            % x.device)
    return (x if x.dtype == BF16 else x.to(BF16)), x.dtype


def from_act(y, dtype):
    return y if y.dtype == dtype else y.to(dtype)


def lens_from_mask(mask, T, B, device):
    """Per-row valid length from a bool padding mask (True = pad, prefix masks as produced by
    get_mask_from_lengths, lightning/utils/tool.py:63-74).  mask may be [B,T] or [B,Tq,Tk]."""
    if mask is None:
        return torch.full((B,), T, dtype=torch.int64, device=device)
    if mask.dim() == 3:
        mask = mask[:, 0, :]
    return (mask.shape[1] - mask.sum(dim=1)).to(torch.int64)
