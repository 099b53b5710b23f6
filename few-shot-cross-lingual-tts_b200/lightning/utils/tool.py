"""Host helpers the hot path needs from the reference's lightning/utils/tool.py."""
import torch
import torch.nn.functional as F


def get_mask_from_lengths(lengths, max_len=None):
    """True = padding.  Definition: lightning/utils/tool.py:63-74 (the in-repo copy of the dlhlp_lib
    function the reference imports, fastspeech2m.py:11).  max_len=None costs one host sync."""
    if max_len is None:
        max_len = int(torch.max(lengths).item())
    ids = torch.arange(0, int(max_len), device=lengths.device).unsqueeze(0)
    return ids >= lengths.unsqueeze(1)


def pad(input_ele, mel_max_length=None):
    """Zero-pad (or crop) a list of [T_i, ...] tensors to a common length and stack them
    (lightning/utils/tool.py:168-186).  Host-side utility; the CUDA LengthRegulator does not use it."""
    max_len = mel_max_length if mel_max_length else max(x.size(0) for x in input_ele)
    out = []
    for x in input_ele:
        if x.dim() == 1:
            out.append(F.pad(x, (0, max_len - x.size(0)), "constant", 0.0))
        else:
            out.append(F.pad(x, (0, 0, 0, max_len - x.size(0)), "constant", 0.0))
    return torch.stack(out)
