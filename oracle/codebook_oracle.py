"""TEST INFRASTRUCTURE ONLY: fp32 restatement of the few-shot phoneme-embedding front-end.

  * phoneme_query      -- PhonemeQueryExtractor(mode="average") (reference lightning/model/reduction.py:42-110):
                          per segment mean of the frames (two_stage) then class-wise mean; empty classes -> zeros;
  * codebook2_forward  -- SoftMultiAttCodebook2.forward (lightning/systems/language/embeddings.py:109-142) with
                          MultiheadAttention (transformer/Modules.py:28-47): NaN -> 0, softmax(weight_raw) layer sum,
                          q_linear, H-head softmax attention over att_banks / emb_banks at temperature sqrt(E / H).

Pinned by tests/golden/codebook.pt (oracle/make_golden_codebook.py runs the reference classes themselves).
"""
import torch


def phoneme_query(representations, avg_frames, n_symbols, phonemes, two_stage=True):
    dims = representations[0].shape[1:]
    table = {i: [] for i in range(n_symbols)}
    for ph, d_list, rep in zip(phonemes, avg_frames, representations):
        pos = 0
        for p, d in zip(ph, d_list):
            if d > 0:
                if two_stage:
                    table[int(p)].append(rep[pos:pos + d].mean(dim=0))
                else:
                    table[int(p)].extend(list(rep[pos:pos + d]))
            pos += d
    rows = [torch.stack(table[c]).mean(0) if table[c] else torch.zeros(dims) for c in range(n_symbols)]
    return torch.stack(rows).float().unsqueeze(0)


def codebook2_forward(sd, ref, num_heads, layered=True):
    """sd: {emb_banks, att_banks, weight_raw, q_linear.weight, q_linear.bias}; ref [B, L, n_layer, D] -> [B, L, E]."""
    ref = torch.where(ref != ref, torch.zeros_like(ref), ref)
    B = ref.shape[0]
    if layered:
        w = torch.softmax(sd["weight_raw"].unsqueeze(0), dim=2)
        ref = (w * ref).sum(dim=2)
    E = sd["emb_banks"].shape[1]
    dh = E // num_heads
    q = torch.nn.functional.linear(ref, sd["q_linear.weight"], sd["q_linear.bias"]).view(B, -1, num_heads, dh)
    q = q.transpose(1, 2)
    k = sd["att_banks"].view(-1, num_heads, dh).transpose(0, 1).unsqueeze(0)
    v = sd["emb_banks"].view(-1, num_heads, dh).transpose(0, 1).unsqueeze(0)
    attn = torch.softmax(q @ k.transpose(2, 3) / dh ** 0.5, dim=3)
    return (attn @ v).transpose(1, 2).contiguous().view(B, -1, E)
