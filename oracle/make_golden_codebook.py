"""Generate tests/golden/codebook.pt by running the reference's own PhonemeQueryExtractor
(lightning/model/reduction.py) and SoftMultiAttCodebook2 (lightning/systems/language/embeddings.py) on seeded
inputs, with outputs and parameter gradients; the oracle restatement must reproduce them first.

Run in the build container:  python oracle/make_golden_codebook.py
Third-party modules missing from the image are stubbed exactly as in oracle/ref_loader.py; the reference's global
`Define` is given a small upstream geometry (dim 64, 5 layers would not fit its hard-coded 25 -> 25 layers, dim 64).
"""
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import codebook_oracle  # noqa: E402

REF = "/root/reference"


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load_reference():
    class LightningModule(nn.Module):
        @property
        def device(self):
            return torch.device("cpu")

    _stub("pytorch_lightning", LightningModule=LightningModule)
    _stub("dlhlp_lib")
    _stub("dlhlp_lib.utils", DataPool=object)
    _stub("dlhlp_lib.utils.numeric", torch_exist_nan=lambda x: bool((x != x).any()))
    t = _stub("text")
    t.__path__ = []
    _stub("text.symbols", symbols=["s%d" % i for i in range(300)])
    _stub("Define", UPSTREAM="hubert_large_ll60k", UPSTREAM_DIM=64, LAYER_IDX=7, UPSTREAM_LAYER=25, DEBUG=False,
          ATTTEMP=False)
    sys.path.insert(0, REF)

    def load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    emb = load("ref_embeddings", "lightning/systems/language/embeddings.py")
    red = load("ref_reduction", "lightning/model/reduction.py")
    sys.path.remove(REF)
    return emb, red


def main():
    emb, red = load_reference()
    g = torch.Generator().manual_seed(9)
    n_symbols, D, n_layer, E, H, C = 11, 64, 25, 32, 4, 16
    # --- phoneme queries from 3 "utterances" of SSL frames
    reps, durs, phs = [], [], []
    for T, L in ((23, 6), (40, 9), (12, 4)):
        d = torch.randint(0, 6, (L,), generator=g).tolist()
        d[-1] += max(0, 3)  # make sure something is non-empty
        T = max(T, sum(d))
        reps.append(torch.randn(T, n_layer, D, generator=g))
        durs.append(d)
        phs.append(torch.randint(0, n_symbols, (L,), generator=g).tolist())
    out = {}
    for two_stage in (True, False):
        ext = red.PhonemeQueryExtractor(mode="average", two_stage=two_stage)
        q = ext([r.clone() for r in reps], durs, n_symbols, phs)
        mine = codebook_oracle.phoneme_query(reps, durs, n_symbols, phs, two_stage)
        assert torch.allclose(q, mine, atol=1e-6), (q - mine).abs().max()
        out["query_two_stage" if two_stage else "query_frame_level"] = q
    # --- codebook attention, forward + gradients
    torch.manual_seed(3)
    cb = emb.SoftMultiAttCodebook2(codebook_size=C, embed_dim=E, num_heads=H)
    ref_in = out["query_two_stage"].clone()
    ref_in[0, 2, 3, 5] = float("nan")  # the reference zeroes NaNs in place
    w = torch.randn(1, n_symbols, E, generator=g)
    y, _ = cb(ref_in.clone())
    (y * w).sum().backward()
    sd = {k: v.detach().clone() for k, v in cb.state_dict().items()}
    mine = codebook_oracle.codebook2_forward(sd, ref_in.clone(), H)
    assert torch.allclose(y, mine, atol=1e-5), (y - mine).abs().max()
    fx = {"cfg": dict(n_symbols=n_symbols, D=D, n_layer=n_layer, E=E, H=H, C=C, layer_idx=7),
          "reps": reps, "durs": durs, "phs": phs, **out,
          "codebook_sd": sd, "codebook_in": ref_in, "codebook_w": w, "codebook_out": y.detach(),
          "codebook_grads": {k: p.grad.clone() for k, p in cb.named_parameters() if p.grad is not None}}
    path = os.path.join(ROOT, "tests", "golden", "codebook.pt")
    torch.save(fx, path)
    print("wrote", path, os.path.getsize(path), "bytes; grads:", list(fx["codebook_grads"]))


if __name__ == "__main__":
    main()
