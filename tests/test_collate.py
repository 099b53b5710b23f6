"""Device-side collate against the reference's pad_1D / pad_2D semantics
(lightning/utils/tool.py:134-165: zero padding to the longest item, np.stack) and tuple layout
(lightning/collates/utils.py:70-85)."""
import numpy as np
import pytest
import torch

from fs2b200 import sub


def _items(n, seed):
    rng = np.random.default_rng(seed)
    items = []
    for i in range(n):
        L = int(rng.integers(3, 40))
        dur = rng.integers(0, 9, L).astype(np.int64)
        T = max(int(dur.sum()), 1)
        items.append({"id": "utt%d" % i, "raw_text": "text %d" % i, "speaker": int(rng.integers(0, 7)),
                      "lang_id": int(rng.integers(0, 3)), "text": rng.integers(1, 80, L).astype(np.int64),
                      "mel": rng.standard_normal((T, 80)).astype(np.float32),
                      "pitch": rng.standard_normal(L).astype(np.float32),
                      "energy": rng.standard_normal(L).astype(np.float64 if seed % 2 else np.float32),
                      "duration": dur})
    return items


def _pad_ref(arrs):  # pad_1D / pad_2D: zero padding to the longest, stacked
    m = max(a.shape[0] for a in arrs)
    return np.stack([np.pad(a, [(0, m - a.shape[0])] + [(0, 0)] * (a.ndim - 1)) for a in arrs])


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [0, 1])
def test_device_collate_matches_reference_padding(seed):
    C = sub("lightning.collates.utils")
    data = _items(9, seed)
    idxs = [4, 0, 7, 2, 8]
    b = C.reprocess(data, idxs, mode="sup")
    assert len(b) == 13  # lightning/systems/language/FastSpeech2.py:79
    sel = [data[i] for i in idxs]
    assert b[0] == [d["id"] for d in sel] and b[1] == [d["raw_text"] for d in sel]
    assert b[2].dtype == torch.int64 and b[2].tolist() == [d["speaker"] for d in sel]
    for slot, key in ((3, "text"), (6, "mel"), (9, "pitch"), (10, "energy"), (11, "duration")):
        ref = _pad_ref([d[key] for d in sel])
        got = b[slot].cpu().numpy()
        assert got.dtype == ref.dtype and got.shape == ref.shape, (key, got.dtype, ref.dtype)
        assert np.array_equal(got, ref), key  # pure copies + zero fill: bit-exact
    assert b[4].tolist() == [len(d["text"]) for d in sel] and b[5] == max(len(d["text"]) for d in sel)
    assert b[7].tolist() == [d["mel"].shape[0] for d in sel] and b[8] == max(d["mel"].shape[0] for d in sel)
    assert b[12].tolist() == [d["lang_id"] for d in sel]


def test_collate_refuses_cpu():
    C = sub("lightning.collates.utils")
    with pytest.raises(RuntimeError):
        C.reprocess(_items(2, 0), [0, 1], device="cpu")
