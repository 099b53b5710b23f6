"""Device-side collate against the REFERENCE's own `reprocess` / pad_1D / pad_2D
(lightning/collates/utils.py:8-111, lightning/utils/tool.py:134-165), executed from the reference sources by
oracle/make_golden2.py::run_collate -> tests/golden/collate.pt; all three modes, tuple layout and dtypes included."""
import numpy as np
import pytest
import torch

from fs2b200 import sub
from tests.util_parity import load_golden


def _same(got, ref):
    if ref is None:
        return got is None
    if torch.is_tensor(ref):
        return torch.is_tensor(got) and got.dtype == ref.dtype and got.shape == ref.shape and \
            torch.equal(got.cpu(), ref)  # pure copies + zero fill: bit-exact
    if isinstance(ref, (list, tuple)):
        return list(got) == list(ref)
    return int(got) == int(ref)


@pytest.mark.gpu
@pytest.mark.parametrize("case", [0, 1])
@pytest.mark.parametrize("mode", ["sup", "unsup", "inference"])
def test_device_collate_matches_reference_reprocess(case, mode):
    C = sub("lightning.collates.utils")
    fx = load_golden("collate.pt")[case]
    got = C.reprocess(fx["items"], fx["idxs"], mode=mode)
    ref = fx["out"][mode]
    assert len(got) == len(ref) == (6 if mode == "inference" else 13)  # FastSpeech2.py:79 asserts 13
    for slot, (g, r) in enumerate(zip(got, ref)):
        assert _same(g, r), (mode, slot, g, r)


def test_collate_refuses_cpu_and_unknown_modes():
    C = sub("lightning.collates.utils")
    item = {"id": "a", "raw_text": "t", "speaker": 0, "lang_id": 0, "text": np.arange(3), "mel": np.zeros((4, 80)),
            "pitch": np.zeros(3), "energy": np.zeros(3), "duration": np.ones(3, dtype=np.int64)}
    with pytest.raises(RuntimeError):
        C.reprocess([item, item], [0, 1], device="cpu")
    with pytest.raises(NotImplementedError):
        C.reprocess([item], [0], mode="semi")
