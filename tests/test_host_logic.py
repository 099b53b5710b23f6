"""Host-side logic of the boundary package: state_dict contract, sinusoid table, masks, bins (CPU)."""
import torch

from fs2b200 import sub
from oracle import synth
from tests.util_parity import load_golden


def test_state_dict_contract():
    """233 tensors, 34,553,923 trainable parameters, reference key names."""
    M = sub("lightning.model")
    m = M.FastSpeech2(synth.model_cfg())
    sd = m.state_dict()
    assert len(sd) == 233
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == 34553923
    assert sd["encoder.position_enc"].shape == (1, 1001, 256)
    assert sd["decoder.layer_stack.5.pos_ffn.w_1.weight"].shape == (1024, 256, 9)
    assert sd["variance_adaptor.pitch_bins"].shape == (255,)
    assert sd["variance_adaptor.energy_predictor.conv_layer.conv1d_2.conv.weight"].shape == (256, 256, 3)
    assert sd["postnet.convolutions.4.1.running_var"].shape == (80,)
    assert sd["postnet.convolutions.0.1.num_batches_tracked"].dtype == torch.int64
    assert not m.encoder.position_enc.requires_grad and not m.variance_adaptor.pitch_bins.requires_grad
    spk = M.FastSpeech2(synth.model_cfg(multi_speaker=True, multi_lingual=True),
                        spk_config={"emb_type": "table", "speakers": list(range(9))}).state_dict()
    assert spk["speaker_emb.model.weight"].shape == (9, 256)
    assert spk["language_emb.model.weight"].shape == (100, 256)


def test_golden_state_dict_keys_match_reference():
    fx = load_golden("model_small.pt")
    M = sub("lightning.model")
    ours = set(k for k, _ in M.FastSpeech2(fx["cfg"]).named_parameters())
    assert set(fx["grad_digest"]) <= ours


def test_sinusoid_table_matches_reference():
    fx = load_golden("misc.pt")
    tab = sub("transformer.Models").get_sinusoid_encoding_table(37, 256)
    assert torch.equal(tab, fx["sinusoid_37x256"])


def test_pitch_energy_bins():
    """modules.py:40-73 with the shipped stats.json: 255 linspace edges of the normalised range."""
    va = sub("lightning.model.modules").VarianceAdaptor(synth.model_cfg())
    D = sub("Define")
    pmin, pmax, pmean, pstd, emin, emax, emean, estd = D.ALLSTATS["global"]
    assert torch.allclose(va.pitch_bins, torch.linspace((pmin - pmean) / pstd, (pmax - pmean) / pstd, 255))
    assert torch.allclose(va.energy_bins, torch.linspace((emin - emean) / estd, (emax - emean) / estd, 255))


def test_mask_helper():
    tool = sub("lightning.utils.tool")
    m = tool.get_mask_from_lengths(torch.tensor([1, 3]), 4)
    assert m.tolist() == [[False, True, True, True], [False, False, False, True]]
    out = tool.pad([torch.ones(2, 3), torch.ones(5, 3)], 4)
    assert out.shape == (2, 4, 3) and out[0, 2:].abs().sum() == 0  # second entry cropped like F.pad


def test_multilingual_embedding_keys():
    E = sub("lightning.systems.language.embeddings")
    emb = E.MultilingualEmbedding({"en": list(range(40)), "zh": list(range(60)), "xx": []}, 256)
    assert set(emb.state_dict()) == {"tables.table-en", "tables.table-zh"}
    assert emb.tables["table-en"][0].abs().sum() == 0
