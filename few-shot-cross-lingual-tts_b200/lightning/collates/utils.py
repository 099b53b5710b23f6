"""Device-side `reprocess`: the 13-tuple batch of the reference's collate (reference:
lightning/collates/utils.py:8-85, `pad_1D` / `pad_2D` of lightning/utils/tool.py:134-165) with the zero padding done
on the GPU.

The host only concatenates the un-padded per-utterance arrays of each field into one pinned buffer (what the padded
batch would hold minus the padding: ~56 % of the bytes for a LibriTTS-shaped batch) and issues one host -> device copy
per field; `fs2_pad_ragged` then writes the padded `[B, max_len, ...]` tensors.  Same tuple layout and dtypes as the
reference (modes "sup", "unsup", "inference"; "sup" shown): ids, raw_texts, speaker ids (int64) | texts int64 [B,Ts] | src_lens int64 | max_src_len |
mels f32 [B,Tm,n_mel] | mel_lens int64 | max_mel_len | pitches f32 | energies (f32 or f64, as stored) | durations
int64 | lang_ids int64.  Reference-mel speaker slices (`spk_ref_mel_slices`, d-vector speakers) are out of scope.
"""
import numpy as np
import torch

from ... import _cabi


def _ragged_to_padded(arrays, dtype, device, stream):
    """list of [L_i, ...] numpy arrays -> (padded tensor [B, max L, ...] on `device`, lens np.int64 [B])."""
    lens = np.array([a.shape[0] for a in arrays], dtype=np.int64)
    tail = tuple(arrays[0].shape[1:])
    flat = np.concatenate([np.ascontiguousarray(a, dtype=dtype) for a in arrays], axis=0)
    host = torch.from_numpy(flat).pin_memory()
    dev = host.to(device, non_blocking=True)
    offsets = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)).to(device, non_blocking=True)
    B, max_len = len(arrays), int(lens.max())
    out = torch.empty((B, max_len) + tail, dtype=dev.dtype, device=device)
    row_bytes = dev.element_size() * int(np.prod(tail, dtype=np.int64)) if tail else dev.element_size()
    _cabi.check(_cabi.lib().fs2_pad_ragged(dev.data_ptr(), offsets.data_ptr(), B, max_len, row_bytes, out.data_ptr(),
                                           stream), "pad_ragged")
    return out, lens


def reprocess(data, idxs, mode="sup", device=None):
    """Pad the items `idxs` of `data` (dicts with id / speaker / lang_id / text / raw_text / mel / pitch / energy /
    duration, as produced by the reference datasets) into the batch tuple, on `device`.

    mode (lightning/collates/utils.py:8-14): "sup" = the 13-tuple of the training batches; "unsup" = no text
    (slots 1 and 3 are None, text_lens = number of duration segments); "inference" = the 6-tuple
    (ids, raw_texts, speaker_args, texts, text_lens, max_text_len)."""
    if mode not in ("sup", "unsup", "inference"):
        raise NotImplementedError(mode)
    if "spk_ref_mel_slices" in data[0]:
        raise NotImplementedError("reference-mel speaker arguments (d-vector speakers) are not on this path")
    device = torch.device("cuda") if device is None else torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("fs2-b200 collate pads on sm_100a CUDA kernels only (there is no CPU path)")
    stream = torch.cuda.current_stream(device).cuda_stream
    ids = [data[i]["id"] for i in idxs]
    speakers = torch.from_numpy(np.array([data[i]["speaker"] for i in idxs])).long().to(device, non_blocking=True)
    lang_ids = torch.from_numpy(np.array([data[i]["lang_id"] for i in idxs])).long().to(device, non_blocking=True)
    raw_texts = texts = text_lens = None
    with torch.cuda.device(device):
        if mode in ("sup", "inference"):
            raw_texts = [data[i]["raw_text"] for i in idxs]
            texts, text_lens = _ragged_to_padded([data[i]["text"] for i in idxs], np.int64, device, stream)
        if mode == "inference":
            return (ids, raw_texts, speakers, texts, torch.from_numpy(text_lens).to(device, non_blocking=True),
                    max(text_lens))
        mels, mel_lens = _ragged_to_padded([data[i]["mel"] for i in idxs], np.float32, device, stream)
        pitches, _ = _ragged_to_padded([data[i]["pitch"] for i in idxs], np.float32, device, stream)
        e0 = np.asarray(data[idxs[0]]["energy"])
        e_dtype = np.float64 if e0.dtype == np.float64 else np.float32  # the reference does not cast energies
        energies, _ = _ragged_to_padded([data[i]["energy"] for i in idxs], e_dtype, device, stream)
        durations, dur_lens = _ragged_to_padded([data[i]["duration"] for i in idxs], np.int64, device, stream)
    if mode == "unsup":  # one "phoneme" per duration segment (collates/utils.py:35-36)
        text_lens = dur_lens
    return (ids, raw_texts, speakers, texts, torch.from_numpy(text_lens).to(device, non_blocking=True),
            max(text_lens), mels, torch.from_numpy(mel_lens).to(device, non_blocking=True), max(mel_lens), pitches,
            energies, durations, lang_ids)
