"""Process-global switches the reference keeps in its top-level Define.py (Define.py:1-50).

Only what the FastSpeech2 path reads is mirrored: ALLSTATS["global"] (pitch/energy min, max, mean, std
-> the 255 bucket edges, lightning/model/modules.py:40-73), NOLID (fastspeech2m.py:98) and DEVICE.
ALLSTATS is taken from a `stats.json` in the working directory when there is one (what the reference
does at import time, Define.py:15-17); otherwise the defaults below (the values of the reference's
shipped stats.json, pitch then energy) are used.  Set Define.ALLSTATS["global"] before building a
model to override.
"""
import json
import os

import torch

DEBUG = False
NOLID = False
# SSL upstream of the few-shot systems (Define.py:28-48): feature width, number of layers, selected layer
UPSTREAM = "hubert_large_ll60k"
UPSTREAM_DIM = 1024
UPSTREAM_LAYER = 25
LAYER_IDX = None
DEVICE = torch.device("cuda" if torch.cuda.is_available() else "cpu")
ALLSTATS = {
    "global": [56.88630676269531, 953.1358032226562, 186.0852184530204, 46.16604905177577,
               0.0, 533.1392211914062, 51.08978468237829, 40.48262468172912]
}
if os.path.exists("stats.json"):
    try:
        with open("stats.json", "r", encoding="utf-8") as f:
            _s = json.load(f)
        ALLSTATS["global"] = list(_s["pitch"]) + list(_s["energy"])
    except Exception:  # malformed file: keep the defaults
        pass
