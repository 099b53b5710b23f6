"""Re-export of the package's synthetic batch / weight generators (one definition for both sides)."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_m = importlib.import_module("few-shot-cross-lingual-tts_b200.synthetic")
globals().update({k: getattr(_m, k) for k in dir(_m) if not k.startswith("__")})
