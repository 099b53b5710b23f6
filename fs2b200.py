"""Import helper: ``from fs2b200 import pkg`` -> the `few-shot-cross-lingual-tts_b200` package.

(The package directory name has hyphens, so the ``import`` statement cannot spell it.)
"""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

PKG_NAME = "few-shot-cross-lingual-tts_b200"
pkg = importlib.import_module(PKG_NAME)


def sub(name):
    """Import a submodule of the package, e.g. sub("transformer.Models")."""
    return importlib.import_module(PKG_NAME + "." + name)
