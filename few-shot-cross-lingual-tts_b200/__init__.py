"""fs2-b200: B200-native FastSpeech2 training step behind the reference's module signatures.

The directory name contains hyphens (it is fixed by the project layout), so import it with
``importlib.import_module("few-shot-cross-lingual-tts_b200")`` or through the ``fs2b200`` helper at
the repo root.
"""
__version__ = "0.1.0"
