"""Import the REAL reference FastSpeech2 from /root/reference (this container only).  TEST INFRASTRUCTURE.

``jatts/models/__init__.py`` star-imports Matcha/E2-TTS/VALL-E whose dependencies are missing here,
so a namespace stub for the ``jatts.models`` package is registered before importing the one module we
need.  Nothing on the GPU box may call this: /root/reference is absent there.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("JATTS_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "jatts", "models", "fastspeech2.py"))


def load_reference_fastspeech2():
    """Returns the reference ``FastSpeech2`` class (unmodified code, imported in place)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import jatts  # noqa: F401

    if "jatts.models" not in sys.modules or not hasattr(sys.modules["jatts.models"], "__graft_stub__"):
        pkg = types.ModuleType("jatts.models")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "jatts", "models")]
        pkg.__graft_stub__ = True
        sys.modules["jatts.models"] = pkg
    return importlib.import_module("jatts.models.fastspeech2").FastSpeech2


def build_reference_model(cfg: dict, state_dict):
    """Instantiate the reference model with ``cfg`` kwargs and load ``state_dict`` (strict)."""
    import torch

    cls = load_reference_fastspeech2()
    model = cls(**cfg)
    model.load_state_dict(state_dict, strict=True)
    return model.eval()
