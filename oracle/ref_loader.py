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


def _install_matcha_standins():
    """``conformer`` and ``diffusers`` (unpinned third-party dependencies of the reference's Matcha modules, setup.cfg:40-41)
    are not installed here.  Inert stand-ins are registered so that the REAL reference modules import; the one class that
    does arithmetic, ``diffusers...Attention``, is replaced by the restatement documented in oracle/matcha.py -- what the
    pin test then compares is the reference's own code (Decoder, ResnetBlock1D, SnakeBeta, BasicTransformerBlock, CFM,
    MatchaTTS._forward) against the oracle."""
    import torch
    import torch.nn as nn

    if "diffusers" in sys.modules and hasattr(sys.modules["diffusers"], "__graft_stub__"):
        return

    class ConformerBlock(nn.Module):  # decoder.py:206-246 subclasses it; block type "conformer" is not used by any recipe
        def __init__(self, **kw):
            super().__init__()

    conformer = types.ModuleType("conformer")
    conformer.ConformerBlock = ConformerBlock
    sys.modules["conformer"] = conformer

    class Attention(nn.Module):
        """diffusers.models.attention_processor.Attention, self-attention subset [diffusers, unpinned]"""

        def __init__(self, query_dim, heads=8, dim_head=64, dropout=0.0, bias=False, cross_attention_dim=None,
                     upcast_attention=False, **kw):
            super().__init__()
            inner = heads * dim_head
            self.heads, self.scale = heads, dim_head ** -0.5
            self.to_q = nn.Linear(query_dim, inner, bias=bias)
            self.to_k = nn.Linear(query_dim, inner, bias=bias)
            self.to_v = nn.Linear(query_dim, inner, bias=bias)
            self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(dropout)])

        def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kw):
            b, t, _ = hidden_states.shape
            sp = lambda z: z.view(b, t, self.heads, -1).transpose(1, 2)
            q, k, v = sp(self.to_q(hidden_states)), sp(self.to_k(hidden_states)), sp(self.to_v(hidden_states))
            s = torch.matmul(q, k.transpose(-2, -1)) * self.scale
            if attention_mask is not None:   # additive, as diffusers applies it
                s = s + attention_mask.to(s.dtype)[:, None, None, :]
            o = torch.matmul(torch.softmax(s, dim=-1), v).transpose(1, 2).reshape(b, t, -1)
            return self.to_out[1](self.to_out[0](o))

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    dummy = type("Unused", (nn.Module,), {})
    d = mod("diffusers", __graft_stub__=True)
    d.models = mod("diffusers.models")
    mod("diffusers.models.attention", GEGLU=dummy, GELU=dummy, AdaLayerNorm=dummy, AdaLayerNormZero=dummy, ApproximateGELU=dummy)
    mod("diffusers.models.attention_processor", Attention=Attention)
    mod("diffusers.models.lora", LoRACompatibleLinear=nn.Linear)
    d.utils = mod("diffusers.utils")
    mod("diffusers.utils.torch_utils", maybe_allow_in_graph=lambda cls: cls)


def load_reference_matcha():
    """Returns the reference ``MatchaTTS`` class (unmodified code, imported in place, with the stand-ins above)."""
    load_reference_fastspeech2()     # registers the jatts.models namespace stub
    _install_matcha_standins()
    return importlib.import_module("jatts.models.matchatts").MatchaTTS
