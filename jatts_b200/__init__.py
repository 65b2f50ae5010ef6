"""jatts_b200 -- B200-native (sm_100a) batched synthesis path for JATTS:
FastSpeech2 (and Matcha-TTS) text2mel inference -> HiFi-GAN V1 vocoder, behind the reference's own call signatures.

Importing this package loads the CUDA library through its C ABI (include/jatts_b200.h) and fails
loudly if it has not been built; there is no CPU fallback anywhere in the package.
"""
from . import _lib  # noqa: F401  (raises if libjatts_b200.so is missing)
from .fastspeech2 import FastSpeech2
from .matchatts import MatchaTTS
from .vocoder import HiFiGANGenerator, Vocoder
from .shard import shard_utterances

__all__ = ["FastSpeech2", "MatchaTTS", "HiFiGANGenerator", "Vocoder", "shard_utterances"]
