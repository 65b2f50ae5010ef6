"""Host logic of the batched stage-4 front-end (jatts_b200/decode.py), no GPU: token conversion and csv reading
as jatts/utils/token_id_converter.py + jatts/datasets/tts_dataset.py:93-116, batch planning, the PCM_16 writer."""
import wave

import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from jatts_b200 import decode


def test_token_converter_matches_reference_semantics(tmp_path):
    p = tmp_path / "tokens.txt"
    p.write_text("<blank>\n<unk>\na\nb\nky\n<sos/eos>\n", encoding="utf-8")
    conv = decode.TokenIDConverter(str(p))
    assert conv.tokens2ids(["a", "ky", "zz", "b"]) == [2, 4, 1, 3]     # unknown -> <unk>
    with pytest.raises(RuntimeError):
        decode.TokenIDConverter(["a", "a", "<unk>"])                    # duplicated symbol
    with pytest.raises(RuntimeError):
        decode.TokenIDConverter(["a", "b"])                             # no <unk>


def test_read_items_splits_on_single_spaces(tmp_path):
    (tmp_path / "tokens.txt").write_text("<blank>\n<unk>\na\nb\n", encoding="utf-8")
    (tmp_path / "d.csv").write_text("sample_id,phonemes,spk\nu1,a b  a,s1\nu2,b q,s2\n", encoding="utf-8")
    items = decode.read_items(str(tmp_path / "d.csv"), "phonemes", decode.TokenIDConverter(str(tmp_path / "tokens.txt")))
    assert [it["sample_id"] for it in items] == ["u1", "u2"]
    assert items[0]["token_indices"] == [2, 3, 2] and items[0]["tokens"] == ["a", "b", "a"]   # empty fields dropped
    assert items[1]["token_indices"] == [3, 1] and items[1]["spk"] == "s2"
    with pytest.raises(KeyError):
        decode.read_items(str(tmp_path / "d.csv"), "characters", decode.TokenIDConverter(str(tmp_path / "tokens.txt")))


@settings(max_examples=60, deadline=None)
@given(st.lists(st.integers(1, 200), min_size=0, max_size=300), st.integers(1, 64), st.integers(1, 4000))
def test_plan_batches_properties(lengths, max_utts, max_tokens):
    batches = decode.plan_batches(lengths, max_utts, max_tokens)
    flat = [j for b in batches for j in b]
    assert sorted(flat) == list(range(len(lengths)))                   # every utterance exactly once
    assert [lengths[j] for j in flat] == sorted(lengths)               # length-bucketed
    for b in batches:
        assert 1 <= len(b) <= max_utts
        assert sum(lengths[j] for j in b) <= max_tokens or len(b) == 1  # a single over-long utterance still runs


def test_plan_batches_rejects_empty_utterances():
    with pytest.raises(ValueError):
        decode.plan_batches([3, 0, 2])
    with pytest.raises(ValueError):
        decode.plan_batches([3], max_utts=0)


def test_wav_writer_round_trip(tmp_path):
    x = (np.sin(np.arange(2400) * 0.05) * 30000).astype(np.int16)
    p = str(tmp_path / "a.wav")
    decode.write_wav_pcm16(p, x, 24000)
    with wave.open(p, "rb") as w:
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (1, 2, 24000, 2400)
        assert np.array_equal(np.frombuffer(w.readframes(2400), dtype="<i2"), x)


def test_read_stats_npz_and_dict(tmp_path):
    m, s = np.arange(80, dtype=np.float32), np.ones(80, dtype=np.float32) * 2
    np.savez(tmp_path / "stats.npz", mel_mean=m, mel_scale=s)
    a, b = decode.read_stats(str(tmp_path / "stats.npz"), "mel")
    assert np.array_equal(a, m) and np.array_equal(b, s)
    a, b = decode.read_stats({"mel_mean": m, "mel_scale": s}, "mel")
    assert a is m and b is s


def test_speaker_embedding_cache_calls_the_extractor_once_per_reference():
    """tts_decode.py:209-212 runs the extractor per utterance; the cache runs it per distinct ref_wav_path"""
    import torch
    calls = []

    def extractor(path):
        calls.append(path)
        return torch.full((192,), float(len(path)))

    cache = decode.SpeakerEmbeddingCache(table={"uttX": torch.zeros(192)}, extractor=extractor)
    items = [{"sample_id": f"u{i}", "ref_wav_path": f"/spk{i % 2}.wav"} for i in range(6)] + [{"sample_id": "uttX", "ref_wav_path": "/z.wav"}]
    embs = [cache(it) for it in items]
    assert calls == ["/spk0.wav", "/spk1.wav"] and cache.calls == 2
    assert float(embs[-1].abs().sum()) == 0.0 and embs[0].shape == (192,)
    with pytest.raises(KeyError):
        decode.SpeakerEmbeddingCache()({"sample_id": "a"})


def test_pinned_ring_reuses_released_buffers():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("pinned memory needs a CUDA runtime")
    ring = decode.PinnedRing(torch.int16, slots=2)
    k0, a = ring.take(100)
    k1, b = ring.take(50)
    assert a.is_pinned() and a.numel() == 100 and k0 != k1
    ring.release(k0)
    k2, c = ring.take(80)
    assert k2 == k0 and c.data_ptr() == a.data_ptr()
