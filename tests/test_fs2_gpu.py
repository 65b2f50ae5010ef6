"""GPU parity of the FastSpeech2 path (through the public class -> C ABI): batched output row i must
equal the reference's single-utterance inference(x_i) -- durations and LengthRegulator indices
bit-exact, mel within 1e-3 max-abs (north_star tolerance), pitch/energy within 1e-3."""
import os
import sys

import numpy as np
import pytest
import torch

import jatts_b200
from oracle import fs2 as ofs2
from oracle import recipes

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import make_golden  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
MEL_TOL = 1e-3   # north_star: mel within 1e-3 max-abs
VAR_TOL = 1e-3

_MODELS = {}


def get_model(cfg_name, wseed, recipe):
    key = (cfg_name, wseed, recipe)
    if key not in _MODELS:
        cfg = getattr(recipes, cfg_name)
        sd = recipes.make_fs2_state_dict(cfg, seed=wseed, duration_recipe=recipe)
        m = jatts_b200.FastSpeech2(**cfg)
        m.load_state_dict(sd)
        _MODELS.clear()  # one resident model at a time
        _MODELS[key] = (m.eval().to("cuda"), sd, cfg)
    return _MODELS[key]


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(make_golden.CASES))
def test_matches_golden_vectors_of_the_real_reference(name):
    cfg_name, wseed, recipe, lens, tseed, alpha = make_golden.CASES[name]
    model, sd, cfg = get_model(cfg_name, wseed, recipe)
    _, _, _, texts, spembs, _ = make_golden.case_inputs(name)
    outs = model.inference_batch(texts, spembs=spembs, alpha=alpha)
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    for i, o in enumerate(outs):
        assert o["duration"].cpu().tolist() == z[f"duration_{i}"].tolist(), f"{name}[{i}] durations"   # bit-exact
        assert tuple(o["feat_gen"].shape) == z[f"feat_gen_{i}"].shape
        assert np.abs(o["feat_gen"].cpu().numpy() - z[f"feat_gen_{i}"]).max() < MEL_TOL
        assert np.abs(o["pitch"].cpu().numpy() - z[f"pitch_{i}"]).max() < VAR_TOL
        assert np.abs(o["energy"].cpu().numpy() - z[f"energy_{i}"]).max() < VAR_TOL
        assert o["duration"].dtype == torch.int64 and o["pitch"].shape == (len(texts[i]), 1)


@pytest.mark.gpu
def test_single_utterance_inference_signature():
    """reference call: model.inference(x, spembs=None) -> dict (tts_decode.py:230)"""
    model, sd, cfg = get_model("JSUT_FS2", 0, "A")
    x = recipes.make_phonemes(50, 1234, cfg["idim"]).to("cuda")
    out = model.inference(x)
    z = np.load(os.path.join(GOLDEN, "fs2_jsut_A.npz"))
    assert set(out) == {"feat_gen", "duration", "pitch", "energy"}
    assert out["duration"].cpu().tolist() == z["duration_0"].tolist()
    assert np.abs(out["feat_gen"].cpu().numpy() - z["feat_gen_0"]).max() < MEL_TOL


@pytest.mark.gpu
def test_ragged_batch_rows_equal_per_utterance_oracle():
    """variable lengths incl. T=1, zero-duration tokens (recipe B); batch row == oracle(x_i)"""
    model, sd, cfg = get_model("JSUT_FS2", 5, "B")
    lens = [1, 2, 3, 17, 64, 5, 33, 80, 9, 50, 21, 8]
    texts = [recipes.make_phonemes(t, 900 + i, cfg["idim"]) for i, t in enumerate(lens)]
    outs = model.inference_batch(texts, return_lr_index=True)
    torch.set_num_threads(os.cpu_count() or 1)
    for x, o in zip(texts, outs):
        ref = ofs2.fs2_inference(sd, cfg, x, return_intermediates=True)
        assert torch.equal(ref["duration"], o["duration"].cpu())
        assert torch.equal(ref["lr_index"].int(), o["lr_index"].cpu())            # bit-exact indices
        assert float((ref["feat_gen"] - o["feat_gen"].cpu()).abs().max()) < MEL_TOL


@pytest.mark.gpu
def test_batch_composition_does_not_change_an_utterance():
    """no leakage between neighbours in the packed layout: same utterance alone vs inside a batch"""
    model, sd, cfg = get_model("JSUT_FS2", 0, "A")
    texts = [recipes.make_phonemes(t, 70 + i, cfg["idim"]) for i, t in enumerate([30, 50, 12, 41])]
    alone = model.inference_batch([texts[1]])[0]
    batch = model.inference_batch(texts)[1]
    assert torch.equal(alone["duration"], batch["duration"])
    assert float((alone["feat_gen"] - batch["feat_gen"]).abs().max()) < 2e-4


@pytest.mark.gpu
def test_full_size_batch_properties():
    """BASELINE config 2 size (64 x 50 phonemes): frames == sum(durations), deterministic across calls"""
    model, sd, cfg = get_model("JSUT_FS2", 0, "A")
    texts = [recipes.make_phonemes(50, i, cfg["idim"]) for i in range(64)]
    a = model.inference_batch(texts, return_lr_index=True)
    b = model.inference_batch(texts)
    for x, oa, ob in zip(texts, a, b):
        assert int(oa["duration"].sum()) == oa["feat_gen"].shape[0]
        lr = oa["lr_index"].cpu()
        assert torch.equal(lr, torch.repeat_interleave(torch.arange(len(x)), oa["duration"].cpu()).int())
        assert bool((lr[1:] >= lr[:-1]).all())                                      # sortedness
        assert torch.equal(oa["feat_gen"], ob["feat_gen"])                          # idempotence
        assert torch.isfinite(oa["feat_gen"]).all()


@pytest.mark.gpu
def test_error_behaviour():
    model, sd, cfg = get_model("JSUT_FS2", 0, "A")
    with pytest.raises(IndexError):
        model.inference(torch.tensor([1, 2, cfg["idim"]]))          # nn.Embedding would raise IndexError
    with pytest.raises(ValueError):
        model.inference(torch.tensor([1, 2, 3]), spembs=torch.zeros(192))   # model has no spk_embed_dim
    with pytest.raises(NotImplementedError):
        model.inference(torch.tensor([1, 2, 3]), use_teacher_forcing=True)
    with pytest.raises(ValueError):
        model.inference(torch.tensor([1, 2, 3]), alpha=0.0)


@pytest.mark.gpu
def test_cpu_model_refuses_to_run():
    cfg = recipes.TINY_FS2
    m = jatts_b200.FastSpeech2(**cfg)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.inference(torch.tensor([1, 2, 3]))


@pytest.mark.gpu
def test_duration_flip_rate_over_ten_thousand_tokens():
    """VERDICT r1 weak #2 / SURVEY 8(d) recipe B: >= 10 000 tokens with the UN-damped duration predictor
    (durations 0..50+).  Tokens whose fp64 ``exp(x) - 1`` lies within 1e-4 (relative to 1 + value) of a
    rounding boundary are identified with the fp64 oracle and excluded -- there fp32 arithmetic may round
    either way; on every other token the CUDA path must equal the fp64 oracle's duration: zero flips."""
    model, sd, cfg = get_model("JSUT_FS2", 5, "B")
    g = torch.Generator().manual_seed(2024)
    lens = torch.randint(20, 81, (210,), generator=g).tolist()
    texts = [recipes.make_phonemes(t, 5000 + i, cfg["idim"]) for i, t in enumerate(lens)]
    assert sum(lens) >= 10_000
    got = []
    for s in range(0, len(texts), 42):
        got += [o["duration"].cpu() for o in model.inference_batch(texts[s:s + 42])]
    torch.set_num_threads(os.cpu_count() or 1)
    n_tok = n_excl = n_flip = n_flip32 = 0
    for x, d in zip(texts, got):
        logd64 = ofs2.fs2_log_durations(sd, cfg, x, dtype=torch.float64)
        v = logd64.exp() - 1.0
        want = torch.clamp(torch.round(v), min=0).long()
        if int(want.sum()) == 0:
            want = torch.ones_like(want)
        near = ((v - torch.floor(v) - 0.5).abs() < 1e-4 * (1.0 + v.abs())) & (v > -0.5 - 1e-4)
        n_tok += x.numel()
        n_excl += int(near.sum())
        n_flip += int(((want != d) & ~near).sum())
        d32 = ofs2.duration_from_log(ofs2.fs2_log_durations(sd, cfg, x))
        n_flip32 += int((d32 != want).sum())     # informational: the fp32 oracle against the fp64 one
    print(f"tokens {n_tok}, excluded near a .5 boundary {n_excl}, CUDA flips on the rest {n_flip}, "
          f"fp32-oracle flips (all tokens) {n_flip32}")
    assert n_tok >= 10_000 and n_excl < n_tok // 100
    assert n_flip == 0


@pytest.mark.gpu
@pytest.mark.parametrize("n_tokens", [100, 200, 320])
def test_long_utterances_match_the_oracle(n_tokens):
    """VERDICT r1 weak #4 / ADVICE: ~600, ~1200 and ~2000 decoder frames (the old attention kernel changed
    implementation above ~500 frames); one attention kernel serves every length.  Durations bit-exact, mel < 1e-3."""
    model, sd, cfg = get_model("JSUT_FS2", 0, "A")
    texts = [recipes.make_phonemes(n_tokens, 7000 + n_tokens, cfg["idim"]), recipes.make_phonemes(23, 7001, cfg["idim"])]
    outs = model.inference_batch(texts, return_lr_index=True)
    torch.set_num_threads(os.cpu_count() or 1)
    for x, o in zip(texts, outs):
        ref = ofs2.fs2_inference(sd, cfg, x, return_intermediates=True)
        assert torch.equal(ref["duration"], o["duration"].cpu())
        assert torch.equal(ref["lr_index"].int(), o["lr_index"].cpu())
        assert float((ref["feat_gen"] - o["feat_gen"].cpu()).abs().max()) < MEL_TOL
    assert outs[0]["feat_gen"].shape[0] > 5.5 * n_tokens
