"""Parity of the Matcha-TTS CUDA path (jatts_b200.MatchaTTS -> C ABI -> engine_matcha.cu; SURVEY 8f-2, BASELINE
config 5) against the oracle (oracle/matcha.py, pinned to the reference's own modules by tests/test_oracle_matcha.py).

Row i of a batch must equal the reference's single-utterance ``inference(x_i)`` given the same noise: durations
bit-exact, mel within the tolerance stated per test (fp32-faithful split-operand GEMMs; the 10 Euler steps feed their
own output back, so the budget is that of the FastSpeech2 mel, 1e-3, on outputs of magnitude O(1))."""
import pytest
import torch

import jatts_b200
from oracle import matcha as om
from oracle import recipes

_MODELS = {}


def get_model(cfg_name, seed, recipe="A"):
    key = (cfg_name, seed, recipe)
    if key not in _MODELS:
        cfg = getattr(recipes, cfg_name)
        sd = recipes.make_matcha_state_dict(cfg, seed, recipe)
        m = jatts_b200.MatchaTTS(**cfg)
        m.load_state_dict(sd)
        _MODELS[key] = (m.eval().to("cuda"), sd, cfg)
    return _MODELS[key]


def run_and_compare(model, sd, cfg, texts, steps, temperature, seed):
    noise = lambda frames: [recipes.make_noise(f, cfg["odim"], seed + i) for i, f in enumerate(frames)]
    outs = model.inference_batch(texts, n_timesteps=steps, temperature=temperature, noise=noise)
    worst = 0.0
    for i, (x, o) in enumerate(zip(texts, outs)):
        frames = int(o["feat_gen"].shape[0])
        z = recipes.make_noise(frames, cfg["odim"], seed + i).t()          # the oracle takes (odim, T)
        ref = om.matcha_inference(sd, cfg, x, z, steps, temperature)
        assert torch.equal(o["duration"].cpu(), ref["duration"]), f"durations differ (row {i})"
        assert tuple(o["feat_gen"].shape) == tuple(ref["feat_gen"].shape) and frames % 2 == 0
        err = float((o["feat_gen"].cpu() - ref["feat_gen"]).abs().max())
        assert err == err, "non-finite output"
        worst = max(worst, err)
    return worst


@pytest.mark.gpu
@pytest.mark.parametrize("steps", [1, 2, 10])
def test_small_model_batch_rows_equal_per_utterance_oracle(steps):
    model, sd, cfg = get_model("SMALL_MATCHA", 1)
    texts = [recipes.make_phonemes(t, 40 + i, cfg["idim"]) for i, t in enumerate([7, 1, 12, 3, 20])]
    err = run_and_compare(model, sd, cfg, texts, steps, 0.667, seed=steps)
    print(f"steps {steps}: mel max-abs error {err:.3e}")
    assert err < 1e-3, err


@pytest.mark.gpu
def test_long_and_short_utterances_in_one_batch():
    """~900 frames next to ~20: several 127-row attention tiles at both rates, GroupNorm statistics over long rows, and the
    per-utterance boundaries of every convolution in between"""
    model, sd, cfg = get_model("SMALL_MATCHA", 1)
    texts = [recipes.make_phonemes(t, 140 + i, cfg["idim"]) for i, t in enumerate([150, 3, 64])]
    err = run_and_compare(model, sd, cfg, texts, 2, 0.667, seed=77)
    print(f"long / short batch: mel max-abs error {err:.3e}")
    assert err < 1e-3, err


@pytest.mark.gpu
def test_single_utterance_inference_signature_and_batch_independence():
    model, sd, cfg = get_model("SMALL_MATCHA", 1)
    xs = [recipes.make_phonemes(t, 60 + i, cfg["idim"]) for i, t in enumerate([9, 4, 15])]
    noise = lambda frames: [recipes.make_noise(f, cfg["odim"], 300 + f) for f in frames]    # keyed by length: same z alone / batched
    batch = model.inference_batch(xs, n_timesteps=3, temperature=0.5, noise=noise)
    for x, b in zip(xs, batch):
        alone = model.inference_batch([x], n_timesteps=3, temperature=0.5, noise=noise)[0]
        assert torch.equal(alone["duration"], b["duration"])
        assert torch.equal(alone["feat_gen"], b["feat_gen"]), "an utterance's output depends on its batch neighbours"
    out = model.inference(xs[0], n_timesteps=3, temperature=0.5)     # reference call signature; noise drawn on the device
    assert set(out) == {"feat_gen", "duration"} and out["feat_gen"].shape == batch[0]["feat_gen"].shape
    assert out["feat_gen"].dtype == torch.float32 and out["duration"].dtype == torch.long
    assert bool(torch.isfinite(out["feat_gen"]).all())


@pytest.mark.gpu
def test_recipe_size_model_matches_the_oracle():
    """egs/jsut/tts1/conf/matcha_tts.v1.prior.steplr.large.yaml: 512-wide decoder, 2 heads of 256, 10 Euler steps"""
    model, sd, cfg = get_model("JSUT_MATCHA", 2)
    texts = [recipes.make_phonemes(t, 80 + i, cfg["idim"]) for i, t in enumerate([11, 30])]
    err = run_and_compare(model, sd, cfg, texts, recipes.MATCHA_ODE_STEPS, recipes.MATCHA_TEMPERATURE, seed=9)
    print(f"recipe-size model: mel max-abs error {err:.3e}")
    assert err < 1e-3, err


@pytest.mark.gpu
def test_full_size_batch_properties():
    """BASELINE config 5 at size (batch 64 x ~50 phonemes): frames = even-truncated sum of durations, finite output,
    idempotence (the same call twice gives the same bytes)"""
    model, sd, cfg = get_model("JSUT_MATCHA", 2)
    texts = [recipes.make_phonemes(50, 900 + i, cfg["idim"]) for i in range(64)]
    noise = lambda frames: [recipes.make_noise(f, cfg["odim"], i) for i, f in enumerate(frames)]
    a = model.inference_batch(texts, n_timesteps=10, temperature=0.667, noise=noise)
    b = model.inference_batch(texts, n_timesteps=10, temperature=0.667, noise=noise)
    for oa, ob in zip(a, b):
        tot = int(oa["duration"].sum())
        assert int(oa["feat_gen"].shape[0]) == tot - tot % 2
        assert bool(torch.isfinite(oa["feat_gen"]).all())
        assert torch.equal(oa["feat_gen"], ob["feat_gen"])


@pytest.mark.gpu
def test_error_behaviour():
    model, sd, cfg = get_model("SMALL_MATCHA", 1)
    with pytest.raises(ValueError):
        model.inference(recipes.make_phonemes(5, 1, cfg["idim"]))                      # n_timesteps / temperature missing
    with pytest.raises(IndexError):
        model.inference_batch([torch.tensor([1, cfg["idim"] + 3])], n_timesteps=2, temperature=0.5)
    with pytest.raises(ValueError):
        model.inference_batch([recipes.make_phonemes(5, 1, cfg["idim"])], n_timesteps=2, temperature=0.5,
                              noise=[torch.zeros(3, cfg["odim"])])
    cpu = jatts_b200.MatchaTTS(**cfg)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cpu.inference(recipes.make_phonemes(5, 1, cfg["idim"]), n_timesteps=2, temperature=0.5)
