"""Host-side logic on CPU: weight repacking verified by emulating the engines' dataflow from the
PACKED tables (tests/packed_emul.py) against the oracle; drop-in surface (state_dict names, ctor
kwargs, unsupported configurations fail loudly); utterance sharding."""
import pytest
import torch

import jatts_b200
import packed_emul
from jatts_b200 import _pack
from jatts_b200.shard import shard_utterances
from oracle import fs2 as ofs2
from oracle import hifigan as ohg
from oracle import recipes


def test_state_dict_is_reference_compatible():
    for cfg in (recipes.JSUT_FS2, recipes.JVS_FS2, recipes.TINY_FS2):
        m = jatts_b200.FastSpeech2(**cfg)
        want = recipes.fs2_state_shapes(cfg)
        got = m.state_dict()
        assert list(got.keys()) == list(want.keys())
        assert all(tuple(got[k].shape) == tuple(want[k]) for k in want)
        m.load_state_dict(recipes.make_fs2_state_dict(cfg, 0))  # strict


def test_hifigan_state_dict_is_pwg_compatible():
    cfg = recipes.HIFIGAN_V1_HOP300
    g = jatts_b200.HiFiGANGenerator(**cfg)
    keys = [k for k in g.state_dict().keys() if k not in ("mean", "scale")]
    assert keys == list(recipes.hifigan_state_shapes(cfg).keys())
    assert g.hop == 300


@pytest.mark.parametrize("cfg_name,recipe,spk", [("TINY_FS2", "A", False), ("TINY_FS2", "B", False), ("TINY_FS2", "Z", False)])
def test_fs2_repacking_preserves_the_function(cfg_name, recipe, spk):
    cfg = getattr(recipes, cfg_name)
    sd = recipes.make_fs2_state_dict(cfg, seed=4, duration_recipe=recipe)
    m = jatts_b200.FastSpeech2(**cfg)
    packed = _pack.pack_fs2(sd, m._cfg, 256)
    for t, seed in ((9, 1), (23, 2)):
        x = recipes.make_phonemes(t, seed, cfg["idim"])
        ref = ofs2.fs2_inference(sd, cfg, x, return_intermediates=True)
        em = packed_emul.emul_fs2(packed, m._cfg, x)
        assert torch.equal(ref["duration"], em["duration"])
        assert torch.equal(ref["lr_index"], em["lr_index"])
        assert float((ref["feat_gen"] - em["feat_gen"]).abs().max()) < 2e-4
        assert float((ref["pitch"] - em["pitch"]).abs().max()) < 1e-4


def test_fs2_repacking_with_speaker_embedding_and_alpha():
    cfg = dict(recipes.TINY_FS2, spk_embed_dim=192, spk_embed_integration_type="add")
    sd = recipes.make_fs2_state_dict(cfg, seed=6, duration_recipe="A")
    m = jatts_b200.FastSpeech2(**cfg)
    packed = _pack.pack_fs2(sd, m._cfg, 256)
    x = recipes.make_phonemes(14, 3, cfg["idim"])
    sp = recipes.make_spembs(1, 0)[0]
    for alpha in (1.0, 0.7):
        ref = ofs2.fs2_inference(sd, cfg, x, spemb=sp, alpha=alpha)
        em = packed_emul.emul_fs2(packed, m._cfg, x, spemb=sp, alpha=alpha)
        assert torch.equal(ref["duration"], em["duration"])
        assert ref["feat_gen"].shape == em["feat_gen"].shape
        assert float((ref["feat_gen"] - em["feat_gen"]).abs().max()) < 2e-4


def test_hifigan_repacking_preserves_the_function():
    cfg = recipes.HIFIGAN_TINY
    sd = recipes.make_hifigan_state_dict(cfg, 0)
    g = jatts_b200.HiFiGANGenerator(**cfg)
    st = recipes.make_stats(3)
    a, b = 1.0 / st["scale"], -st["mean"] / st["scale"]
    packed = _pack.pack_hifigan(sd, g._cfg, a, b)
    m = recipes.make_mel(13, 1)
    ref = ohg.hifigan_forward(sd, cfg, (m - st["mean"]) / st["scale"]).reshape(-1)
    em = packed_emul.emul_hifigan(packed, g._cfg, m)
    # packed weights are bf16: compare at the bf16-weight level (SNR bound of the GPU test is 35 dB)
    assert ohg.ac_snr_db(ref, em) > 40.0


def test_split16_representation():
    g = torch.Generator().manual_seed(0)
    w = torch.randn(1000, generator=g) * torch.logspace(-6, 3, 1000)
    hi, lo = _pack.split16(w)
    rec = hi.double() + lo.double() / _pack.SPLIT_SCALE
    # 22 significand bits for normal fp16 magnitudes; below 2^-14 the absolute error floor is 2^-35
    bound = 2.0 ** -21 * w.double().abs().clamp_min(2.0 ** -14)
    assert bool(((rec - w.double()).abs() <= bound).all())


def test_glu_interleave_layout():
    d = 128
    w = torch.arange(2 * d, dtype=torch.float32).unsqueeze(1).repeat(1, 4)
    b = torch.arange(2 * d, dtype=torch.float32)
    wi, bi = _pack.glu_interleave(w, b)
    assert bi[:64].tolist() == list(range(64)) and bi[64:128].tolist() == list(range(d, d + 64))
    assert bi[128:192].tolist() == list(range(64, 128)) and torch.equal(wi[:, 0], bi)


def test_unsupported_configurations_fail_loudly():
    base = recipes.JSUT_FS2
    for bad in (dict(encoder_type="transformer"), dict(reduction_factor=2), dict(use_gst=True),
                dict(positionwise_layer_type="linear"), dict(use_macaron_style_in_conformer=False),
                dict(pitch_embed_kernel_size=9), dict(spks=4), dict(conformer_rel_pos_type="latest")):
        with pytest.raises(NotImplementedError):
            jatts_b200.FastSpeech2(**dict(base, **bad))
    with pytest.raises(NotImplementedError):
        jatts_b200.HiFiGANGenerator(**dict(recipes.HIFIGAN_V1_HOP300, out_channels=2))
    with pytest.raises(NotImplementedError):
        jatts_b200.HiFiGANGenerator(**dict(recipes.HIFIGAN_V1_HOP300, upsample_kernel_sizes=(11, 10, 8, 6)))


def test_no_cpu_fallback():
    m = jatts_b200.FastSpeech2(**recipes.TINY_FS2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.inference(torch.tensor([1, 2, 3]))
    with pytest.raises(NotImplementedError):
        m(torch.zeros(1, 3, dtype=torch.long))  # training forward stays with the reference
    g = jatts_b200.HiFiGANGenerator(**recipes.HIFIGAN_TINY)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        g.inference(recipes.make_mel(5, 0))


def test_product_never_imports_the_oracle():
    import os
    import re

    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "jatts_b200")
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"


def test_shard_utterances_partitions_and_balances():
    lens = [300, 10, 250, 40, 40, 500, 5, 90, 120, 33, 77]
    for w in (1, 2, 4, 8):
        shards = shard_utterances(lens, w)
        assert len(shards) == w
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(len(lens)))                      # a partition: every utterance once
        loads = [sum(lens[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(lens)                # LPT bound
    assert shard_utterances(lens, 3) == shard_utterances(lens, 3)  # deterministic
    with pytest.raises(ValueError):
        shard_utterances(lens, 0)


def test_narrow_transposed_convolution_as_one_three_tap_convolution():
    """_pack.py::pack_hifigan: for s * C_out <= 128 the whole ConvTranspose1d(2s, s, p) is ONE 3-tap convolution with
    N = s * C_out whose output row i is frames i*s .. i*s + s - 1 (engine_hifigan.cu uses it for the last stage).
    Evaluated with torch on the PACKED (bf16) weights against F.conv_transpose1d with the same bf16 weights: exact."""
    import torch.nn.functional as F

    for cfg, i in ((recipes.HIFIGAN_V1_HOP300, 3), (recipes.HIFIGAN_TINY, 1)):
        sd = recipes.make_hifigan_state_dict(cfg, 0)
        out = _pack.pack_hifigan(sd, cfg, torch.ones(cfg["in_channels"]), torch.zeros(cfg["in_channels"]))
        s = cfg["upsample_scales"][i]
        ci, co = cfg["channels"] >> i, cfg["channels"] >> (i + 1)
        assert s * co <= 128 and f"ups{i}.t3.hi" in out and tuple(out[f"ups{i}.t3.hi"].shape) == (3, 128, ci)
        assert all(f"ups{j}.t3.hi" not in out for j in range(len(cfg["upsample_scales"])) if cfg["upsample_scales"][j] * (cfg["channels"] >> (j + 1)) > 128)
        w3 = out[f"ups{i}.t3.hi"].double()
        wref = sd[f"upsamples.{i}.1.weight"].to(torch.bfloat16).double()
        x = torch.randn(11, ci, generator=torch.Generator().manual_seed(i), dtype=torch.float64)
        want = F.conv_transpose1d(x.t().unsqueeze(0), wref, sd[f"upsamples.{i}.1.bias"].double(), stride=s,
                                  padding=s // 2 + s % 2, output_padding=s % 2)[0].t()
        z = torch.zeros(1, ci, dtype=torch.float64)
        xp = torch.cat([z, x, z], 0)
        n = s * co
        got = xp[:-2] @ w3[0, :n].t() + xp[1:-1] @ w3[1, :n].t() + xp[2:] @ w3[2, :n].t() + out[f"ups{i}.t3.b"][:n].double()
        assert float(w3[:, n:].abs().max()) == 0.0 and float(out[f"ups{i}.t3.b"][n:].abs().max()) == 0.0
        assert want.shape == (11 * s, co) and float((got.reshape(11 * s, co) - want).abs().max()) < 1e-12
