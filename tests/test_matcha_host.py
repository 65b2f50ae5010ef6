"""Host-side checks of the Matcha-TTS drop-in (BASELINE config 5) that need no GPU: the class is
state_dict compatible with the reference, and the weight repacking of jatts_b200/_pack.py::pack_matcha (stride-2
convolution as a 2-tap convolution over frame pairs, ConvTranspose1d(4, 2, 1) as a 3-tap convolution emitting frame
pairs, stacked q / k / v, SnakeBeta constants, the per-step time-embedding table) preserves the function."""
import pytest
import torch
import torch.nn.functional as F

import jatts_b200
from jatts_b200 import _pack
from oracle import matcha as om
from oracle import recipes, ref_loader


def unsplit(out, name):
    return out[name + ".hi"].double() + out[name + ".lo"].double() / _pack.SPLIT_SCALE


def test_state_dict_matches_the_recipe_layout():
    for cfg in (recipes.SMALL_MATCHA, recipes.JSUT_MATCHA, dict(recipes.SMALL_MATCHA, spk_embed_dim=24)):
        m = jatts_b200.MatchaTTS(**cfg)
        want = recipes.matcha_state_shapes(cfg)
        got = m.state_dict()
        assert list(got.keys()) == list(want.keys())
        assert all(tuple(got[k].shape) == tuple(want[k]) for k in want)
        if not cfg.get("spk_embed_dim"):
            m.load_state_dict(recipes.make_matcha_state_dict(cfg, 0))  # strict


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not present (GPU box)")
def test_state_dict_is_reference_compatible():
    """key names, order and shapes against the REAL reference class (matchatts.py:30) built with the same kwargs"""
    import logging

    logging.disable(logging.WARNING)
    try:
        cls = ref_loader.load_reference_matcha()
        for cfg in (recipes.SMALL_MATCHA, dict(recipes.SMALL_MATCHA, spk_embed_dim=24, decoder_n_blocks=2, decoder_num_mid_blocks=1)):
            ref = cls(**cfg).state_dict()
            got = jatts_b200.MatchaTTS(**cfg).state_dict()
            assert list(got.keys()) == list(ref.keys())
            assert all(tuple(got[k].shape) == tuple(ref[k].shape) for k in ref)
    finally:
        logging.disable(logging.NOTSET)


def test_unsupported_configurations_fail_loudly():
    for bad in (dict(decoder_channels=[64, 128]), dict(decoder_attention_head_dim=32), dict(decoder_act_fn="gelu"),
                dict(encoder_type="transformer"), dict(decoder_channels=[64, 64, 64])):
        with pytest.raises(NotImplementedError):
            jatts_b200.MatchaTTS(**dict(recipes.SMALL_MATCHA, **bad))


def test_paired_row_views_of_the_resampling_convolutions():
    """the two layout tricks of engine_matcha.cu evaluated with torch on the PACKED weights"""
    cfg = recipes.SMALL_MATCHA
    sd = recipes.make_matcha_state_dict(cfg, 3)
    m = jatts_b200.MatchaTTS(**cfg)
    out = _pack.pack_matcha(sd, m._cfg, 64)
    c = cfg["decoder_channels"][0]
    g = torch.Generator().manual_seed(1)
    T = 10
    x = torch.randn(T, c, generator=g, dtype=torch.float64)                        # (frames, channels), even length
    e = "decoder.estimator."
    # ---- Downsample1D: Conv1d(k3, stride 2, padding 1)
    want = F.conv1d(x.t().unsqueeze(0), sd[e + "down_blocks.0.2.conv.weight"].double(), sd[e + "down_blocks.0.2.conv.bias"].double(),
                    stride=2, padding=1)[0].t()
    w = unsplit(out, "dec.down0")                                                   # [2 taps, N_pad, 2C]
    pairs = torch.cat([torch.zeros(1, 2 * c, dtype=torch.float64), x.reshape(T // 2, 2 * c)], 0)   # row j+1 = pair j; row 0 = zero gap
    got = pairs[:-1] @ w[0, :c].t() + pairs[1:] @ w[1, :c].t() + out["dec.down0.b"].double()
    assert float((got - want).abs().max()) < 1e-5
    # ---- Upsample1D: ConvTranspose1d(4, 2, 1)
    xh = x[: T // 2]
    want = F.conv_transpose1d(xh.t().unsqueeze(0), sd[e + "up_blocks.0.2.conv.weight"].double(), sd[e + "up_blocks.0.2.conv.bias"].double(),
                              stride=2, padding=1)[0].t()                           # (T, C)
    w = unsplit(out, "dec.up0")                                                     # [3 taps, 2C (padded), C]
    z = torch.zeros(1, c, dtype=torch.float64)
    xp = torch.cat([z, xh, z], 0)
    got = xp[:-2] @ w[0, :2 * c].t() + xp[1:-1] @ w[1, :2 * c].t() + xp[2:] @ w[2, :2 * c].t() + out["dec.up0.b"].double()
    assert float((got.reshape(T, c) - want).abs().max()) < 1e-5


def test_time_table_equals_the_oracle_embedding():
    cfg = recipes.SMALL_MATCHA
    sd = recipes.make_matcha_state_dict(cfg, 5)
    m = jatts_b200.MatchaTTS(**cfg)
    m.load_state_dict(sd)
    steps = 4
    temb, dts = m._time_table(steps, torch.device("cpu"))
    names = _pack.matcha_resnet_names(cfg["decoder_num_mid_blocks"])
    assert tuple(temb.shape) == (steps, len(names), cfg["decoder_channels"][0])
    t_span = torch.linspace(0, 1, steps + 1)
    t, dt = t_span[0], t_span[1] - t_span[0]
    p = "decoder.estimator."
    for step in range(1, steps + 1):
        te = om.timestep_embedding(om.sinusoidal_pos_emb(t, 2 * cfg["odim"]), sd, p + "time_mlp.")
        for r, name in enumerate(names):
            want = F.linear(F.mish(te), sd[f"{p}{name}.0.mlp.1.weight"], sd[f"{p}{name}.0.mlp.1.bias"])[0]
            assert float((temb[step - 1, r] - want).abs().max()) < 1e-6
        assert abs(dts[step - 1] - float(dt)) < 1e-9
        t = t + dt
        if step < steps:
            dt = t_span[step + 1] - t


def test_oracle_runs_on_the_recipe_weights():
    cfg = recipes.SMALL_MATCHA
    sd = recipes.make_matcha_state_dict(cfg, 0)
    x = recipes.make_phonemes(6, 1, cfg["idim"])
    frames = int(om.matcha_inference(sd, cfg, x, recipes.make_noise(400, cfg["odim"], 0).t(), 2, 0.667)["feat_gen"].shape[0])
    assert frames > 0 and frames % 2 == 0


def test_packed_dataflow_equals_the_oracle():
    """the whole dataflow of engine_matcha.cu (paired-row down-sampling, 3-tap up-sampling, concatenation operands, stacked
    q / k / v, SnakeBeta constants, time table, Euler update in the projection) evaluated on the CPU FROM THE PACKED TABLE
    (tests/packed_emul.py::emul_matcha) against the oracle: a packing mistake shows up without a GPU"""
    import packed_emul

    cfg = recipes.SMALL_MATCHA
    sd = recipes.make_matcha_state_dict(cfg, 7)
    m = jatts_b200.MatchaTTS(**cfg)
    m.load_state_dict(sd)
    packed = _pack.pack_matcha(sd, m._cfg, 128)
    steps, temp = 3, 0.667
    temb, dts = m._time_table(steps, torch.device("cpu"))
    for n_tok, seed in ((5, 1), (11, 2)):
        x = recipes.make_phonemes(n_tok, 30 + seed, cfg["idim"])
        z = recipes.make_noise(512, cfg["odim"], seed)
        ref = om.matcha_inference(sd, cfg, x, z.t(), steps, temp)
        em = packed_emul.emul_matcha(packed, cfg, x, z * temp, temb, list(dts))
        assert torch.equal(ref["duration"], em["duration"])
        assert em["feat_gen"].shape == ref["feat_gen"].shape
        assert float((em["feat_gen"] - ref["feat_gen"]).abs().max()) < 2e-4
