"""Structural checks of the HiFi-GAN restatement (parity unpinned upstream: see oracle/hifigan.py)."""
import torch
import torch.nn.functional as F

from oracle import hifigan, recipes


def test_canonical_v1_parameter_count():
    # arXiv:2010.05646 table 1: generator V1 = 13.92 M parameters
    assert hifigan.count_params(recipes.hifigan_state_shapes(recipes.HIFIGAN_V1_CANONICAL)) == 13_926_017
    assert hifigan.count_params(recipes.hifigan_state_shapes(recipes.HIFIGAN_V1_HOP300)) == 12_979_841


def test_output_length_is_hop_times_frames_and_bounded():
    cfg = recipes.HIFIGAN_TINY
    sd = recipes.make_hifigan_state_dict(cfg, 0)
    y = hifigan.hifigan_forward(sd, cfg, recipes.make_mel(11, 0))
    assert y.shape == (11 * 6, 1) and float(y.abs().max()) <= 1.0


def test_weight_norm_fold():
    g = torch.Generator().manual_seed(0)
    v = torch.randn(6, 4, 3, generator=g)
    gg = torch.rand(6, 1, 1, generator=g) + 0.5
    conv = torch.nn.utils.weight_norm(torch.nn.Conv1d(4, 6, 3))
    conv.weight_v.data.copy_(v)
    conv.weight_g.data.copy_(gg)
    folded = hifigan.fold_weight_norm({"c.weight_g": gg, "c.weight_v": v, "c.bias": conv.bias.data})
    x = torch.randn(1, 4, 9, generator=g)
    assert torch.allclose(F.conv1d(x, folded["c.weight"], folded["c.bias"]), conv(x), atol=1e-6)


def test_polyphase_identity_for_transposed_conv():
    """ConvTranspose1d(k=2s, stride s, padding s//2+s%2, output_padding s%2) is a
    2-tap conv per output phase; output length is exactly s*L."""
    g = torch.Generator().manual_seed(1)
    for s in (2, 3, 4, 5, 8):
        ci, co, L = 6, 4, 9
        w = torch.randn(ci, co, 2 * s, generator=g)
        x = torch.randn(1, ci, L, generator=g)
        p = s // 2 + s % 2
        ref = F.conv_transpose1d(x, w, stride=s, padding=p, output_padding=s % 2)[0].t()
        assert ref.shape[0] == s * L
        out = torch.zeros(s * L, co)
        xs = x[0].t()
        for t in range(s * L):
            j, q = (t + p) // s, (t + p) % s
            if j < L:
                out[t] += xs[j] @ w[:, :, q]
            if 0 <= j - 1 < L:
                out[t] += xs[j - 1] @ w[:, :, q + s]
        assert torch.allclose(out, ref, atol=1e-5)


def test_vocoder_decode_affine():
    cfg = recipes.HIFIGAN_TINY
    sd = recipes.make_hifigan_state_dict(cfg, 0)
    c = recipes.make_mel(7, 3)
    st, tg = recipes.make_stats(0), recipes.make_stats(1)
    y = hifigan.vocoder_decode(sd, cfg, c, st, tg)
    c2 = (c * tg["scale"] + tg["mean"] - st["mean"]) / st["scale"]
    assert torch.allclose(y, hifigan.hifigan_forward(sd, cfg, c2).reshape(-1))


def _to_speecht5(sd, cfg):
    """Rename a parallel_wavegan-style state_dict to ``transformers.SpeechT5HifiGan``'s names."""
    nb, nd = len(cfg["resblock_kernel_sizes"]), len(cfg["resblock_dilations"][0])
    out = {"mean": torch.zeros(cfg["in_channels"]), "scale": torch.ones(cfg["in_channels"]),
           "conv_pre.weight": sd["input_conv.weight"], "conv_pre.bias": sd["input_conv.bias"],
           "conv_post.weight": sd["output_conv.1.weight"], "conv_post.bias": sd["output_conv.1.bias"]}
    for i in range(len(cfg["upsample_scales"])):
        out[f"upsampler.{i}.weight"] = sd[f"upsamples.{i}.1.weight"]
        out[f"upsampler.{i}.bias"] = sd[f"upsamples.{i}.1.bias"]
        for j in range(nb):
            for d in range(nd):
                for c in ("convs1", "convs2"):
                    for p in ("weight", "bias"):
                        out[f"resblocks.{i * nb + j}.{c}.{d}.{p}"] = sd[f"blocks.{i * nb + j}.{c}.{d}.1.{p}"]
    return out


def test_restatement_equals_independent_hifigan_v1_implementation():
    """Anchor for the otherwise unpinned vocoder oracle: ``transformers``'
    ``SpeechT5HifiGan`` is an independent implementation of the published HiFi-GAN V1 generator.  On the
    canonical (8,8,2,2)/(16,16,4,4) config, with the same weights, the restatement must reproduce it.

    What this does NOT cover: parallel_wavegan's rule for ODD upsampling scales (hop 300 uses 5 and 3):
    ``padding = s//2 + s%2, output_padding = s%2`` [upstream, unverified] -- SpeechT5HifiGan pads with
    ``(k - s)//2`` and no output padding, which coincides only for even s."""
    transformers = __import__("pytest").importorskip("transformers")
    cfg = recipes.HIFIGAN_V1_CANONICAL
    sd = recipes.make_hifigan_state_dict(cfg, 0)
    hcfg = transformers.SpeechT5HifiGanConfig(
        model_in_dim=cfg["in_channels"], upsample_initial_channel=cfg["channels"],
        upsample_rates=list(cfg["upsample_scales"]), upsample_kernel_sizes=list(cfg["upsample_kernel_sizes"]),
        resblock_kernel_sizes=list(cfg["resblock_kernel_sizes"]),
        resblock_dilation_sizes=[list(d) for d in cfg["resblock_dilations"]],
        leaky_relu_slope=0.1, normalize_before=False)
    ref = transformers.SpeechT5HifiGan(hcfg).eval()
    mapped = _to_speecht5(sd, cfg)
    assert set(mapped) == set(ref.state_dict()), "state_dict name mapping is incomplete"
    ref.load_state_dict(mapped)
    assert sum(p.numel() for p in ref.parameters()) == 13_926_017
    mel = recipes.make_mel(37, 5)
    with torch.no_grad():
        want = ref(mel)
    got = hifigan.hifigan_forward(sd, cfg, mel).reshape(-1)
    assert want.shape == got.shape == (37 * 256,)
    assert float((want - got).abs().max()) <= 1e-6, float((want - got).abs().max())
