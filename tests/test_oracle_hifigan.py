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
