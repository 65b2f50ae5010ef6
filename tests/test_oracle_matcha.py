"""Pins oracle/matcha.py (BASELINE config 5) to the REAL reference Matcha-TTS code imported from
/root/reference: MatchaTTS._forward(is_inference), CFM.solve_euler, the U-Net Decoder, ResnetBlock1D / Block1D,
SnakeBeta feed-forward and BasicTransformerBlock.  ``diffusers`` is not installed: its ``Attention`` class is the one
piece restated on both sides (oracle/ref_loader.py::_install_matcha_standins), marked [diffusers, unpinned]."""
import math

import pytest
import torch

from oracle import matcha as om
from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not present (GPU box)")

TINY = dict(
    idim=20, odim=16, adim=64, aheads=1, elayers=1, eunits=128, positionwise_layer_type="conv1d",
    positionwise_conv_kernel_size=3, duration_predictor_layers=2, duration_predictor_chans=32,
    duration_predictor_kernel_size=3, use_masking=True, encoder_normalize_before=True, reduction_factor=1,
    encoder_type="conformer", conformer_pos_enc_layer_type="rel_pos", conformer_self_attn_layer_type="rel_selfattn",
    conformer_activation_type="swish", use_macaron_style_in_conformer=True, use_cnn_in_conformer=True,
    conformer_enc_kernel_size=7, init_type="xavier_uniform",
    decoder_channels=[32, 64], decoder_dropout=0.05, decoder_attention_head_dim=16, decoder_n_blocks=1,
    decoder_num_mid_blocks=2, decoder_num_heads=2, decoder_act_fn="snakebeta",
)


def build(seed):
    cls = ref_loader.load_reference_matcha()
    torch.manual_seed(seed)
    model = cls(**TINY).eval()
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():   # biases / norm affines / snake parameters away from their trivial init values
        for n, p in model.named_parameters():
            if n.endswith("bias") or n.endswith(".alpha") or n.endswith(".beta") or "norm" in n:
                p.add_(torch.randn(p.shape, generator=g) * 0.2)
        model.duration_predictor.linear.bias.fill_(math.log(4.0))
    return model


@pytest.mark.parametrize("seed,n_tok,steps", [(0, 7, 3), (3, 12, 10), (5, 2, 2), (8, 5, 4), (9, 9, 2)])
def test_matcha_inference_matches_the_reference(seed, n_tok, steps):
    model = build(seed)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    text = torch.randint(2, TINY["idim"] - 1, (n_tok,), generator=torch.Generator().manual_seed(seed + 7))
    torch.manual_seed(100 + seed)
    with torch.no_grad():
        ref = model.inference(text, n_timesteps=steps, temperature=0.667)
    t_out = ref["feat_gen"].shape[0]
    torch.manual_seed(100 + seed)
    # what CFM.inference draws: randn_like(mu) of the PERMUTED (B, T, odim) -> (B, odim, T) view keeps its strides, and
    # for a non-contiguous tensor the CPU generator takes its scalar path in memory order -- a different stream from
    # torch.randn(...) of a contiguous tensor.  Reproduce the call itself:
    z = torch.randn_like(torch.empty(1, t_out, TINY["odim"]).permute(0, 2, 1))[0]
    got = om.matcha_inference(sd, TINY, text, z, steps, 0.667)
    assert torch.equal(got["duration"], ref["duration"])
    print("total duration", int(ref["duration"].sum()), "frames out", t_out)
    assert got["feat_gen"].shape == ref["feat_gen"].shape and t_out % 2 == 0 and t_out > 0
    assert float((got["feat_gen"] - ref["feat_gen"]).abs().max()) < 2e-4


def test_decoder_block_by_block():
    """the estimator alone at a few (t, T): catches a compensating pair of mistakes the end-to-end pin could hide"""
    model = build(11)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(2)
    for T, t in ((8, 0.0), (20, 0.35), (34, 1.0)):
        x, mu = torch.randn(1, 16, T, generator=g), torch.randn(1, 16, T, generator=g)
        mask = torch.ones(1, 1, T)
        with torch.no_grad():
            want = model.decoder.estimator(x, mask, mu, torch.tensor(t))
        got = om.decoder_forward(sd, "decoder.estimator.", x, mask, mu, torch.tensor(t), (32, 64), 1, 2, 2)
        assert float((got - want).abs().max()) < 1e-4
