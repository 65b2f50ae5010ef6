"""Seeded synthetic weights / inputs for the oracle, the tests and bench.py.  TEST INFRASTRUCTURE.

Follows SURVEY.md section 8(d).  Everything is generated with explicit ``torch.Generator`` objects
on the CPU so the same tensors appear in this container (where the real reference can be imported
and golden vectors are produced) and on the GPU box (where /root/reference does not exist).

The weight recipe mirrors the statistics of the reference's own initialisation
(jatts/modules/initialize.py:70-96: xavier_uniform on every parameter with dim > 1, zero biases,
default ``reset_parameters`` for Embedding / LayerNorm; BatchNorm running stats 0/1) and, in
``stress`` mode, perturbs every bias / norm affine / BatchNorm statistic so that a kernel which drops
one of those terms fails parity instead of passing by accident.
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict

import torch

# --------------------------------------------------------------------------------------------
# configs (egs/jsut/tts1/conf/fastspeech2.v1.yaml:27-79, egs/jvs/tts1/conf/fastspeech2.v1.yaml:81-82)
# --------------------------------------------------------------------------------------------
JSUT_FS2 = dict(
    idim=45, odim=80, adim=384, aheads=2, elayers=4, eunits=1536, dlayers=4, dunits=1536,
    positionwise_layer_type="conv1d", positionwise_conv_kernel_size=3,
    duration_predictor_layers=2, duration_predictor_chans=256, duration_predictor_kernel_size=3,
    postnet_layers=5, postnet_filts=5, postnet_chans=256, use_masking=True,
    encoder_normalize_before=True, decoder_normalize_before=True, reduction_factor=1,
    encoder_type="conformer", decoder_type="conformer",
    conformer_pos_enc_layer_type="rel_pos", conformer_self_attn_layer_type="rel_selfattn",
    conformer_activation_type="swish", use_macaron_style_in_conformer=True, use_cnn_in_conformer=True,
    conformer_enc_kernel_size=7, conformer_dec_kernel_size=31, init_type="xavier_uniform",
    transformer_enc_dropout_rate=0.2, transformer_enc_positional_dropout_rate=0.2,
    transformer_enc_attn_dropout_rate=0.2, transformer_dec_dropout_rate=0.2,
    transformer_dec_positional_dropout_rate=0.2, transformer_dec_attn_dropout_rate=0.2,
    pitch_predictor_layers=5, pitch_predictor_chans=256, pitch_predictor_kernel_size=5,
    pitch_predictor_dropout=0.5, pitch_embed_kernel_size=1, pitch_embed_dropout=0.0,
    stop_gradient_from_pitch_predictor=True,
    energy_predictor_layers=2, energy_predictor_chans=256, energy_predictor_kernel_size=3,
    energy_predictor_dropout=0.5, energy_embed_kernel_size=1, energy_embed_dropout=0.0,
    stop_gradient_from_energy_predictor=False,
)
JVS_FS2 = dict(JSUT_FS2, spk_embed_dim=192, spk_embed_integration_type="add")

# a shrunken config with the same structure, for CPU-speed tests of host logic
TINY_FS2 = dict(
    JSUT_FS2, idim=20, adim=64, aheads=1, eunits=128, dunits=128, elayers=1, dlayers=1,
    duration_predictor_chans=64, pitch_predictor_chans=64, pitch_predictor_layers=2,
    energy_predictor_chans=64, postnet_chans=64, postnet_layers=3,
)

HIFIGAN_V1_HOP300 = dict(
    in_channels=80, out_channels=1, channels=512, kernel_size=7,
    upsample_scales=(5, 5, 4, 3), upsample_kernel_sizes=(10, 10, 8, 6),
    resblock_kernel_sizes=(3, 7, 11), resblock_dilations=((1, 3, 5), (1, 3, 5), (1, 3, 5)),
    use_additional_convs=True, bias=True,
    nonlinear_activation="LeakyReLU", nonlinear_activation_params={"negative_slope": 0.1},
)
HIFIGAN_V1_CANONICAL = dict(
    HIFIGAN_V1_HOP300, upsample_scales=(8, 8, 2, 2), upsample_kernel_sizes=(16, 16, 4, 4)
)
HIFIGAN_TINY = dict(
    HIFIGAN_V1_HOP300, channels=128, upsample_scales=(3, 2), upsample_kernel_sizes=(6, 4),
    resblock_kernel_sizes=(3, 7), resblock_dilations=((1, 3), (1, 3)),
)
SAMPLING_RATE = 24000
HOP_SIZE = 300


def _gen(name: str, seed: int) -> torch.Generator:
    g = torch.Generator()
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


# --------------------------------------------------------------------------------------------
# FastSpeech2 state_dict layout (checked against the real reference in
# tests/test_oracle_pin.py::test_state_dict_layout_matches_reference)
# --------------------------------------------------------------------------------------------
def fs2_state_shapes(cfg: dict) -> "OrderedDict[str, tuple]":
    D, H = cfg["adim"], cfg["aheads"]
    k = cfg["positionwise_conv_kernel_size"]
    s: "OrderedDict[str, tuple]" = OrderedDict()
    s["encoder.embed.0.weight"] = (cfg["idim"], D)

    def conformer(prefix, nlayers, units, ck):
        for i in range(nlayers):
            p = f"{prefix}.encoders.{i}."
            s[p + "self_attn.pos_bias_u"] = (H, D // H)
            s[p + "self_attn.pos_bias_v"] = (H, D // H)
            for n in ("q", "k", "v", "out"):
                s[p + f"self_attn.linear_{n}.weight"] = (D, D)
                s[p + f"self_attn.linear_{n}.bias"] = (D,)
            s[p + "self_attn.linear_pos.weight"] = (D, D)
            for ff in ("feed_forward", "feed_forward_macaron"):
                s[p + ff + ".w_1.weight"] = (units, D, k)
                s[p + ff + ".w_1.bias"] = (units,)
                s[p + ff + ".w_2.weight"] = (D, units, k)
                s[p + ff + ".w_2.bias"] = (D,)
            s[p + "conv_module.pointwise_conv1.weight"] = (2 * D, D, 1)
            s[p + "conv_module.pointwise_conv1.bias"] = (2 * D,)
            s[p + "conv_module.depthwise_conv.weight"] = (D, 1, ck)
            s[p + "conv_module.depthwise_conv.bias"] = (D,)
            for n in ("weight", "bias", "running_mean", "running_var"):
                s[p + "conv_module.norm." + n] = (D,)
            s[p + "conv_module.norm.num_batches_tracked"] = ()
            s[p + "conv_module.pointwise_conv2.weight"] = (D, D, 1)
            s[p + "conv_module.pointwise_conv2.bias"] = (D,)
            for n in ("norm_ff", "norm_mha", "norm_ff_macaron", "norm_conv", "norm_final"):
                s[p + n + ".weight"] = (D,)
                s[p + n + ".bias"] = (D,)
        s[prefix + ".after_norm.weight"] = (D,)
        s[prefix + ".after_norm.bias"] = (D,)

    conformer("encoder", cfg["elayers"], cfg["eunits"], cfg["conformer_enc_kernel_size"])
    if cfg.get("spk_embed_dim"):
        s["projection.weight"] = (D, cfg["spk_embed_dim"])
        s["projection.bias"] = (D,)

    def predictor(prefix, nl, ch, kk):
        for i in range(nl):
            s[f"{prefix}.conv.{i}.0.weight"] = (ch, D if i == 0 else ch, kk)
            s[f"{prefix}.conv.{i}.0.bias"] = (ch,)
            s[f"{prefix}.conv.{i}.2.weight"] = (ch,)
            s[f"{prefix}.conv.{i}.2.bias"] = (ch,)
        s[f"{prefix}.linear.weight"] = (1, ch)
        s[f"{prefix}.linear.bias"] = (1,)

    predictor("duration_predictor", cfg["duration_predictor_layers"], cfg["duration_predictor_chans"],
              cfg["duration_predictor_kernel_size"])
    predictor("pitch_predictor", cfg["pitch_predictor_layers"], cfg["pitch_predictor_chans"],
              cfg["pitch_predictor_kernel_size"])
    s["pitch_embed.0.weight"] = (D, 1, cfg["pitch_embed_kernel_size"])
    s["pitch_embed.0.bias"] = (D,)
    predictor("energy_predictor", cfg["energy_predictor_layers"], cfg["energy_predictor_chans"],
              cfg["energy_predictor_kernel_size"])
    s["energy_embed.0.weight"] = (D, 1, cfg["energy_embed_kernel_size"])
    s["energy_embed.0.bias"] = (D,)
    conformer("decoder", cfg["dlayers"], cfg["dunits"], cfg["conformer_dec_kernel_size"])
    s["feat_out.weight"] = (cfg["odim"] * cfg["reduction_factor"], D)
    s["feat_out.bias"] = (cfg["odim"] * cfg["reduction_factor"],)
    nl, ch, od = cfg["postnet_layers"], cfg["postnet_chans"], cfg["odim"]
    for i in range(nl):
        ci = od if i == 0 else ch
        co = od if i == nl - 1 else ch
        s[f"postnet.postnet.{i}.0.weight"] = (co, ci, cfg["postnet_filts"])
        for n in ("weight", "bias", "running_mean", "running_var"):
            s[f"postnet.postnet.{i}.1.{n}"] = (co,)
        s[f"postnet.postnet.{i}.1.num_batches_tracked"] = ()
    return s


def _xavier(shape, g):
    rf = 1
    for d in shape[2:]:
        rf *= d
    fan_in, fan_out = shape[1] * rf, shape[0] * rf
    bound = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(shape, generator=g) * 2 - 1) * bound


def make_fs2_state_dict(cfg: dict, seed: int = 0, stress: bool = True,
                        duration_recipe: str = "A") -> "OrderedDict[str, torch.Tensor]":
    """Seeded FastSpeech2 weights.

    duration_recipe "A": ``duration_predictor.linear.weight *= 0.1; bias = log 7`` (durations ~5-8,
    50 phonemes -> ~300 frames: the throughput workload).  "B": weight unscaled, ``bias = log 3``
    (wide spread including zeros: the parity stress case).  "Z": all durations predicted 0 (the
    regulator's all-zero fallback).  SURVEY.md 8(d).
    """
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, shape in fs2_state_shapes(cfg).items():
        g = _gen(name, seed)
        leaf = name.rsplit(".", 1)[-1]
        is_norm = (".norm" in name) or ("after_norm" in name) or name.endswith((".2.weight", ".2.bias")) \
            or (name.startswith("postnet") and ".1." in name)
        if leaf == "num_batches_tracked":
            t = torch.zeros((), dtype=torch.int64)
        elif name == "encoder.embed.0.weight":
            t = torch.randn(shape, generator=g)
            t[0].zero_()  # padding_idx=0 (fastspeech2.py:270-272)
        elif leaf == "running_mean":
            t = 0.1 * torch.randn(shape, generator=g) if stress else torch.zeros(shape)
        elif leaf == "running_var":
            t = 0.5 + torch.rand(shape, generator=g) if stress else torch.ones(shape)
        elif is_norm and leaf == "weight":
            t = 1.0 + 0.1 * (torch.rand(shape, generator=g) * 2 - 1) if stress else torch.ones(shape)
        elif is_norm and leaf == "bias":
            t = 0.05 * (torch.rand(shape, generator=g) * 2 - 1) if stress else torch.zeros(shape)
        elif len(shape) >= 2:
            t = _xavier(shape, g)
        elif leaf == "bias":
            t = 0.05 * (torch.rand(shape, generator=g) * 2 - 1) if stress else torch.zeros(shape)
        else:
            raise AssertionError(name)
        sd[name] = t.float() if t.dtype != torch.int64 else t
    if duration_recipe == "A":
        sd["duration_predictor.linear.weight"] = sd["duration_predictor.linear.weight"] * 0.1
        sd["duration_predictor.linear.bias"] = torch.full((1,), math.log(7.0))
    elif duration_recipe == "B":
        sd["duration_predictor.linear.bias"] = torch.full((1,), math.log(3.0))
    elif duration_recipe == "Z":  # every predicted duration rounds to 0: length_regulator.py:86-94 fallback
        sd["duration_predictor.linear.weight"] = sd["duration_predictor.linear.weight"] * 0.1
        sd["duration_predictor.linear.bias"] = torch.full((1,), -3.0)
    else:
        raise ValueError(duration_recipe)
    return sd


# --------------------------------------------------------------------------------------------
# Matcha-TTS (egs/jsut/tts1/conf/matcha_tts.v1.prior.steplr.large.yaml:22-61; BASELINE config 5)
# --------------------------------------------------------------------------------------------
JSUT_MATCHA = dict(
    idim=45, odim=80, adim=384, aheads=2, elayers=4, eunits=1536,
    positionwise_layer_type="conv1d", positionwise_conv_kernel_size=3,
    duration_predictor_layers=2, duration_predictor_chans=256, duration_predictor_kernel_size=3, use_masking=True,
    encoder_normalize_before=True, reduction_factor=1, encoder_type="conformer",
    conformer_pos_enc_layer_type="rel_pos", conformer_self_attn_layer_type="rel_selfattn",
    conformer_activation_type="swish", use_macaron_style_in_conformer=True, use_cnn_in_conformer=True,
    conformer_enc_kernel_size=7, conformer_dec_kernel_size=31, init_type="xavier_uniform",
    transformer_enc_dropout_rate=0.2, transformer_enc_positional_dropout_rate=0.2, transformer_enc_attn_dropout_rate=0.2,
    decoder_channels=[512, 512], decoder_dropout=0.05, decoder_attention_head_dim=256, decoder_n_blocks=1,
    decoder_num_mid_blocks=2, decoder_num_heads=2, decoder_act_fn="snakebeta",
)
MATCHA_ODE_STEPS, MATCHA_TEMPERATURE = 10, 0.667          # ode_steps / temperature of the same yaml (:78-79)
# the same structure at a size the CPU oracle finishes in seconds (every width a shape the CUDA path implements)
SMALL_MATCHA = dict(
    JSUT_MATCHA, idim=30, odim=16, adim=128, eunits=256, elayers=2, duration_predictor_chans=64,
    decoder_channels=[64, 64], decoder_attention_head_dim=64, decoder_num_heads=2,
)


def matcha_state_shapes(cfg: dict) -> "OrderedDict[str, tuple]":
    """state_dict layout of jatts.models.MatchaTTS at this configuration, in the reference's registration order
    (matchatts.py:196-313, jatts/modules/matchatts/decoder.py:243-392)"""
    text = dict(cfg, dlayers=0, dunits=cfg["eunits"], postnet_layers=0, postnet_chans=0, postnet_filts=1,
                pitch_predictor_layers=0, pitch_predictor_chans=0, pitch_predictor_kernel_size=1, pitch_embed_kernel_size=1,
                energy_predictor_layers=0, energy_predictor_chans=0, energy_predictor_kernel_size=1, energy_embed_kernel_size=1)
    fs2 = fs2_state_shapes(text)
    s: "OrderedDict[str, tuple]" = OrderedDict()
    D, od = cfg["adim"], cfg["odim"]
    for k, v in fs2.items():
        if k.startswith("encoder.") or k.startswith("projection."):
            s[k] = v
    s["encoder_proj.weight"] = (od, D)
    s["encoder_proj.bias"] = (od,)
    for k, v in fs2.items():
        if k.startswith("duration_predictor."):
            s[k] = v
    e = "decoder.estimator."
    chans = list(cfg["decoder_channels"])
    in_ch, ted = 2 * od, chans[0] * 4
    inner = cfg["decoder_num_heads"] * cfg["decoder_attention_head_dim"]
    s[e + "time_mlp.linear_1.weight"] = (ted, in_ch)
    s[e + "time_mlp.linear_1.bias"] = (ted,)
    s[e + "time_mlp.linear_2.weight"] = (ted, ted)
    s[e + "time_mlp.linear_2.bias"] = (ted,)

    def resnet(p, ci, co):
        s[p + "mlp.1.weight"] = (co, ted)
        s[p + "mlp.1.bias"] = (co,)
        for b, c_in in (("block1", ci), ("block2", co)):
            s[p + b + ".block.0.weight"] = (co, c_in, 3)
            s[p + b + ".block.0.bias"] = (co,)
            s[p + b + ".block.1.weight"] = (co,)
            s[p + b + ".block.1.bias"] = (co,)
        s[p + "res_conv.weight"] = (co, ci, 1)
        s[p + "res_conv.bias"] = (co,)

    def transformer(p, c):
        s[p + "norm1.weight"] = (c,)
        s[p + "norm1.bias"] = (c,)
        for n in ("q", "k", "v"):
            s[p + f"attn1.to_{n}.weight"] = (inner, c)
        s[p + "attn1.to_out.0.weight"] = (c, inner)
        s[p + "attn1.to_out.0.bias"] = (c,)
        s[p + "norm3.weight"] = (c,)
        s[p + "norm3.bias"] = (c,)
        s[p + "ff.net.0.alpha"] = (4 * c,)
        s[p + "ff.net.0.beta"] = (4 * c,)
        s[p + "ff.net.0.proj.weight"] = (4 * c, c)
        s[p + "ff.net.0.proj.bias"] = (4 * c,)
        s[p + "ff.net.2.weight"] = (c, 4 * c)
        s[p + "ff.net.2.bias"] = (c,)

    nb = cfg["decoder_n_blocks"]
    co = in_ch
    for i, c in enumerate(chans):
        ci, co = co, c
        resnet(f"{e}down_blocks.{i}.0.", ci, co)
        for j in range(nb):
            transformer(f"{e}down_blocks.{i}.1.{j}.", co)
        q = f"{e}down_blocks.{i}.2." + ("" if i == len(chans) - 1 else "conv.")
        s[q + "weight"] = (co, co, 3)
        s[q + "bias"] = (co,)
    for i in range(cfg["decoder_num_mid_blocks"]):
        resnet(f"{e}mid_blocks.{i}.0.", chans[-1], chans[-1])
        for j in range(nb):
            transformer(f"{e}mid_blocks.{i}.1.{j}.", chans[-1])
    up = chans[::-1] + [chans[0]]
    for i in range(len(up) - 1):
        ci, co = up[i], up[i + 1]
        resnet(f"{e}up_blocks.{i}.0.", 2 * ci, co)
        for j in range(nb):
            transformer(f"{e}up_blocks.{i}.1.{j}.", co)
        if i == len(up) - 2:
            s[f"{e}up_blocks.{i}.2.weight"] = (co, co, 3)
            s[f"{e}up_blocks.{i}.2.bias"] = (co,)
        else:
            s[f"{e}up_blocks.{i}.2.conv.weight"] = (co, co, 4)   # ConvTranspose1d (C_in, C_out, k)
            s[f"{e}up_blocks.{i}.2.conv.bias"] = (co,)
    s[e + "final_block.block.0.weight"] = (up[-1], up[-1], 3)
    s[e + "final_block.block.0.bias"] = (up[-1],)
    s[e + "final_block.block.1.weight"] = (up[-1],)
    s[e + "final_block.block.1.bias"] = (up[-1],)
    s[e + "final_proj.weight"] = (od, up[-1], 1)
    s[e + "final_proj.bias"] = (od,)
    return s


def make_matcha_state_dict(cfg: dict, seed: int = 0, duration_recipe: str = "A") -> "OrderedDict[str, torch.Tensor]":
    """Seeded Matcha-TTS weights: the text side exactly as ``make_fs2_state_dict`` (stress mode), the decoder with
    xavier weights, perturbed biases / norm affines and SnakeBeta log-scale parameters away from their zero init."""
    text = dict(cfg, dlayers=0, dunits=cfg["eunits"], postnet_layers=0, postnet_chans=0, postnet_filts=1,
                pitch_predictor_layers=0, pitch_predictor_chans=0, pitch_predictor_kernel_size=1, pitch_embed_kernel_size=1,
                energy_predictor_layers=0, energy_predictor_chans=0, energy_predictor_kernel_size=1, energy_embed_kernel_size=1)
    fs2 = make_fs2_state_dict(text, seed, True, duration_recipe)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, shape in matcha_state_shapes(cfg).items():
        if name in fs2:
            sd[name] = fs2[name]
            continue
        g = _gen("matcha." + name, seed)
        leaf = name.rsplit(".", 1)[-1]
        is_norm = ".norm1." in name or ".norm3." in name or ".block.1." in name
        if leaf in ("alpha", "beta"):
            t = 0.3 * (torch.rand(shape, generator=g) * 2 - 1)
        elif is_norm and leaf == "weight":
            t = 1.0 + 0.1 * (torch.rand(shape, generator=g) * 2 - 1)
        elif len(shape) >= 2:
            t = _xavier(shape, g)
        elif leaf == "bias":
            t = 0.05 * (torch.rand(shape, generator=g) * 2 - 1)
        else:
            raise AssertionError(name)
        sd[name] = t.float()
    return sd


def make_noise(frames: int, odim: int, seed: int) -> torch.Tensor:
    """standard-normal z (frames, odim) for the flow-matching decoder"""
    return torch.randn(frames, odim, generator=torch.Generator().manual_seed(7000 + seed))


def make_phonemes(t_text: int, seed: int, idim: int = 45) -> torch.Tensor:
    """ids 0 (<blank>/pad) and 1 (<unk>) avoided, idim-1 (<sos/eos>) unused."""
    g = torch.Generator().manual_seed(seed)
    return torch.randint(2, idim - 1, (t_text,), generator=g)


def make_spembs(n: int, seed: int, dim: int = 192) -> torch.Tensor:
    return torch.randn(n, dim, generator=torch.Generator().manual_seed(100 + seed))


# --------------------------------------------------------------------------------------------
# HiFi-GAN weights (weight-norm already removed; parallel_wavegan state_dict names)
# --------------------------------------------------------------------------------------------
def hifigan_state_shapes(cfg: dict) -> "OrderedDict[str, tuple]":
    """parallel_wavegan HiFiGANGenerator key order: input_conv, upsamples.*, blocks.* (convs1 then convs2), output_conv"""
    s: "OrderedDict[str, tuple]" = OrderedDict()
    ch, k = cfg["channels"], cfg["kernel_size"]
    s["input_conv.weight"] = (ch, cfg["in_channels"], k)
    s["input_conv.bias"] = (ch,)
    nb = len(cfg["resblock_kernel_sizes"])
    for i, (sc, uk) in enumerate(zip(cfg["upsample_scales"], cfg["upsample_kernel_sizes"])):
        ci, co = ch // (2 ** i), ch // (2 ** (i + 1))
        s[f"upsamples.{i}.1.weight"] = (ci, co, uk)  # ConvTranspose1d layout (C_in, C_out, k)
        s[f"upsamples.{i}.1.bias"] = (co,)
    for i in range(len(cfg["upsample_scales"])):
        co = ch // (2 ** (i + 1))
        for j, (rk, dils) in enumerate(zip(cfg["resblock_kernel_sizes"], cfg["resblock_dilations"])):
            for cv in ("convs1", "convs2"):
                for d in range(len(dils)):
                    s[f"blocks.{i * nb + j}.{cv}.{d}.1.weight"] = (co, co, rk)
                    s[f"blocks.{i * nb + j}.{cv}.{d}.1.bias"] = (co,)
    cl = ch // (2 ** len(cfg["upsample_scales"]))
    s["output_conv.1.weight"] = (cfg["out_channels"], cl, k)
    s["output_conv.1.bias"] = (cfg["out_channels"],)
    return s


def make_hifigan_state_dict(cfg: dict, seed: int = 0, bias_scale: float = 0.02):
    """weight ~ N(0, 1/sqrt(fan_in)) (fan_in = C_in*k, or C_in*k/stride for ConvTranspose1d).

    SURVEY 8(d): upstream's N(0, 0.01) init collapses activations to the bias; this gain keeps
    activations O(1) through the stack with no tanh saturation.  Biases are small but non-zero so a
    dropped bias is visible.
    """
    sd = OrderedDict()
    scales = dict((f"upsamples.{i}.1.weight", sc) for i, sc in enumerate(cfg["upsample_scales"]))
    for name, shape in hifigan_state_shapes(cfg).items():
        g = _gen("hifigan." + name, seed)
        if name.endswith("weight"):
            if name in scales:
                fan_in = shape[0] * shape[2] / scales[name]
            else:
                fan_in = shape[1] * shape[2]
            sd[name] = torch.randn(shape, generator=g) / math.sqrt(fan_in)
        else:
            sd[name] = bias_scale * torch.randn(shape, generator=g)
    return sd


def make_mel(t: int, seed: int, nmel: int = 80) -> torch.Tensor:
    """'vocoder-normalised' mel clip, N(0,1), temporally smoothed a little so it is not white."""
    g = torch.Generator().manual_seed(7919 * (seed + 1))
    x = torch.randn(t + 4, nmel, generator=g)
    return (0.4 * x[:-4] + 0.3 * x[1:-3] + 0.6 * x[2:-2] + 0.3 * x[3:-1] + 0.4 * x[4:]).contiguous()


def make_stats(seed: int, nmel: int = 80):
    """random per-bin mean/scale pair (exercises SURVEY quirk 9)."""
    g = torch.Generator().manual_seed(31337 + seed)
    return {"mean": torch.randn(nmel, generator=g) * 0.5 - 1.0,
            "scale": 0.5 + torch.rand(nmel, generator=g)}
