"""CPU restatement of the reference Matcha-TTS inference arithmetic (BASELINE config 5).
TEST INFRASTRUCTURE ONLY -- no CUDA path consumes it yet; it exists so that the decoder kernels of the next round
have an oracle to be checked against from their first line.

Plain functional fp32 PyTorch over a reference-format ``state_dict``:

  jatts/models/matchatts.py:390-480          ``_forward(is_inference=True)``: conformer encoder -> duration predictor
                                             -> LengthRegulator -> ``encoder_proj`` -> even-length truncation -> CFM
  jatts/modules/matchatts/flow_matching.py:48-95   ``CFM.inference`` / ``solve_euler`` (fixed-step Euler)
  jatts/modules/matchatts/decoder.py:47-487        U-Net ``Decoder``: SinusoidalPosEmb, TimestepEmbedding, ResnetBlock1D
                                             (Conv1d k3 -> GroupNorm(8) -> Mish), stride-2 down / ConvTranspose1d(4,2,1) up
  jatts/modules/matchatts/transformer.py:28-364    ``BasicTransformerBlock`` (pre-LN self attention + SnakeBeta feed-forward)

PINNING.  Everything above is checked against the REAL reference modules imported from /root/reference
(tests/test_oracle_matcha.py) -- with ONE exception: ``diffusers.models.attention_processor.Attention`` is an unpinned,
un-vendored third-party class (setup.cfg:40-41; diffusers is not installed here).  It is restated below from its
published definition [diffusers, unpinned]: ``to_q / to_k / to_v`` (no bias), ``heads`` x ``dim_head``, scores scaled by
``dim_head ** -0.5``, softmax, ``to_out.0`` (with bias).  Its ``attention_mask`` handling differs between diffusers
versions, but the reference passes the 0/1 frame mask as an ADDITIVE bias (transformer.py:291-300 -> Attention), and at
inference the reference runs one utterance with an all-ones mask (matchatts.py:446-449), for which every version reduces
to unmasked softmax attention (a constant added to all scores).  The pin test runs the reference modules with this
restated attention standing in for the missing class, so the comparison covers the reference's own code only.

The noise ``z`` that ``CFM.inference`` draws internally (``torch.randn_like``, flow_matching.py:64) is an explicit input.
(Quirk for anyone reproducing a seeded reference run: ``mu`` reaches ``randn_like`` as a permuted (B, T, odim) ->
(B, odim, T) view; ``randn_like`` preserves the strides and the CPU generator fills a non-contiguous tensor through its
scalar path in memory order, which is NOT the stream ``torch.randn`` produces for a contiguous tensor:
``z = torch.randn_like(torch.empty(1, T, odim).permute(0, 2, 1))`` -- pinned by tests/test_oracle_matcha.py.)
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import fs2 as ofs2

GN_GROUPS = 8      # decoder.py:67 Block1D(groups=8)
LN_EPS = 1e-5      # torch.nn.LayerNorm default (transformer.py:216, 250)


def sinusoidal_pos_emb(t: torch.Tensor, dim: int, scale: float = 1000.0) -> torch.Tensor:
    """decoder.py:47-62: t () or (B,) -> (B, dim) = [sin | cos] of scale * t * exp(-i * log(1e4) / (dim/2 - 1))."""
    if t.ndim < 1:
        t = t.unsqueeze(0)
    half = dim // 2
    emb = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(10000) / (half - 1)))
    emb = scale * t.unsqueeze(1) * emb.unsqueeze(0)
    return torch.cat((emb.sin(), emb.cos()), dim=-1)


def timestep_embedding(x, sd, p):
    """decoder.py:107-150 with act_fn='silu', no condition, no post activation."""
    h = F.silu(F.linear(x, sd[p + "linear_1.weight"], sd[p + "linear_1.bias"]))
    return F.linear(h, sd[p + "linear_2.weight"], sd[p + "linear_2.bias"])


def block1d(x, mask, sd, p):
    """decoder.py:65-76: (Conv1d k3 p1 -> GroupNorm(8) -> Mish)(x * mask) * mask; x (B, C, T), mask (B, 1, T)."""
    h = F.conv1d(x * mask, sd[p + "block.0.weight"], sd[p + "block.0.bias"], padding=1)
    h = F.group_norm(h, GN_GROUPS, sd[p + "block.1.weight"], sd[p + "block.1.bias"])
    return F.mish(h) * mask


def resnet_block1d(x, mask, temb, sd, p):
    """decoder.py:79-96."""
    h = block1d(x, mask, sd, p + "block1.")
    h = h + F.linear(F.mish(temb), sd[p + "mlp.1.weight"], sd[p + "mlp.1.bias"]).unsqueeze(-1)
    h = block1d(h, mask, sd, p + "block2.")
    return h + F.conv1d(x * mask, sd[p + "res_conv.weight"], sd[p + "res_conv.bias"])


def attention(x, sd, p, heads: int):
    """diffusers ``Attention`` (self attention, no bias on q/k/v, no mask -- see the module docstring) [diffusers, unpinned];
    x (B, T, C)."""
    b, t, _ = x.shape
    q = F.linear(x, sd[p + "to_q.weight"])
    k = F.linear(x, sd[p + "to_k.weight"])
    v = F.linear(x, sd[p + "to_v.weight"])
    dh = q.shape[-1] // heads
    sp = lambda z: z.view(b, t, heads, dh).transpose(1, 2)
    att = torch.softmax(torch.matmul(sp(q), sp(k).transpose(-2, -1)) * dh ** -0.5, dim=-1)
    o = torch.matmul(att, sp(v)).transpose(1, 2).reshape(b, t, heads * dh)
    return F.linear(o, sd[p + "to_out.0.weight"], sd[p + "to_out.0.bias"])


def snake_beta(x, sd, p):
    """transformer.py:28-102, alpha_logscale=True: y = proj(x); y + sin(y * e^alpha)^2 / (e^beta + 1e-9)."""
    y = F.linear(x, sd[p + "proj.weight"], sd[p + "proj.bias"])
    alpha, beta = torch.exp(sd[p + "alpha"]), torch.exp(sd[p + "beta"])
    return y + (1.0 / (beta + 0.000000001)) * torch.pow(torch.sin(y * alpha), 2)


def basic_transformer_block(x, sd, p, heads: int):
    """transformer.py:160-364 as configured by decoder.py:354-362 (layer_norm, self attention only, snakebeta FF, eval)."""
    h = F.layer_norm(x, (x.shape[-1],), sd[p + "norm1.weight"], sd[p + "norm1.bias"], LN_EPS)
    x = attention(h, sd, p + "attn1.", heads) + x
    h = F.layer_norm(x, (x.shape[-1],), sd[p + "norm3.weight"], sd[p + "norm3.bias"], LN_EPS)
    h = snake_beta(h, sd, p + "ff.net.0.")
    return F.linear(h, sd[p + "ff.net.2.weight"], sd[p + "ff.net.2.bias"]) + x


def decoder_forward(sd, p, x, mask, mu, t, channels, n_blocks: int, num_mid_blocks: int, heads: int):
    """decoder.py:394-487 ``Decoder.forward``: x, mu (B, C_feat, T), mask (B, 1, T), t () or (B,) -> (B, C_feat, T)."""
    in_ch = x.shape[1] + mu.shape[1]
    temb = timestep_embedding(sinusoidal_pos_emb(t, in_ch), sd, p + "time_mlp.")
    x = torch.cat([x, mu], dim=1)
    hiddens, masks = [], [mask]

    def transformers(x, q, m):
        h = x.transpose(1, 2)
        for j in range(n_blocks):
            h = basic_transformer_block(h, sd, f"{q}{j}.", heads)
        return h.transpose(1, 2)

    n = len(channels)
    for i in range(n):
        m = masks[-1]
        x = resnet_block1d(x, m, temb, sd, f"{p}down_blocks.{i}.0.")
        x = transformers(x, f"{p}down_blocks.{i}.1.", m)
        hiddens.append(x)
        if i != n - 1:   # Downsample1D: Conv1d(k3, stride 2, padding 1)
            x = F.conv1d(x * m, sd[f"{p}down_blocks.{i}.2.conv.weight"], sd[f"{p}down_blocks.{i}.2.conv.bias"], stride=2, padding=1)
        else:
            x = F.conv1d(x * m, sd[f"{p}down_blocks.{i}.2.weight"], sd[f"{p}down_blocks.{i}.2.bias"], padding=1)
        masks.append(m[:, :, ::2])
    masks = masks[:-1]
    m_mid = masks[-1]
    for i in range(num_mid_blocks):
        x = resnet_block1d(x, m_mid, temb, sd, f"{p}mid_blocks.{i}.0.")
        x = transformers(x, f"{p}mid_blocks.{i}.1.", m_mid)
    for i in range(n):
        m = masks.pop()
        x = resnet_block1d(torch.cat([x, hiddens.pop()], dim=1), m, temb, sd, f"{p}up_blocks.{i}.0.")
        x = transformers(x, f"{p}up_blocks.{i}.1.", m)
        if i != n - 1:   # Upsample1D: ConvTranspose1d(4, 2, 1)
            x = F.conv_transpose1d(x * m, sd[f"{p}up_blocks.{i}.2.conv.weight"], sd[f"{p}up_blocks.{i}.2.conv.bias"], stride=2, padding=1)
        else:
            x = F.conv1d(x * m, sd[f"{p}up_blocks.{i}.2.weight"], sd[f"{p}up_blocks.{i}.2.bias"], padding=1)
    x = block1d(x, m, sd, p + "final_block.")
    return F.conv1d(x * m, sd[p + "final_proj.weight"], sd[p + "final_proj.bias"]) * mask


@torch.no_grad()
def cfm_solve_euler(sd, p, z, mu, mask, n_timesteps: int, channels, n_blocks, num_mid_blocks, heads):
    """flow_matching.py:48-95: x_0 = z (already scaled by the temperature), t_span = linspace(0, 1, n+1), fixed Euler."""
    t_span = torch.linspace(0, 1, n_timesteps + 1)
    t, dt = t_span[0], t_span[1] - t_span[0]
    x = z
    for step in range(1, len(t_span)):
        x = x + dt * decoder_forward(sd, p + "estimator.", x, mask, mu, t, channels, n_blocks, num_mid_blocks, heads)
        t = t + dt
        if step < len(t_span) - 1:
            dt = t_span[step + 1] - t
    return x


@torch.no_grad()
def matcha_inference(sd: Dict[str, torch.Tensor], cfg: dict, text: torch.Tensor, z: torch.Tensor, n_timesteps: int,
                     temperature: float, spemb: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """matchatts.py:482-560 ``inference`` -> :390-480 ``_forward(is_inference=True)`` for ONE utterance.
    ``z``: standard-normal noise (odim, T_even) as ``torch.randn_like(mu)`` would draw it (before the temperature)."""
    hs = ofs2.conformer_stack(sd["encoder.embed.0.weight"][text], sd, "encoder", cfg["elayers"], cfg["aheads"])
    if cfg.get("spk_embed_dim"):
        e = F.normalize(spemb.unsqueeze(0)).squeeze(0)
        hs = hs + F.linear(e, sd["projection.weight"], sd["projection.bias"]).unsqueeze(0)
    d_outs = ofs2.duration_from_log(ofs2.predictor_stack(hs, sd, "duration_predictor", cfg["duration_predictor_layers"]))
    hs_lr, d_used, _ = ofs2.length_regulate(hs, d_outs, 1.0)
    olen = max(int(d_outs.sum()), 1)                       # clamp_min(d_outs.sum(), 1)  (:446)
    olen -= olen % 2                                       # the decoder halves and doubles the time axis (:454)
    mu = F.linear(hs_lr, sd["encoder_proj.weight"], sd["encoder_proj.bias"])[:olen]      # (:451, :457)
    mu = mu.t().unsqueeze(0)
    mask = torch.ones(1, 1, olen)
    x = cfm_solve_euler(sd, "decoder.", z[:, :olen].unsqueeze(0) * temperature, mu, mask, n_timesteps,
                        tuple(cfg["decoder_channels"]), cfg["decoder_n_blocks"], cfg["decoder_num_mid_blocks"],
                        cfg["decoder_num_heads"])
    return dict(feat_gen=x[0].t(), duration=d_used)
