"""CPU restatement of the reference FastSpeech2 inference arithmetic.  TEST INFRASTRUCTURE ONLY.

Plain functional fp32 PyTorch over a reference-format ``state_dict``; no ``nn.Module`` from the
reference is used, so this file travels to the GPU box (where /root/reference does not exist).
Every function cites the reference lines it restates.  Pinned against the real reference code by
tests/test_oracle_pin.py and by tests/golden/fs2_*.npz (made with tests/golden/make_golden.py).

Supported configuration = what every shipped recipe uses (SURVEY.md finding 5): Conformer encoder
and decoder with legacy relative-position attention, macaron conv1d FFN, CNN module,
``normalize_before``; ``reduction_factor`` 1; optional x-vector conditioning (``add``).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

LN_EPS = 1e-12  # jatts/modules/transformer/layer_norm.py:23
BN_EPS = 1e-5   # torch.nn.BatchNorm1d default (conformer/convolution.py:45, pre_postnets.py:121)
PE_MAX_LEN = 5000  # jatts/modules/positional_encoding.py:26,212


def legacy_rel_pe_table(d_model: int, length: int, dtype=torch.float32) -> torch.Tensor:
    """positional_encoding.py:36-57 with reverse=True at the *initial* max_len=5000:
    row n holds the sinusoid of position (4999 - n); ``forward`` returns rows [0, T) (:231-235)."""
    assert length <= PE_MAX_LEN, "reference rebuilds the table above 5000 frames; not supported"
    position = torch.arange(PE_MAX_LEN - 1, -1, -1.0, dtype=torch.float32).unsqueeze(1)
    div_term = torch.exp(
        torch.arange(0, d_model, 2, dtype=torch.float32) * -(math.log(10000.0) / d_model)
    )
    pe = torch.zeros(PE_MAX_LEN, d_model)
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe[:length].to(dtype)


def layer_norm(x, w, b):
    """layer_norm.py:12-42 (eps 1e-12) on the last dim."""
    return F.layer_norm(x, (x.shape[-1],), w, b, LN_EPS)


def conv_ffn(x, sd, p):
    """multi_layer_conv.py:52-63: Conv1d(k) -> ReLU -> Conv1d(k); x is (T, D)."""
    k = sd[p + "w_1.weight"].shape[-1]
    h = F.conv1d(x.t().unsqueeze(0), sd[p + "w_1.weight"], sd[p + "w_1.bias"], padding=(k - 1) // 2)
    h = torch.relu(h)
    h = F.conv1d(h, sd[p + "w_2.weight"], sd[p + "w_2.bias"], padding=(k - 1) // 2)
    return h.squeeze(0).t()


def rel_shift_legacy(x):
    """attention.py:142-162 (the known-buggy ESPnet 'legacy' variant), x is (H, T, T)."""
    h, t1, t2 = x.shape
    zero_pad = torch.zeros((h, t1, 1), dtype=x.dtype)
    x_padded = torch.cat([zero_pad, x], dim=-1)
    x_padded = x_padded.view(h, t2 + 1, t1)
    return x_padded[:, 1:].view_as(x)


def rel_shift_closed_form(bd):
    """SURVEY.md 8(a) quirk 3, the closed form the CUDA kernel uses; (H,T,T) -> (H,T,T)."""
    h, t, _ = bd.shape
    out = torch.zeros_like(bd)
    for a in range(t):
        for b in range(t):
            if b <= a:
                out[:, a, b] = bd[:, a, t - 1 - a + b]
            elif b == a + 1:
                out[:, a, b] = 0.0
            else:
                out[:, a, b] = bd[:, a + 1, b - a - 2]
    return out


def rel_mhsa(x, pos_emb, sd, p, n_head):
    """attention.py:164-206 (+ forward_qkv :39-61, forward_attention :63-93, mask=None / all-true)."""
    t, d = x.shape
    dk = d // n_head
    q = F.linear(x, sd[p + "linear_q.weight"], sd[p + "linear_q.bias"]).view(t, n_head, dk)
    k = F.linear(x, sd[p + "linear_k.weight"], sd[p + "linear_k.bias"]).view(t, n_head, dk)
    v = F.linear(x, sd[p + "linear_v.weight"], sd[p + "linear_v.bias"]).view(t, n_head, dk)
    pp = F.linear(pos_emb, sd[p + "linear_pos.weight"]).view(t, n_head, dk)
    q_u = (q + sd[p + "pos_bias_u"]).transpose(0, 1)  # (H, T, dk)
    q_v = (q + sd[p + "pos_bias_v"]).transpose(0, 1)
    k = k.transpose(0, 1)
    v = v.transpose(0, 1)
    pp = pp.transpose(0, 1)
    ac = torch.matmul(q_u, k.transpose(-2, -1))
    bd = rel_shift_legacy(torch.matmul(q_v, pp.transpose(-2, -1)))
    scores = (ac + bd) / math.sqrt(dk)
    attn = torch.softmax(scores, dim=-1)
    ctx = torch.matmul(attn, v).transpose(0, 1).contiguous().view(t, d)
    return F.linear(ctx, sd[p + "linear_out.weight"], sd[p + "linear_out.bias"])


def conv_module(x, sd, p):
    """conformer/convolution.py:56-79: pw1 -> GLU -> depthwise -> BatchNorm(eval) -> Swish -> pw2."""
    d = x.shape[1]
    h = x.t().unsqueeze(0)
    h = F.conv1d(h, sd[p + "pointwise_conv1.weight"], sd[p + "pointwise_conv1.bias"])
    h = F.glu(h, dim=1)
    kk = sd[p + "depthwise_conv.weight"].shape[-1]
    h = F.conv1d(h, sd[p + "depthwise_conv.weight"], sd[p + "depthwise_conv.bias"],
                 padding=(kk - 1) // 2, groups=d)
    h = F.batch_norm(h, sd[p + "norm.running_mean"], sd[p + "norm.running_var"],
                     sd[p + "norm.weight"], sd[p + "norm.bias"], False, 0.0, BN_EPS)
    h = h * torch.sigmoid(h)  # swish.py:13-18
    h = F.conv1d(h, sd[p + "pointwise_conv2.weight"], sd[p + "pointwise_conv2.bias"])
    return h.squeeze(0).t()


def conformer_layer(x, pos_emb, sd, p, n_head):
    """conformer/encoder_layer.py:78-178, eval mode, normalize_before, macaron, cnn module."""
    x = x + 0.5 * conv_ffn(layer_norm(x, sd[p + "norm_ff_macaron.weight"], sd[p + "norm_ff_macaron.bias"]),
                           sd, p + "feed_forward_macaron.")
    x = x + rel_mhsa(layer_norm(x, sd[p + "norm_mha.weight"], sd[p + "norm_mha.bias"]),
                     pos_emb, sd, p + "self_attn.", n_head)
    x = x + conv_module(layer_norm(x, sd[p + "norm_conv.weight"], sd[p + "norm_conv.bias"]),
                        sd, p + "conv_module.")
    x = x + 0.5 * conv_ffn(layer_norm(x, sd[p + "norm_ff.weight"], sd[p + "norm_ff.bias"]),
                           sd, p + "feed_forward.")
    return layer_norm(x, sd[p + "norm_final.weight"], sd[p + "norm_final.bias"])


def conformer_stack(x, sd, prefix, n_layers, n_head):
    """conformer/encoder.py:233-289 after the embedding: x*sqrt(D) and the reversed PE slice
    (positional_encoding.py:221-235; also applied to the decoder input, encoder.py:138-141),
    N layers, after_norm."""
    t, d = x.shape
    pos_emb = legacy_rel_pe_table(d, t, x.dtype)
    x = x * math.sqrt(d)
    for i in range(n_layers):
        x = conformer_layer(x, pos_emb, sd, f"{prefix}.encoders.{i}.", n_head)
    return layer_norm(x, sd[prefix + ".after_norm.weight"], sd[prefix + ".after_norm.bias"])


def predictor_stack(x, sd, prefix, n_layers):
    """duration_predictor.py:78-84 / variance_predictor.py:75-81:
    n x [Conv1d -> ReLU -> LayerNorm(channel dim)] -> Linear(->1); x is (T, D) -> (T,)."""
    h = x.t().unsqueeze(0)
    for i in range(n_layers):
        w = sd[f"{prefix}.conv.{i}.0.weight"]
        h = torch.relu(F.conv1d(h, w, sd[f"{prefix}.conv.{i}.0.bias"], padding=(w.shape[-1] - 1) // 2))
        h = layer_norm(h.transpose(1, 2), sd[f"{prefix}.conv.{i}.2.weight"],
                       sd[f"{prefix}.conv.{i}.2.bias"]).transpose(1, 2)
    h = h.squeeze(0).t()
    return F.linear(h, sd[prefix + ".linear.weight"], sd[prefix + ".linear.bias"]).squeeze(-1)


def duration_from_log(xs, offset: float = 1.0):
    """duration_predictor.py:86-90: clamp(round(exp(x) - offset), min=0).long()
    (torch.round is round-half-to-even)."""
    return torch.clamp(torch.round(xs.exp() - offset), min=0).long()


def length_regulate(hs, ds, alpha: float = 1.0):
    """length_regulator.py:70-97 for one utterance (B=1).  Returns (expanded, durations_used, index).

    ``ds.sum() == 0`` sets EVERY token of the all-zero row to 1 (:86-94) and, when alpha == 1.0, the
    mutation is in place so the duration returned by ``inference`` shows the 1s.
    """
    if alpha != 1.0:
        assert alpha > 0
        ds = torch.round(ds.float() * alpha).long()
    if int(ds.sum()) == 0:
        ds = torch.ones_like(ds)
    idx = torch.repeat_interleave(torch.arange(ds.numel()), ds)
    return hs[idx], ds, idx


def postnet(x, sd, n_layers):
    """pre_postnets.py:108-185: (n-1) x [Conv1d(no bias) -> BN -> tanh] + [Conv1d -> BN]; x is (T, odim)."""
    h = x.t().unsqueeze(0)
    for i in range(n_layers):
        w = sd[f"postnet.postnet.{i}.0.weight"]
        h = F.conv1d(h, w, None, padding=(w.shape[-1] - 1) // 2)
        q = f"postnet.postnet.{i}.1."
        h = F.batch_norm(h, sd[q + "running_mean"], sd[q + "running_var"], sd[q + "weight"],
                         sd[q + "bias"], False, 0.0, BN_EPS)
        if i != n_layers - 1:
            h = torch.tanh(h)
    return h.squeeze(0).t()


@torch.no_grad()
def fs2_log_durations(sd: Dict[str, torch.Tensor], cfg: dict, text: torch.Tensor,
                      spemb: Optional[torch.Tensor] = None, dtype=torch.float32) -> torch.Tensor:
    """The front half of ``_forward`` only (fastspeech2.py:583-612): encoder -> (speaker add) -> duration
    predictor, returning the LOG durations before ``round(exp(x) - 1)``.  With ``dtype=torch.float64`` this
    is the screening oracle of SURVEY.md 8(d) recipe B: a token whose fp64 ``exp(x) - 1`` lies next to a
    rounding boundary may legitimately round either way in fp32 arithmetic."""
    sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
    hs = conformer_stack(sd["encoder.embed.0.weight"][text], sd, "encoder", cfg["elayers"], cfg["aheads"])
    if cfg.get("spk_embed_dim"):
        e = F.normalize(spemb.to(dtype).unsqueeze(0)).squeeze(0)
        hs = hs + F.linear(e, sd["projection.weight"], sd["projection.bias"]).unsqueeze(0)
    return predictor_stack(hs, sd, "duration_predictor", cfg["duration_predictor_layers"])


@torch.no_grad()
def fs2_inference(sd: Dict[str, torch.Tensor], cfg: dict, text: torch.Tensor,
                  spemb: Optional[torch.Tensor] = None, alpha: float = 1.0,
                  return_intermediates: bool = False) -> Dict[str, torch.Tensor]:
    """fastspeech2.py:655-735 ``inference`` -> :566-653 ``_forward(is_inference=True)`` for ONE
    utterance (the reference has no correct batched inference: SURVEY finding 6)."""
    d, h = cfg["adim"], cfg["aheads"]
    x = sd["encoder.embed.0.weight"][text]                                  # fastspeech2.py:270-272
    hs = conformer_stack(x, sd, "encoder", cfg["elayers"], h)               # :583
    if cfg.get("spk_embed_dim"):                                             # :596-597, :737-761
        assert cfg.get("spk_embed_integration_type", "add") == "add"
        e = F.normalize(spemb.unsqueeze(0)).squeeze(0)
        hs = hs + F.linear(e, sd["projection.weight"], sd["projection.bias"]).unsqueeze(0)
    p_outs = predictor_stack(hs, sd, "pitch_predictor", cfg["pitch_predictor_layers"])      # :603-605
    e_outs = predictor_stack(hs, sd, "energy_predictor", cfg["energy_predictor_layers"])    # :606-609
    logd = predictor_stack(hs, sd, "duration_predictor", cfg["duration_predictor_layers"])
    d_outs = duration_from_log(logd)                                         # :612
    # pitch/energy embed are Conv1d(1 -> D, k=1): per-token scale + bias (:614-616)
    assert cfg["pitch_embed_kernel_size"] == 1 and cfg["energy_embed_kernel_size"] == 1
    p_embs = p_outs.unsqueeze(1) * sd["pitch_embed.0.weight"][:, 0, 0] + sd["pitch_embed.0.bias"]
    e_embs = e_outs.unsqueeze(1) * sd["energy_embed.0.weight"][:, 0, 0] + sd["energy_embed.0.bias"]
    hs2 = hs + e_embs + p_embs
    hs_lr, d_used, lr_index = length_regulate(hs2, d_outs, alpha)           # :617
    # alpha != 1: the regulator rounds into a NEW tensor, inference() returns the raw prediction;
    # alpha == 1: the all-zero fix mutates d_outs in place and is visible to the caller (quirk 6)
    d_ret = d_outs if alpha != 1.0 else d_used
    zs = conformer_stack(hs_lr, sd, "decoder", cfg["dlayers"], h)           # :640
    before = F.linear(zs, sd["feat_out.weight"], sd["feat_out.bias"])       # :641-643
    after = before + postnet(before, sd, cfg["postnet_layers"])             # :646-651
    out = dict(feat_gen=after, duration=d_ret, pitch=p_outs.unsqueeze(-1), energy=e_outs.unsqueeze(-1))
    if return_intermediates:
        out.update(enc_out=hs, log_duration=logd, lr_index=lr_index, dec_out=zs, before=before,
                   lr_out=hs_lr)
    return out
