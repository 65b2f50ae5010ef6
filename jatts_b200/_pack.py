"""Weight repacking for the CUDA engines (host side, done once at load).

Reference-format tensors -> the layouts the kernels consume: tap-major [taps][N_pad][K_pad] bf16
(fp16 hi / scaled-lo split pairs for the fp32-faithful text2mel GEMMs), BatchNorm folded into the preceding
convolution, GLU halves interleaved per 128-wide tile, ConvTranspose1d split into its two polyphase
taps, relative-position tables pre-multiplied by ``linear_pos``.
"""
from __future__ import annotations

import math

import torch

BN_EPS = 1e-5
PE_MAX_LEN = 5000  # jatts/modules/positional_encoding.py:212 (initial max_len of the legacy table)
GAP_ROWS = 8       # zero rows between utterances of the packed layout (csrc/common.cuh kGapRows)


def round_up(a: int, b: int) -> int:
    return (a + b - 1) // b * b


def pick_block_n(n_cols: int, split: bool) -> int:
    """mirror of csrc/engine_common.cuh::pick_block_n"""
    if split:
        return 128
    for b in (256, 128, 64):
        if n_cols % b == 0:
            return b
    return 32


SPLIT_SCALE = 2048.0  # csrc/common.cuh::kSplitScale


def split16(w: torch.Tensor):
    """fp32 -> fp16 pair (hi, lo) with w ~= hi + lo / 2^11 (22 significand bits); see csrc/common.cuh."""
    w = w.clamp(-65504.0, 65504.0)
    hi = w.to(torch.float16)
    lo = ((w - hi.float()) * SPLIT_SCALE).to(torch.float16)
    return hi, lo


def pack_taps(w_tnk: torch.Tensor, split: bool, out: dict, name: str, bias=None):
    """w_tnk: fp32 [taps, N, K] -> name.hi (/.lo) as [taps, N_pad, K_pad] (bf16, or fp16 pair) (+ name.b)."""
    taps, n, k = w_tnk.shape
    bn = pick_block_n(n, split)
    n_pad, k_pad = round_up(n, bn), round_up(k, 64)
    buf = torch.zeros(taps, n_pad, k_pad, dtype=torch.float32)
    buf[:, :n, :k] = w_tnk
    if split:
        hi, lo = split16(buf)
        out[name + ".hi"], out[name + ".lo"] = hi.contiguous(), lo.contiguous()
    else:
        out[name + ".hi"] = buf.to(torch.bfloat16).contiguous()
    if bias is not None:
        out[name + ".b"] = bias.float().contiguous()


def conv1d_taps(weight: torch.Tensor) -> torch.Tensor:
    """torch Conv1d weight [C_out, C_in, k] (or Linear [C_out, C_in]) -> [k, C_out, C_in]."""
    if weight.dim() == 2:
        weight = weight.unsqueeze(-1)
    return weight.permute(2, 0, 1).contiguous().float()


def glu_interleave(w_nk: torch.Tensor, bias: torch.Tensor, block_n: int = 128):
    """pointwise_conv1 rows [a(0..D) | gate(D..2D)] -> per 128-wide tile [64 a | 64 gate]
    (F.glu(dim=1): first half * sigmoid(second half), convolution.py:71)."""
    d = w_nk.shape[0] // 2
    half = block_n // 2
    assert d % half == 0
    idx = []
    for t in range(d // half):
        idx += list(range(t * half, (t + 1) * half))
        idx += list(range(d + t * half, d + (t + 1) * half))
    idx = torch.tensor(idx)
    return w_nk[idx], bias[idx]


def legacy_pe_table(d_model: int, length: int) -> torch.Tensor:
    """positional_encoding.py:36-57 with reverse=True at max_len 5000: row n = sinusoid(4999 - n)."""
    position = torch.arange(PE_MAX_LEN - 1, -1, -1.0, dtype=torch.float32).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32) * -(math.log(10000.0) / d_model))
    pe = torch.zeros(PE_MAX_LEN, d_model)
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe[:length]


def _ln(out: dict, sd: dict, dst: str, src: str):
    out[dst + ".g"] = sd[src + ".weight"].float().contiguous()
    out[dst + ".b"] = sd[src + ".bias"].float().contiguous()


def _pack_conformer(out: dict, sd: dict, dst: str, src: str, n_layers: int, pe: torch.Tensor):
    """``src``.encoders.* (jatts/modules/conformer/encoder.py) -> ``dst``.<layer>.*; pe: legacy table (max_len, D) fp64"""
    for i in range(n_layers):
        p, q = f"{dst}.{i}.", f"{src}.encoders.{i}."
        for a, b in (("ln_ffm", "norm_ff_macaron"), ("ln_mha", "norm_mha"), ("ln_conv", "norm_conv"),
                     ("ln_ff", "norm_ff"), ("ln_final", "norm_final")):
            _ln(out, sd, p + a, q + b)
        for a, b in (("ffm", "feed_forward_macaron"), ("ff", "feed_forward")):
            pack_taps(conv1d_taps(sd[q + b + ".w_1.weight"]), True, out, p + a + "_w1", sd[q + b + ".w_1.bias"])
            pack_taps(conv1d_taps(sd[q + b + ".w_2.weight"]), True, out, p + a + "_w2", sd[q + b + ".w_2.bias"])
        sa = q + "self_attn."
        # one projection GEMM emits [q + pos_bias_u | q + pos_bias_v | k | v] (attention.py:186-196 adds the two
        # biases to q before the two score products): the second copy of W_q costs one N block of that GEMM
        # and lets the attention kernel take every operand by TMA
        wq, bq = sd[sa + "linear_q.weight"], sd[sa + "linear_q.bias"].float()
        wqkv = torch.cat([wq, wq, sd[sa + "linear_k.weight"], sd[sa + "linear_v.weight"]], 0)
        bqkv = torch.cat([bq + sd[sa + "pos_bias_u"].float().reshape(-1), bq + sd[sa + "pos_bias_v"].float().reshape(-1),
                          sd[sa + "linear_k.bias"].float(), sd[sa + "linear_v.bias"].float()], 0)
        pack_taps(conv1d_taps(wqkv), True, out, p + "qkv", bqkv)
        pack_taps(conv1d_taps(sd[sa + "linear_out.weight"]), True, out, p + "out", sd[sa + "linear_out.bias"])
        # p = linear_pos(pos_emb) is batch independent (attention.py:182-184): precompute per layer
        pos_hi, pos_lo = split16((pe @ sd[sa + "linear_pos.weight"].double().t()).float())
        out[p + "pos.hi"], out[p + "pos.lo"] = pos_hi.contiguous(), pos_lo.contiguous()
        cm = q + "conv_module."
        w1, b1 = glu_interleave(sd[cm + "pointwise_conv1.weight"][:, :, 0].float(), sd[cm + "pointwise_conv1.bias"].float())
        pack_taps(w1.unsqueeze(0), True, out, p + "pw1", b1)
        # depthwise conv + eval BatchNorm folded: y = (conv(x)+b - mean) * g/sqrt(var+eps) + beta
        scale = sd[cm + "norm.weight"].double() / torch.sqrt(sd[cm + "norm.running_var"].double() + BN_EPS)
        wdw = sd[cm + "depthwise_conv.weight"][:, 0, :].double() * scale[:, None]          # [D, k]
        bdw = (sd[cm + "depthwise_conv.bias"].double() - sd[cm + "norm.running_mean"].double()) * scale \
            + sd[cm + "norm.bias"].double()
        out[p + "dw.wT"] = wdw.t().float().contiguous()                                       # [k, D]
        out[p + "dw.b"] = bdw.float().contiguous()
        pack_taps(conv1d_taps(sd[cm + "pointwise_conv2.weight"]), True, out, p + "pw2", sd[cm + "pointwise_conv2.bias"])
    _ln(out, sd, dst + ".after_norm", src + ".after_norm")


def _pack_predictor(out: dict, sd: dict, dst: str, src: str, n_layers: int):
    for i in range(n_layers):
        pack_taps(conv1d_taps(sd[f"{src}.conv.{i}.0.weight"]), True, out, f"{dst}.conv{i}", sd[f"{src}.conv.{i}.0.bias"])
        _ln(out, sd, f"{dst}.ln{i}", f"{src}.conv.{i}.2")
    out[dst + ".lin_w"] = sd[src + ".linear.weight"].float().reshape(-1).contiguous()
    out[dst + ".lin_b"] = sd[src + ".linear.bias"].float().reshape(1).contiguous()


def pack_fs2(sd: dict, cfg: dict, max_len: int) -> dict:
    """reference FastSpeech2 state_dict (CPU fp32) -> engine weight table (CPU tensors)."""
    out: dict = {}
    pe = legacy_pe_table(cfg["adim"], max_len).double()
    _pack_conformer(out, sd, "enc", "encoder", cfg["elayers"], pe)
    _pack_conformer(out, sd, "dec", "decoder", cfg["dlayers"], pe)
    _pack_predictor(out, sd, "dur", "duration_predictor", cfg["duration_predictor_layers"])
    _pack_predictor(out, sd, "pitch", "pitch_predictor", cfg["pitch_predictor_layers"])
    _pack_predictor(out, sd, "energy", "energy_predictor", cfg["energy_predictor_layers"])
    out["emb"] = sd["encoder.embed.0.weight"].float().contiguous()
    for n in ("pitch", "energy"):
        out[n + "_embed.w"] = sd[n + "_embed.0.weight"][:, 0, 0].float().contiguous()
        out[n + "_embed.b"] = sd[n + "_embed.0.bias"].float().contiguous()
    if cfg.get("spk_embed_dim"):
        out["spk.w"] = sd["projection.weight"].float().contiguous()
        out["spk.b"] = sd["projection.bias"].float().contiguous()
    pack_taps(conv1d_taps(sd["feat_out.weight"]), True, out, "feat_out", sd["feat_out.bias"])
    for i in range(cfg["postnet_layers"]):
        q = f"postnet.postnet.{i}."
        w = sd[q + "0.weight"].double()
        scale = sd[q + "1.weight"].double() / torch.sqrt(sd[q + "1.running_var"].double() + BN_EPS)
        wf = (w * scale[:, None, None]).float()
        bf = (sd[q + "1.bias"].double() - sd[q + "1.running_mean"].double() * scale).float()
        pack_taps(conv1d_taps(wf), True, out, f"postnet{i}", bf)
    return out


def matcha_resnet_names(n_mid: int):
    """reference module path of every ResnetBlock1D (+ its transformer list) in the order the engine numbers them"""
    return (["down_blocks.0", "down_blocks.1"] + [f"mid_blocks.{i}" for i in range(n_mid)] + ["up_blocks.0", "up_blocks.1"])


def pack_matcha(sd: dict, cfg: dict, max_len: int) -> dict:
    """reference MatchaTTS state_dict (CPU fp32) -> engine weight table (jatts_b200/csrc/engine_matcha.cu).

    Text side as FastSpeech2.  Decoder (jatts/modules/matchatts/decoder.py): every Conv1d / Linear as tap-major split
    pairs; q / k / v (no bias) stacked into one projection; the stride-2 ``Downsample1D`` convolution as the 2-tap
    convolution over PAIRS of frames it is on the packed layout, ``Upsample1D``'s ConvTranspose1d(4, 2, 1) as a 3-tap
    convolution whose output row is a pair of frames; SnakeBeta's exp(alpha), 1 / (exp(beta) + 1e-9) per channel.  The
    time-embedding weights stay with the Python class (jatts_b200/matchatts.py builds the per-step table)."""
    out: dict = {}
    pe = legacy_pe_table(cfg["adim"], max_len).double()
    _pack_conformer(out, sd, "enc", "encoder", cfg["elayers"], pe)
    _pack_predictor(out, sd, "dur", "duration_predictor", cfg["duration_predictor_layers"])
    out["emb"] = sd["encoder.embed.0.weight"].float().contiguous()
    if cfg.get("spk_embed_dim"):
        out["spk.w"] = sd["projection.weight"].float().contiguous()
        out["spk.b"] = sd["projection.bias"].float().contiguous()
    pack_taps(conv1d_taps(sd["encoder_proj.weight"]), True, out, "enc_proj", sd["encoder_proj.bias"])
    e = "decoder.estimator."
    c = cfg["decoder_channels"][0]
    for r, name in enumerate(matcha_resnet_names(cfg["decoder_num_mid_blocks"])):
        q, p = f"{e}{name}.0.", f"dec.res{r}"
        pack_taps(conv1d_taps(sd[q + "block1.block.0.weight"]), True, out, p + ".c1", sd[q + "block1.block.0.bias"])
        pack_taps(conv1d_taps(sd[q + "block2.block.0.weight"]), True, out, p + ".c2", sd[q + "block2.block.0.bias"])
        pack_taps(conv1d_taps(sd[q + "res_conv.weight"]), True, out, p + ".res", sd[q + "res_conv.bias"])
        _ln(out, sd, p + ".gn1", q + "block1.block.1")
        _ln(out, sd, p + ".gn2", q + "block2.block.1")
        for j in range(cfg["decoder_n_blocks"]):
            q, p = f"{e}{name}.1.{j}.", f"dec.tr{r}_{j}"
            _ln(out, sd, p + ".ln1", q + "norm1")
            _ln(out, sd, p + ".ln3", q + "norm3")
            wqkv = torch.cat([sd[q + "attn1.to_q.weight"], sd[q + "attn1.to_k.weight"], sd[q + "attn1.to_v.weight"]], 0)
            pack_taps(conv1d_taps(wqkv), True, out, p + ".qkv")
            pack_taps(conv1d_taps(sd[q + "attn1.to_out.0.weight"]), True, out, p + ".out", sd[q + "attn1.to_out.0.bias"])
            pack_taps(conv1d_taps(sd[q + "ff.net.0.proj.weight"]), True, out, p + ".ff1", sd[q + "ff.net.0.proj.bias"])
            pack_taps(conv1d_taps(sd[q + "ff.net.2.weight"]), True, out, p + ".ff2", sd[q + "ff.net.2.bias"])
            out[p + ".snake.a"] = torch.exp(sd[q + "ff.net.0.alpha"].double()).float().contiguous()
            out[p + ".snake.ib"] = (1.0 / (torch.exp(sd[q + "ff.net.0.beta"].double()) + 1e-9)).float().contiguous()
    # Downsample1D Conv1d(C, C, 3, stride 2, padding 1): out[j] = W0 x[2j-1] + W1 x[2j] + W2 x[2j+1]; with two frames per
    # row [x[2j] | x[2j+1]] (K = 2C): tap(-1) = [0 | W0], tap(0) = [W1 | W2]
    w = sd[e + "down_blocks.0.2.conv.weight"].float()                      # [C_out, C_in, 3]
    z = torch.zeros_like(w[:, :, 0])
    pack_taps(torch.stack([torch.cat([z, w[:, :, 0]], 1), torch.cat([w[:, :, 1], w[:, :, 2]], 1)], 0), True, out, "dec.down0",
              sd[e + "down_blocks.0.2.conv.bias"])
    pack_taps(conv1d_taps(sd[e + "down_blocks.1.2.weight"]), True, out, "dec.down1", sd[e + "down_blocks.1.2.bias"])
    # Upsample1D ConvTranspose1d(C, C, 4, 2, 1), weight [C_in, C_out, 4]: out[2j] = W1^T x[j] + W3^T x[j-1],
    # out[2j+1] = W2^T x[j] + W0^T x[j+1]; output row j = [out[2j] | out[2j+1]] (N = 2C), taps (-1, 0, +1)
    w = sd[e + "up_blocks.0.2.conv.weight"].float()
    wt = [w[:, :, k].t() for k in range(4)]                                # [C_out, C_in]
    z = torch.zeros_like(wt[0])
    pack_taps(torch.stack([torch.cat([wt[3], z], 0), torch.cat([wt[1], wt[2]], 0), torch.cat([z, wt[0]], 0)], 0), True, out,
              "dec.up0", torch.cat([sd[e + "up_blocks.0.2.conv.bias"]] * 2))
    pack_taps(conv1d_taps(sd[e + "up_blocks.1.2.weight"]), True, out, "dec.up1", sd[e + "up_blocks.1.2.bias"])
    pack_taps(conv1d_taps(sd[e + "final_block.block.0.weight"]), True, out, "dec.final", sd[e + "final_block.block.0.bias"])
    _ln(out, sd, "dec.final.gn", e + "final_block.block.1")
    pack_taps(conv1d_taps(sd[e + "final_proj.weight"]), True, out, "dec.proj", sd[e + "final_proj.bias"])
    assert c == cfg["decoder_channels"][1]
    return out


def pack_hifigan(sd: dict, cfg: dict, mel_scale: torch.Tensor, mel_shift: torch.Tensor) -> dict:
    """parallel_wavegan HiFiGANGenerator state_dict (weight norm already folded) -> engine table."""
    out: dict = {}
    pack_taps(conv1d_taps(sd["input_conv.weight"]), False, out, "input_conv", sd["input_conv.bias"])
    nb = len(cfg["resblock_kernel_sizes"])
    for i, s in enumerate(cfg["upsample_scales"]):
        w = sd[f"upsamples.{i}.1.weight"].float()  # ConvTranspose1d: [C_in, C_out, 2s]
        ci, co, k = w.shape
        assert k == 2 * s
        # out[t] = x[j] w[:,:,q] + x[j-1] w[:,:,q+s],  j=(t+p)//s, q=(t+p)%s  
        tap0 = w[:, :, :s].permute(2, 1, 0).reshape(s * co, ci)
        tap1 = w[:, :, s:].permute(2, 1, 0).reshape(s * co, ci)
        pack_taps(torch.stack([tap0, tap1], 0), False, out, f"ups{i}", sd[f"upsamples.{i}.1.bias"])
        if s * co <= 128:
            # narrow last stage(s): the WHOLE transposed convolution as one 3-tap convolution with N = s * C_out, output
            # row i = frames i*s .. i*s + s - 1 side by side (contiguous in the channels-last output).  Output frame i*s + r
            # has kernel index k = r + p (mod s): r + p < s -> x[i] w_k + x[i-1] w_{k+s}; else x[i+1] w_{k-s}... i.e.
            # with q = (r + p) % s: class A (r < s - p) taps (-1, 0) = (w_{q+s}, w_q), class B taps (0, +1) = (w_{q+s}, w_q).
            # One N = 128 tile per input tile instead of s tiles of N = C_out: the tiles of this stage are too small to pay
            # for their own pipeline hand-offs (measured 248 us for three phases of 1.96 M rows x 64 -> 32 channels)
            pp = s // 2 + s % 2
            t3 = torch.zeros(3, 128, ci)      # N padded to ONE 128-wide tile (pick_block_n would cut 96 into 3 x 32)
            for r in range(s):
                q = (r + pp) % s
                w_lo, w_hi = w[:, :, q].t(), w[:, :, q + s].t()          # [C_out, C_in]
                blk = slice(r * co, (r + 1) * co)
                if r < s - pp:
                    t3[0, blk], t3[1, blk] = w_hi, w_lo
                else:
                    t3[1, blk], t3[2, blk] = w_hi, w_lo
            bias = torch.zeros(128)
            bias[:s * co] = sd[f"upsamples.{i}.1.bias"].float().repeat(s)
            pack_taps(t3, False, out, f"ups{i}.t3", bias)
        for j in range(nb):
            for d in range(len(cfg["resblock_dilations"][j])):
                b = f"blocks.{i * nb + j}."
                pack_taps(conv1d_taps(sd[b + f"convs1.{d}.1.weight"]), False, out, f"rb{i}_{j}.c1_{d}", sd[b + f"convs1.{d}.1.bias"])
                pack_taps(conv1d_taps(sd[b + f"convs2.{d}.1.weight"]), False, out, f"rb{i}_{j}.c2_{d}", sd[b + f"convs2.{d}.1.bias"])
    out["output.w"] = sd["output_conv.1.weight"][0].t().float().contiguous()  # [k, C]
    out["output.b"] = sd["output_conv.1.bias"].float().reshape(1).contiguous()
    out["mel_scale"] = mel_scale.float().contiguous()
    out["mel_shift"] = mel_shift.float().contiguous()
    return out
