"""CPU restatement of the HiFi-GAN V1 generator behind ``parallel_wavegan``.  TEST INFRASTRUCTURE ONLY.

**PARITY UNPINNED.**  The arithmetic lives in the third-party package ``parallel-wavegan``
(kan-bayashi/ParallelWaveGAN), an unpinned dependency of the reference (setup.cfg:17) that is neither
vendored under /root/reference nor installed in this image; the reference has no test or golden
vector for it.  This file restates the published algorithm (Kong et al., arXiv:2010.05646, generator
V1 with "ResBlock1") in the arrangement and state_dict naming of
``parallel_wavegan.models.HiFiGANGenerator`` as reached from the reference's own call sites:

  jatts/vocoder/vocoder.py:41  load_model(checkpoint, config)            -> HiFiGANGenerator(**generator_params)
  jatts/vocoder/vocoder.py:43  model.remove_weight_norm()
  jatts/vocoder/vocoder.py:57-61  c = (c*scale_t + mean_t - mean_v) / scale_v     (restated in ``vocoder_decode``)
  jatts/vocoder/vocoder.py:64  model.inference(c, normalize_before=False).view(-1)

Generator: input_conv Conv1d(80->C,k7,p3); for each upsample stage i: LeakyReLU(0.1) ->
ConvTranspose1d(C/2^i -> C/2^(i+1), k=2s, stride s, padding s//2 + s%2, output_padding s%2) ->
mean over the 3 residual blocks (kernel 3/7/11; each 3 x [LReLU -> Conv1d(dilation d) -> LReLU ->
Conv1d(dilation 1) -> + x]); then LeakyReLU(0.01, torch default) -> Conv1d(C/16 -> 1, k7, p3) -> tanh.
Structural anchor: the canonical (8,8,2,2)/(16,16,4,4) variant counts 13.926 M parameters
(tests/test_oracle_hifigan.py), matching the paper's 13.92 M.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F


def _slope(cfg):
    assert cfg.get("nonlinear_activation", "LeakyReLU") == "LeakyReLU"
    return float(cfg.get("nonlinear_activation_params", {"negative_slope": 0.1})["negative_slope"])


def fold_weight_norm(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """``remove_weight_norm()``: w = g * v / ||v|| with the norm over all dims but 0
    (torch.nn.utils.weight_norm default dim=0, also for ConvTranspose1d)."""
    out = {}
    for k, v in sd.items():
        if k.endswith("weight_g"):
            base = k[: -len("weight_g")]
            vv = sd[base + "weight_v"]
            norm = vv.flatten(1).norm(dim=1).view(-1, *([1] * (vv.dim() - 1)))
            out[base + "weight"] = v * vv / norm
        elif k.endswith("weight_v"):
            continue
        else:
            out[k] = v
    return out


def residual_block(x, sd, prefix, kernel, dilations, slope, additional=True):
    for d, dil in enumerate(dilations):
        xt = F.conv1d(F.leaky_relu(x, slope), sd[f"{prefix}.convs1.{d}.1.weight"],
                      sd.get(f"{prefix}.convs1.{d}.1.bias"), dilation=dil, padding=(kernel - 1) // 2 * dil)
        if additional:
            xt = F.conv1d(F.leaky_relu(xt, slope), sd[f"{prefix}.convs2.{d}.1.weight"],
                          sd.get(f"{prefix}.convs2.{d}.1.bias"), dilation=1, padding=(kernel - 1) // 2)
        x = xt + x
    return x


@torch.no_grad()
def hifigan_forward(sd: Dict[str, torch.Tensor], cfg: dict, c: torch.Tensor,
                    return_stages: bool = False):
    """``HiFiGANGenerator.inference(c, normalize_before=False)``: c (T, 80) -> (T*hop, out_channels)."""
    slope = _slope(cfg)
    k = cfg["kernel_size"]
    x = c.t().unsqueeze(0)
    x = F.conv1d(x, sd["input_conv.weight"], sd["input_conv.bias"], padding=(k - 1) // 2)
    nb = len(cfg["resblock_kernel_sizes"])
    stages = [x]
    for i, (s, uk) in enumerate(zip(cfg["upsample_scales"], cfg["upsample_kernel_sizes"])):
        assert uk == 2 * s
        x = F.conv_transpose1d(F.leaky_relu(x, slope), sd[f"upsamples.{i}.1.weight"],
                               sd[f"upsamples.{i}.1.bias"], stride=s, padding=s // 2 + s % 2,
                               output_padding=s % 2)
        cs = 0.0
        for j in range(nb):
            cs = cs + residual_block(x, sd, f"blocks.{i * nb + j}", cfg["resblock_kernel_sizes"][j],
                                     cfg["resblock_dilations"][j], slope, cfg["use_additional_convs"])
        x = cs / nb
        stages.append(x)
    x = F.conv1d(F.leaky_relu(x, 0.01), sd["output_conv.1.weight"], sd["output_conv.1.bias"],
                 padding=(k - 1) // 2)
    y = torch.tanh(x).squeeze(0).t()
    return (y, stages) if return_stages else y


@torch.no_grad()
def vocoder_decode(sd, cfg, c, stats, trg_stats=None, take_norm_feat=True):
    """jatts/vocoder/vocoder.py:56-67 ``Vocoder.decode``: returns the flat waveform (T*hop,)."""
    if take_norm_feat:
        c = c * trg_stats["scale"] + trg_stats["mean"]
    c = (c - stats["mean"]) / stats["scale"]
    return hifigan_forward(sd, cfg, c).reshape(-1)


def count_params(shapes) -> int:
    n = 0
    for shp in shapes.values():
        m = 1
        for d in shp:
            m *= d
        n += m
    return n


def ac_snr_db(ref: torch.Tensor, test: torch.Tensor) -> float:
    """AC-SNR (signal = reference minus its mean), the waveform parity figure."""
    ref = ref.double().flatten()
    test = test.double().flatten()
    sig = ref - ref.mean()
    err = test - ref
    return float(10.0 * torch.log10(sig.pow(2).sum() / err.pow(2).sum().clamp_min(1e-300)))
