"""``Vocoder`` / ``HiFiGANGenerator`` -- B200-native drop-ins for the vocoder half of the path.

``Vocoder`` mirrors jatts/vocoder/vocoder.py:16-67 (same constructor arguments, ``decode(c) -> (y, sr)``).
Underneath, ``HiFiGANGenerator`` honours the ``parallel_wavegan`` contract the reference relies on
(vocoder.py:41-44,64): constructor = ``generator_params`` of the vocoder's ``config.yml``,
``load_state_dict`` with parallel_wavegan key names (weight-norm ``weight_g``/``weight_v`` pairs
accepted), ``remove_weight_norm()``, ``.eval()``, ``.to(device)``,
``inference(c, normalize_before=False) -> (T*hop, out_channels)``.  ``decode_batch`` /
``inference_batch`` are the batched forms.  All arithmetic runs in the CUDA library.
"""
from __future__ import annotations

import ctypes as C
import logging
from collections import OrderedDict
from typing import List, Optional, Sequence

import torch

from . import _lib, _pack


def _hifigan_shapes(cfg: dict) -> "OrderedDict[str, tuple]":
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


def _register(root: torch.nn.Module, name: str, shape: tuple) -> None:
    *path, leaf = name.split(".")
    mod = root
    for part in path:
        if part not in mod._modules:
            mod.add_module(part, torch.nn.Module())
        mod = mod._modules[part]
    mod.register_parameter(leaf, torch.nn.Parameter(torch.zeros(shape), requires_grad=False))


class HiFiGANGenerator(torch.nn.Module):
    """parallel_wavegan.models.HiFiGANGenerator-compatible host object (inference only)."""

    def __init__(self, in_channels=80, out_channels=1, channels=512, kernel_size=7,
                 upsample_scales=(8, 8, 2, 2), upsample_kernel_sizes=(16, 16, 4, 4),
                 resblock_kernel_sizes=(3, 7, 11), resblock_dilations=((1, 3, 5), (1, 3, 5), (1, 3, 5)),
                 use_additional_convs=True, bias=True, nonlinear_activation="LeakyReLU",
                 nonlinear_activation_params={"negative_slope": 0.1}, use_causal_conv=False,
                 use_weight_norm=True):
        super().__init__()

        def need(cond, what):
            if not cond:
                raise NotImplementedError(f"jatts_b200.HiFiGANGenerator: {what} is not supported by the B200 path")

        need(out_channels == 1, "out_channels != 1")
        need(use_additional_convs and bias, "use_additional_convs=False / bias=False")
        need(nonlinear_activation == "LeakyReLU", "activations other than LeakyReLU")
        need(not use_causal_conv, "causal convolutions")
        # activations are stored as LeakyReLU(x) and x is recovered as r / slope for r < 0: slope must be in (0, 1]
        need(0.0 < float(nonlinear_activation_params.get("negative_slope", 0.01)) <= 1.0,
             "LeakyReLU negative_slope outside (0, 1]")
        need(kernel_size % 2 == 1 and all(k % 2 == 1 for k in resblock_kernel_sizes), "even kernel sizes")
        need(all(uk == 2 * s for uk, s in zip(upsample_kernel_sizes, upsample_scales)),
             "upsample_kernel_size != 2 * upsample_scale")
        need(len(set(len(d) for d in resblock_dilations)) == 1 and len(resblock_dilations) == len(resblock_kernel_sizes),
             "ragged resblock_dilations")
        need(len(upsample_scales) <= 8 and len(resblock_kernel_sizes) <= 8 and len(resblock_dilations[0]) <= 8,
             "more than 8 stages / blocks / dilations")
        self._cfg = dict(in_channels=in_channels, out_channels=out_channels, channels=channels,
                         kernel_size=kernel_size, upsample_scales=tuple(upsample_scales),
                         upsample_kernel_sizes=tuple(upsample_kernel_sizes),
                         resblock_kernel_sizes=tuple(resblock_kernel_sizes),
                         resblock_dilations=tuple(tuple(d) for d in resblock_dilations),
                         slope=float(nonlinear_activation_params.get("negative_slope", 0.01)))
        self.hop = 1
        for s in upsample_scales:
            self.hop *= int(s)
        for name, shape in _hifigan_shapes(self._cfg).items():
            _register(self, name, shape)
        # parallel_wavegan registers the vocoder's own statistics on the model (register_stats)
        self.register_buffer("mean", torch.zeros(in_channels))
        self.register_buffer("scale", torch.ones(in_channels))
        self._affine = (torch.ones(in_channels), torch.zeros(in_channels))  # applied on load: c*a + b
        self._engine = None

    # ---- parallel_wavegan surface -------------------------------------------------------------
    def remove_weight_norm(self):
        """Weight norm is folded while loading (``load_state_dict``); kept for call compatibility."""
        return None

    def apply_weight_norm(self):
        raise NotImplementedError("training-side API")

    def register_stats(self, stats):
        if isinstance(stats, dict):
            mean, scale = stats["mean"], stats["scale"]
        else:
            raise NotImplementedError("pass a dict with 'mean' and 'scale'")
        self.mean.copy_(torch.as_tensor(mean, dtype=torch.float32).reshape(-1))
        self.scale.copy_(torch.as_tensor(scale, dtype=torch.float32).reshape(-1))

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        sd = OrderedDict()
        for k, v in state_dict.items():
            if k.endswith("weight_g"):
                base = k[: -len("weight_g")]
                vv = state_dict[base + "weight_v"]
                norm = vv.flatten(1).norm(dim=1).view(-1, *([1] * (vv.dim() - 1)))
                sd[base + "weight"] = v * vv / norm
            elif not k.endswith("weight_v"):
                sd[k] = v
        if not strict or ("mean" not in sd and "scale" not in sd):
            sd.setdefault("mean", self.mean)
            sd.setdefault("scale", self.scale)
        self._drop_engine()
        return super().load_state_dict(sd, strict=strict, **kw)

    def set_input_affine(self, a: torch.Tensor, b: torch.Tensor):
        """c_vocoder = c * a + b per mel bin, fused into the operand load (vocoder.py:57-61)."""
        self._affine = (a.detach().float().cpu().reshape(-1), b.detach().float().cpu().reshape(-1))
        self._drop_engine()

    # ---- engine lifetime ------------------------------------------------------------------------
    def _drop_engine(self):
        if self._engine is not None:
            _lib.lib.jatts_hifigan_destroy(self._engine[0])
            self._engine = None

    def __del__(self):
        try:
            self._drop_engine()
        except Exception:
            pass

    def _apply(self, fn, *a, **kw):
        self._drop_engine()
        return super()._apply(fn, *a, **kw)

    def _device(self):
        return self.input_conv.weight.device

    def _get_engine(self, normalize_before: bool):
        dev = self._device()
        if dev.type != "cuda":
            raise RuntimeError("jatts_b200.HiFiGANGenerator runs on CUDA only (there is no CPU fallback)")
        if self._engine is not None and self._engine[2] == dev and self._engine[3] == normalize_before:
            return self._engine[0]
        self._drop_engine()
        sd = {k: v.detach().float().cpu() for k, v in self.state_dict().items()}
        a, b = self._affine
        if normalize_before:  # (c - mean) / scale before the generator
            a, b = a / sd["scale"], (b - sd["mean"]) / sd["scale"]
        c = self._cfg
        packed = _pack.pack_hifigan(sd, c, a, b)
        with torch.cuda.device(dev):
            table = {k: v.to(dev) for k, v in packed.items()}
            arr, keep = _lib.tensor_table(table)
            cfg = _lib.HifiganConfig()
            cfg.in_channels, cfg.out_channels, cfg.channels, cfg.kernel_size = (
                c["in_channels"], c["out_channels"], c["channels"], c["kernel_size"])
            cfg.n_upsamples = len(c["upsample_scales"])
            for i, s in enumerate(c["upsample_scales"]):
                cfg.upsample_scales[i] = s
            cfg.n_resblocks = len(c["resblock_kernel_sizes"])
            cfg.n_dilations = len(c["resblock_dilations"][0])
            for j, k in enumerate(c["resblock_kernel_sizes"]):
                cfg.resblock_kernels[j] = k
                for d, dil in enumerate(c["resblock_dilations"][j]):
                    cfg.resblock_dilations[j][d] = dil
            cfg.lrelu_slope = c["slope"]
            handle = C.c_void_p()
            torch.cuda.synchronize(dev)
            _lib.check(_lib.lib.jatts_hifigan_create(C.byref(cfg), arr, len(table), C.byref(handle)), "hifigan_create")
        self._engine = (handle, (table, keep), dev, normalize_before)
        return handle

    # ---- inference --------------------------------------------------------------------------------
    def forward(self, c):
        raise NotImplementedError("inference only: use inference() / inference_batch()")

    @torch.no_grad()
    def inference_batch(self, mels: Sequence[torch.Tensor], normalize_before: bool = False,
                        pcm16: bool = False) -> List[torch.Tensor]:
        """list of (T_i, in_channels) -> list of (T_i*hop, 1) fp32 waveforms, or int16 PCM samples
        (``lrintf(wave * 32767)``, what ``sf.write(..., "PCM_16")`` stores) when ``pcm16`` is set."""
        if len(mels) == 0:
            return []
        handle = self._get_engine(normalize_before)
        dev = self._device()
        lens = [int(m.shape[0]) for m in mels]
        if min(lens) <= 0:
            raise ValueError("empty mel clip")
        for m in mels:
            if m.dim() != 2 or m.shape[1] != self._cfg["in_channels"]:
                raise ValueError(f"expected (T, {self._cfg['in_channels']}) mel, got {tuple(m.shape)}")
        cat = torch.cat([m.to(device=dev, dtype=torch.float32) for m in mels], 0).contiguous()
        wave = torch.empty(sum(lens) * self.hop, device=dev, dtype=torch.int16 if pcm16 else torch.float32)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            h_lens = (C.c_int32 * len(lens))(*lens)
            run = _lib.lib.jatts_hifigan_run_pcm16 if pcm16 else _lib.lib.jatts_hifigan_run
            _lib.check(run(handle, cat.data_ptr(), h_lens, len(lens), wave.data_ptr(), stream), "hifigan_run")
        outs, o = [], 0
        for n in lens:
            outs.append(wave[o:o + n * self.hop].unsqueeze(-1))
            o += n * self.hop
        return outs

    @torch.no_grad()
    def inference(self, c, normalize_before: bool = False) -> torch.Tensor:
        if not isinstance(c, torch.Tensor):
            c = torch.tensor(c, dtype=torch.float)
        return self.inference_batch([c], normalize_before=normalize_before)[0]


def _read_stats(stats):
    """``stats`` may be a dict (tests), an .npz path, or the recipe's ``stats.h5`` (vocoder.py:47-54 reads ``mean`` and
    ``scale`` with ``read_hdf5``): h5py when it is installed, else the built-in reader of the 'earliest' HDF5 format
    (jatts_b200/_h5lite.py)."""
    if isinstance(stats, dict):
        return stats["mean"], stats["scale"]
    if str(stats).endswith(".npz"):
        import numpy as np

        z = np.load(stats)
        return z["mean"], z["scale"]
    try:
        import h5py
    except ImportError:
        from ._h5lite import read_hdf5

        return read_hdf5(stats, "mean"), read_hdf5(stats, "scale")
    with h5py.File(stats, "r") as f:
        return f["mean"][()], f["scale"][()]


class Vocoder(object):
    """Drop-in for jatts.vocoder.Vocoder (vocoder.py:16-67)."""

    def __init__(self, checkpoint, config, stats, device, trg_stats=None, take_norm_feat=True):
        self.device = device
        if take_norm_feat:
            assert trg_stats is not None, "trg_stats must be given if take_norm_feat=True"
            self.trg_stats = {
                "mean": torch.as_tensor(trg_stats["mean"], dtype=torch.float),
                "scale": torch.as_tensor(trg_stats["scale"], dtype=torch.float),
            }
        self.take_norm_feat = take_norm_feat
        if isinstance(config, dict):
            self.config = config
        else:
            import yaml

            with open(config) as f:
                self.config = yaml.load(f, Loader=yaml.Loader)
        gtype = self.config.get("generator_type", "HiFiGANGenerator")
        if gtype != "HiFiGANGenerator":
            raise NotImplementedError(f"generator_type {gtype}: only HiFiGANGenerator has a B200 path")
        self.model = HiFiGANGenerator(**self.config["generator_params"])
        if isinstance(checkpoint, dict):
            sd = checkpoint
        else:
            sd = torch.load(checkpoint, map_location="cpu")
        if "model" in sd and "generator" in sd["model"]:
            sd = sd["model"]["generator"]
        self.model.load_state_dict(sd)
        logging.info("Loaded vocoder parameters.")
        self.model.remove_weight_norm()
        mean, scale = _read_stats(stats)
        self.stats = {"mean": torch.as_tensor(mean, dtype=torch.float).reshape(-1),
                      "scale": torch.as_tensor(scale, dtype=torch.float).reshape(-1)}
        # c_voc = ((c*scale_t + mean_t) - mean_v) / scale_v  ==  c*a + b   (vocoder.py:57-61)
        if take_norm_feat:
            a = self.trg_stats["scale"].reshape(-1) / self.stats["scale"]
            b = (self.trg_stats["mean"].reshape(-1) - self.stats["mean"]) / self.stats["scale"]
        else:
            a = 1.0 / self.stats["scale"]
            b = -self.stats["mean"] / self.stats["scale"]
        self.model.set_input_affine(a, b)
        self.model = self.model.eval().to(device)

    @torch.no_grad()
    def decode_batch(self, cs: Sequence[torch.Tensor], pcm16: bool = False) -> List[torch.Tensor]:
        """batched ``decode``: list of (T_i, 80) normalised mels -> list of (T_i*hop,) waveforms (fp32, or the
        int16 PCM_16 samples tts_decode.py writes when ``pcm16``)."""
        return [y.view(-1) for y in self.model.inference_batch(list(cs), normalize_before=False, pcm16=pcm16)]

    @torch.no_grad()
    def decode(self, c):
        y = self.decode_batch([c])[0]
        return y, self.config["sampling_rate"]
