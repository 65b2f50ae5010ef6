"""``MatchaTTS`` -- B200-native drop-in for ``jatts.models.MatchaTTS`` on the inference path (BASELINE config 5).

Mirrors the reference class (jatts/models/matchatts.py:30): same constructor keyword arguments (:42-103), same
``state_dict`` key names / shapes / order (``load_state_dict(torch.load(ckpt)["model"])`` of
jatts/bin/tts_decode.py:139-142 works unchanged), ``.eval()`` / ``.to(device)``, and
``inference(text, ..., spembs=None, ..., n_timesteps=None, temperature=None)`` (:482-558) returning
``dict(feat_gen, duration)``.  ``inference_batch`` is the batched form: row *i* equals the reference's
single-utterance ``inference(x_i)`` given the same noise.

The noise: the reference draws ``z = torch.randn_like(mu)`` inside ``CFM.inference`` (flow_matching.py:64) from the
global generator of mu's device.  Here it is drawn with ``torch.randn`` on the model's device, or passed in
(``noise=`` list of (T_i, odim) tensors; ``plan_batch`` returns the T_i) -- which is what the parity tests do.

All arithmetic runs in the CUDA library (jatts_b200/csrc/engine_matcha.cu) through the C ABI of include/jatts_b200.h,
except the time embedding of the Euler steps (decoder.py:47-62, :107-150, :91): it depends on the step index only, a
(steps x blocks x C) table computed once per step count from the model's own parameters.
Training (``forward``) is out of scope and raises.
"""
from __future__ import annotations

import ctypes as C
import math
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

from . import _lib, _pack
from .fastspeech2 import _register


def _state_shapes(cfg: dict) -> "OrderedDict[str, tuple]":
    """state_dict layout of the reference module at this configuration (same order as the reference registers it)."""
    D, H = cfg["adim"], cfg["aheads"]
    k = cfg["positionwise_conv_kernel_size"]
    s: "OrderedDict[str, tuple]" = OrderedDict()
    s["encoder.embed.0.weight"] = (cfg["idim"], D)
    units, ck = cfg["eunits"], cfg["conformer_enc_kernel_size"]
    for i in range(cfg["elayers"]):
        p = f"encoder.encoders.{i}."
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
    s["encoder.after_norm.weight"] = (D,)
    s["encoder.after_norm.bias"] = (D,)
    if cfg["spk_embed_dim"]:
        s["projection.weight"] = (D, cfg["spk_embed_dim"])
        s["projection.bias"] = (D,)
    od = cfg["odim"]
    s["encoder_proj.weight"] = (od, D)
    s["encoder_proj.bias"] = (od,)
    ch, kk = cfg["duration_predictor_chans"], cfg["duration_predictor_kernel_size"]
    for i in range(cfg["duration_predictor_layers"]):
        s[f"duration_predictor.conv.{i}.0.weight"] = (ch, D if i == 0 else ch, kk)
        s[f"duration_predictor.conv.{i}.0.bias"] = (ch,)
        s[f"duration_predictor.conv.{i}.2.weight"] = (ch,)
        s[f"duration_predictor.conv.{i}.2.bias"] = (ch,)
    s["duration_predictor.linear.weight"] = (1, ch)
    s["duration_predictor.linear.bias"] = (1,)

    # ---- decoder.estimator (decoder.py:243-392)
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
        last = i == len(chans) - 1
        q = f"{e}down_blocks.{i}.2." + ("" if last else "conv.")
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
        last = i == len(up) - 2
        if last:
            s[f"{e}up_blocks.{i}.2.weight"] = (co, co, 3)
        else:
            s[f"{e}up_blocks.{i}.2.conv.weight"] = (co, co, 4)     # ConvTranspose1d: [C_in, C_out, k]
        s[f"{e}up_blocks.{i}.2." + ("" if last else "conv.") + "bias"] = (co,)
    s[e + "final_block.block.0.weight"] = (up[-1], up[-1], 3)
    s[e + "final_block.block.0.bias"] = (up[-1],)
    s[e + "final_block.block.1.weight"] = (up[-1],)
    s[e + "final_block.block.1.bias"] = (up[-1],)
    s[e + "final_proj.weight"] = (od, up[-1], 1)
    s[e + "final_proj.bias"] = (od,)
    return s


class MatchaTTS(torch.nn.Module):
    """See module docstring.  Keyword arguments are those of the reference (matchatts.py:42-103)."""

    def __init__(
        self,
        idim: int, odim: int, adim: int = 384, aheads: int = 4, elayers: int = 6, eunits: int = 1536,
        positionwise_layer_type: str = "conv1d", positionwise_conv_kernel_size: int = 1,
        use_scaled_pos_enc: bool = True, use_batch_norm: bool = True, encoder_normalize_before: bool = True,
        encoder_concat_after: bool = False, reduction_factor: int = 1, encoder_type: str = "transformer",
        transformer_enc_dropout_rate: float = 0.1, transformer_enc_positional_dropout_rate: float = 0.1,
        transformer_enc_attn_dropout_rate: float = 0.1,
        conformer_rel_pos_type: str = "legacy", conformer_pos_enc_layer_type: str = "rel_pos",
        conformer_self_attn_layer_type: str = "rel_selfattn", conformer_activation_type: str = "swish",
        use_macaron_style_in_conformer: bool = True, use_cnn_in_conformer: bool = True, zero_triu: bool = False,
        conformer_enc_kernel_size: int = 7, conformer_dec_kernel_size: int = 31,
        decoder_channels=(256, 256), decoder_dropout: float = 0.05, decoder_attention_head_dim: int = 64,
        decoder_n_blocks: int = 1, decoder_num_mid_blocks: int = 2, decoder_num_heads: int = 2,
        decoder_act_fn: str = "snakebeta",
        duration_predictor_layers: int = 2, duration_predictor_chans: int = 384,
        duration_predictor_kernel_size: int = 3, duration_predictor_dropout_rate: float = 0.1,
        spks: Optional[int] = None, spk_embed_dim: Optional[int] = None, spk_embed_integration_type: str = "add",
        use_gst: bool = False, gst_tokens: int = 10, gst_heads: int = 4, gst_conv_layers: int = 6,
        gst_conv_chans_list: Sequence[int] = (32, 32, 64, 64, 128, 128), gst_conv_kernel_size: int = 3,
        gst_conv_stride: int = 2, gst_gru_layers: int = 1, gst_gru_units: int = 128,
        init_type: str = "xavier_uniform", init_enc_alpha: float = 1.0,
        use_masking: bool = False, use_weighted_masking: bool = False,
        max_len: int = 2048,
    ):
        super().__init__()

        def need(cond, what):
            if not cond:
                raise NotImplementedError(f"jatts_b200.MatchaTTS: {what} is not supported by the B200 path")

        need(encoder_type == "conformer", "encoder_type other than 'conformer'")
        need(conformer_rel_pos_type == "legacy" and conformer_pos_enc_layer_type in ("rel_pos", "legacy_rel_pos")
             and conformer_self_attn_layer_type in ("rel_selfattn", "legacy_rel_selfattn"),
             "non-legacy relative position attention")
        need(positionwise_layer_type == "conv1d" and positionwise_conv_kernel_size % 2 == 1,
             "positionwise_layer_type != 'conv1d' / even kernel size")
        need(use_macaron_style_in_conformer and use_cnn_in_conformer, "conformer without macaron FFN / CNN module")
        need(conformer_activation_type == "swish", "conformer_activation_type != 'swish'")
        need(encoder_normalize_before and not encoder_concat_after, "normalize_before=False / concat_after=True")
        need(reduction_factor == 1, "reduction_factor != 1")
        need(not zero_triu and not use_gst, "zero_triu / GST")
        need(spks is None or spks <= 1, "speaker-id embeddings (spks)")
        need(spk_embed_dim is None or spk_embed_dim <= 0 or spk_embed_integration_type == "add",
             "spk_embed_integration_type != 'add'")
        need(aheads > 0 and adim % aheads == 0 and adim // aheads in (64, 128, 192, 256),
             f"adim/aheads = {adim}/{aheads} (the tcgen05 attention kernel implements head sizes 64 ... 256)")
        need(0 < max_len <= _pack.PE_MAX_LEN, "max_len outside (0, 5000]")
        for what, k in (("positionwise_conv_kernel_size", positionwise_conv_kernel_size),
                        ("duration_predictor_kernel_size", duration_predictor_kernel_size)):
            need(k % 2 == 1 and (k - 1) // 2 <= _pack.GAP_ROWS, f"{what}={k} (must be odd and <= {2 * _pack.GAP_ROWS + 1})")
        need(conformer_enc_kernel_size % 2 == 1, "even conformer kernel size")
        chans = list(decoder_channels)
        need(len(chans) == 2 and chans[0] == chans[1], f"decoder_channels={chans} (two equal widths, as every shipped recipe)")
        need(chans[0] % 64 == 0 and chans[0] <= 512, f"decoder width {chans[0]} (multiple of 64, <= 512)")
        need(decoder_act_fn == "snakebeta", "decoder_act_fn != 'snakebeta'")
        need(decoder_attention_head_dim in (64, 128, 192, 256), f"decoder_attention_head_dim={decoder_attention_head_dim}")
        need(1 <= decoder_n_blocks <= 8 and 1 <= decoder_num_mid_blocks <= 8, "decoder_n_blocks / decoder_num_mid_blocks outside 1..8")
        need(odim % 8 == 0, "odim not a multiple of 8")

        self.idim, self.odim = idim, odim
        self.eos = idim - 1
        self.reduction_factor = reduction_factor
        self.spk_embed_dim = spk_embed_dim if spk_embed_dim and spk_embed_dim > 0 else None
        self.max_len = max_len
        self._cfg = dict(
            idim=idim, odim=odim, adim=adim, aheads=aheads, elayers=elayers, eunits=eunits,
            positionwise_conv_kernel_size=positionwise_conv_kernel_size,
            conformer_enc_kernel_size=conformer_enc_kernel_size,
            duration_predictor_layers=duration_predictor_layers, duration_predictor_chans=duration_predictor_chans,
            duration_predictor_kernel_size=duration_predictor_kernel_size,
            decoder_channels=chans, decoder_attention_head_dim=decoder_attention_head_dim, decoder_n_blocks=decoder_n_blocks,
            decoder_num_mid_blocks=decoder_num_mid_blocks, decoder_num_heads=decoder_num_heads,
            spk_embed_dim=self.spk_embed_dim or 0,
        )
        for name, shape in _state_shapes(self._cfg).items():
            _register(self, name, shape)
        self._engine = None
        self._temb_cache: Dict[int, tuple] = {}

    # ------------------------------------------------------------------ engine lifetime
    def _drop_engine(self):
        if self._engine is not None:
            _lib.lib.jatts_matcha_destroy(self._engine[0])
            self._engine = None
        self._temb_cache = {}

    def __del__(self):
        try:
            self._drop_engine()
        except Exception:
            pass

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        self._drop_engine()
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def _apply(self, fn, *a, **kw):
        self._drop_engine()
        return super()._apply(fn, *a, **kw)

    def _parameters_device(self) -> torch.device:
        return self.encoder_proj.weight.device

    def _get_engine(self):
        dev = self._parameters_device()
        if dev.type != "cuda":
            raise RuntimeError("jatts_b200.MatchaTTS runs on CUDA only (there is no CPU fallback); call .to('cuda')")
        if self._engine is not None and self._engine[2] == dev:
            return self._engine[0]
        self._drop_engine()
        sd = {k: v.detach().float().cpu() for k, v in self.state_dict().items()}
        packed = _pack.pack_matcha(sd, self._cfg, self.max_len)
        with torch.cuda.device(dev):
            table = {k: v.to(dev) for k, v in packed.items()}
            arr, keep = _lib.tensor_table(table)
            c = self._cfg
            ffk, dk, dc = c["positionwise_conv_kernel_size"], c["duration_predictor_kernel_size"], c["duration_predictor_chans"]
            text = _lib.Fs2Config(
                idim=c["idim"], odim=c["odim"], adim=c["adim"], aheads=c["aheads"], elayers=c["elayers"], eunits=c["eunits"],
                dlayers=0, dunits=c["eunits"], ffn_kernel=ffk, enc_cnn_kernel=c["conformer_enc_kernel_size"],
                dec_cnn_kernel=c["conformer_enc_kernel_size"], dur_layers=c["duration_predictor_layers"], dur_chans=dc,
                dur_kernel=dk, pitch_layers=1, pitch_chans=dc, pitch_kernel=dk, energy_layers=1, energy_chans=dc,
                energy_kernel=dk, postnet_layers=1, postnet_chans=64, postnet_filts=dk, spk_embed_dim=c["spk_embed_dim"],
                max_len=self.max_len)
            cfg = _lib.MatchaConfig(text=text, n_channels=2, channels=(C.c_int32 * 4)(*c["decoder_channels"], 0, 0),
                                    n_blocks=c["decoder_n_blocks"], n_mid_blocks=c["decoder_num_mid_blocks"],
                                    n_heads=c["decoder_num_heads"], head_dim=c["decoder_attention_head_dim"])
            handle = C.c_void_p()
            torch.cuda.synchronize(dev)
            _lib.check(_lib.lib.jatts_matcha_create(C.byref(cfg), arr, len(table), C.byref(handle)), "matcha_create")
        self._engine = (handle, (table, keep), dev)
        return handle

    # ------------------------------------------------------------------ time embedding of the Euler steps
    def _time_table(self, n_timesteps: int, dev: torch.device):
        """(temb [steps, blocks, C] on ``dev``, dt float[steps]) for flow_matching.py:70-95's schedule.

        t_span = linspace(0, 1, n + 1); t and dt advance exactly as ``solve_euler`` does (fp32 on the host).  Per step:
        SinusoidalPosEmb(2 * odim)(t) (decoder.py:47-62, scale 1000) -> TimestepEmbedding (linear_1, SiLU, linear_2;
        :107-150) -> per ResnetBlock1D ``mlp = Sequential(Mish, Linear)`` (:85, :91)."""
        if n_timesteps in self._temb_cache:
            return self._temb_cache[n_timesteps]
        e = self.decoder.estimator
        f = lambda t: t.detach().float().cpu()
        dim = 2 * self.odim
        half = dim // 2
        freq = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(10000) / (half - 1)))
        t_span = torch.linspace(0, 1, n_timesteps + 1)
        t, dt = t_span[0], t_span[1] - t_span[0]
        names = _pack.matcha_resnet_names(self._cfg["decoder_num_mid_blocks"])
        rows, dts = [], []
        for step in range(1, n_timesteps + 1):
            emb = 1000.0 * t.reshape(1, 1) * freq.unsqueeze(0)
            emb = torch.cat((emb.sin(), emb.cos()), dim=-1)
            h = F.silu(F.linear(emb, f(e.time_mlp.linear_1.weight), f(e.time_mlp.linear_1.bias)))
            temb = F.mish(F.linear(h, f(e.time_mlp.linear_2.weight), f(e.time_mlp.linear_2.bias)))
            per_block = []
            for name in names:
                blocks, idx = name.split(".")
                mlp = getattr(getattr(getattr(e, blocks), idx), "0").mlp
                per_block.append(F.linear(temb, f(getattr(mlp, "1").weight), f(getattr(mlp, "1").bias))[0])
            rows.append(torch.stack(per_block, 0))
            dts.append(float(dt))
            t = t + dt
            if step < n_timesteps:
                dt = t_span[step + 1] - t
        out = (torch.stack(rows, 0).contiguous().to(dev), (C.c_float * n_timesteps)(*dts))
        self._temb_cache[n_timesteps] = out
        return out

    # ------------------------------------------------------------------ inference
    def forward(self, *args, **kwargs):
        raise NotImplementedError(
            "jatts_b200.MatchaTTS implements the inference path only (matchatts.py:482-558); "
            "training (forward, :317-388) stays with the reference implementation")

    @torch.no_grad()
    def inference_batch(self, texts: Sequence[torch.Tensor], spembs: Optional[torch.Tensor] = None,
                        n_timesteps: int = 10, temperature: float = 0.667, noise=None) -> List[Dict[str, torch.Tensor]]:
        """Batched ``inference``.  ``noise``: optional callable ``frames -> list of (T_i, odim)`` standard-normal tensors, or
        such a list when the frame counts are known (T_i = even-truncated sum of the predicted durations)."""
        handle = self._get_engine()
        dev = self._parameters_device()
        n = len(texts)
        if n == 0:
            return []
        if n_timesteps is None or temperature is None:
            raise ValueError("n_timesteps and temperature must be given (matchatts.py:482-492 passes None into the solver, "
                             "which fails in torch.linspace)")
        if int(n_timesteps) < 1:
            raise ValueError("n_timesteps must be >= 1")
        lens = [int(t.shape[0]) for t in texts]
        if min(lens) <= 0:
            raise ValueError("empty utterance")
        tok = torch.cat([t.reshape(-1) for t in texts]).to(device=dev, dtype=torch.long).contiguous()
        # out-of-range ids raise IndexError as nn.Embedding does -- but the flag is only READ after the plan call below has
        # synchronised the stream anyway (the kernels clamp such ids, nothing is indexed out of the table): reading it here
        # would be a second host<->device round trip per batch, with the GPU idle while the host catches up
        bad_ids = ((tok < 0) | (tok >= self.idim)).any()
        if (self.spk_embed_dim is None) != (spembs is None):
            raise ValueError("spembs must be given iff the model was built with spk_embed_dim")
        sp_ptr = None
        if spembs is not None:
            spembs = spembs.to(device=dev, dtype=torch.float32).reshape(n, self.spk_embed_dim).contiguous()
            sp_ptr = spembs.data_ptr()
        temb, dts = self._time_table(int(n_timesteps), dev)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            h_lens = (C.c_int32 * n)(*lens)
            h_frames = (C.c_int32 * n)()
            rc = _lib.lib.jatts_matcha_plan(handle, tok.data_ptr(), h_lens, n, sp_ptr, h_frames, stream)
            if bool(bad_ids):
                raise IndexError("token id out of range for the embedding table")  # what nn.Embedding raises
            _lib.check(rc, "matcha_plan")
            frames = list(h_frames)
            tot_f, tot_t = sum(frames), sum(lens)
            if callable(noise):
                noise = noise(frames)
            if noise is None:
                z = torch.randn(tot_f, self.odim, device=dev, dtype=torch.float32)
            else:
                if len(noise) != n or any(tuple(zi.shape) != (fi, self.odim) for zi, fi in zip(noise, frames)):
                    raise ValueError(f"noise must be a list of (T_i, odim) tensors with T_i = {frames}")
                z = torch.cat([zi.to(device=dev, dtype=torch.float32) for zi in noise], 0).contiguous() if tot_f else \
                    torch.empty(0, self.odim, device=dev)
            mel = torch.empty(tot_f, self.odim, device=dev, dtype=torch.float32)
            dur = torch.empty(tot_t, device=dev, dtype=torch.long)
            _lib.check(_lib.lib.jatts_matcha_run(handle, z.data_ptr(), float(temperature), temb.data_ptr(), dts,
                                                 int(n_timesteps), mel.data_ptr(), dur.data_ptr(), stream), "matcha_run")
        outs, fo, to = [], 0, 0
        for i in range(n):
            outs.append(dict(feat_gen=mel[fo:fo + frames[i]], duration=dur[to:to + lens[i]]))
            fo += frames[i]
            to += lens[i]
        return outs

    @torch.no_grad()
    def inference(self, text: torch.Tensor, feats: Optional[torch.Tensor] = None,
                  durations: Optional[torch.Tensor] = None, spembs: torch.Tensor = None,
                  sids: Optional[torch.Tensor] = None, lids: Optional[torch.Tensor] = None,
                  n_timesteps: int = None, temperature: float = None,
                  use_teacher_forcing: bool = False) -> Dict[str, torch.Tensor]:
        """Reference signature (matchatts.py:482-493): one utterance in, ``dict(feat_gen, duration)`` out."""
        if use_teacher_forcing or durations is not None:
            raise NotImplementedError("teacher forcing is not on the shipped inference path")
        if sids is not None or lids is not None:
            raise NotImplementedError("sids / lids conditioning is not supported")
        sp = None if spembs is None else spembs.reshape(1, -1)
        return self.inference_batch([text], spembs=sp, n_timesteps=n_timesteps, temperature=temperature)[0]
