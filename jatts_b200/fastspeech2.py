"""``FastSpeech2`` -- B200-native drop-in for ``jatts.models.FastSpeech2`` on the inference path.

Mirrors the reference class (jatts/models/fastspeech2.py:30): same constructor keyword arguments
(:46-128), same ``state_dict`` key names / shapes / order (so ``load_state_dict(torch.load(ckpt)["model"])``
from jatts/bin/tts_decode.py:139-142 works unchanged), ``.eval()`` / ``.to(device)``, and
``inference(text, ..., spembs=None, ..., alpha=1.0)`` (:655-735) returning
``dict(feat_gen, duration, pitch, energy)``.  ``inference_batch`` is the batched form of the same
call: row *i* equals the reference's single-utterance ``inference(x_i)`` (the reference has no correct
batched inference, SURVEY.md finding 6).

All arithmetic runs in the CUDA library (jatts_b200/csrc) through the C ABI of include/jatts_b200.h.
Training (``forward``) is out of scope and raises.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib, _pack

_BUFFER_LEAVES = ("running_mean", "running_var", "num_batches_tracked")


def _state_shapes(cfg: dict) -> "OrderedDict[str, tuple]":
    """state_dict layout of the reference module at this configuration."""
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
    if cfg["spk_embed_dim"]:
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
    s["feat_out.weight"] = (cfg["odim"], D)
    s["feat_out.bias"] = (cfg["odim"],)
    nl, ch, od = cfg["postnet_layers"], cfg["postnet_chans"], cfg["odim"]
    for i in range(nl):
        ci = od if i == 0 else ch
        co = od if i == nl - 1 else ch
        s[f"postnet.postnet.{i}.0.weight"] = (co, ci, cfg["postnet_filts"])
        for n in ("weight", "bias", "running_mean", "running_var"):
            s[f"postnet.postnet.{i}.1.{n}"] = (co,)
        s[f"postnet.postnet.{i}.1.num_batches_tracked"] = ()
    return s


def _register(root: torch.nn.Module, name: str, shape: tuple) -> None:
    *path, leaf = name.split(".")
    mod = root
    for part in path:
        if part not in mod._modules:
            mod.add_module(part, torch.nn.Module())
        mod = mod._modules[part]
    if leaf == "num_batches_tracked":
        mod.register_buffer(leaf, torch.zeros((), dtype=torch.long))
    elif leaf in _BUFFER_LEAVES:
        mod.register_buffer(leaf, torch.ones(shape) if leaf == "running_var" else torch.zeros(shape))
    else:
        mod.register_parameter(leaf, torch.nn.Parameter(torch.zeros(shape), requires_grad=False))


class FastSpeech2(torch.nn.Module):
    """See module docstring.  Keyword arguments are those of the reference (fastspeech2.py:46-128)."""

    def __init__(
        self,
        idim: int, odim: int, adim: int = 384, aheads: int = 4, elayers: int = 6, eunits: int = 1536,
        dlayers: int = 6, dunits: int = 1536, postnet_layers: int = 5, postnet_chans: int = 512,
        postnet_filts: int = 5, postnet_dropout_rate: float = 0.5, positionwise_layer_type: str = "conv1d",
        positionwise_conv_kernel_size: int = 1, use_scaled_pos_enc: bool = True, use_batch_norm: bool = True,
        encoder_normalize_before: bool = True, decoder_normalize_before: bool = True,
        encoder_concat_after: bool = False, decoder_concat_after: bool = False, reduction_factor: int = 1,
        encoder_type: str = "transformer", decoder_type: str = "transformer",
        transformer_enc_dropout_rate: float = 0.1, transformer_enc_positional_dropout_rate: float = 0.1,
        transformer_enc_attn_dropout_rate: float = 0.1, transformer_dec_dropout_rate: float = 0.1,
        transformer_dec_positional_dropout_rate: float = 0.1, transformer_dec_attn_dropout_rate: float = 0.1,
        conformer_rel_pos_type: str = "legacy", conformer_pos_enc_layer_type: str = "rel_pos",
        conformer_self_attn_layer_type: str = "rel_selfattn", conformer_activation_type: str = "swish",
        use_macaron_style_in_conformer: bool = True, use_cnn_in_conformer: bool = True, zero_triu: bool = False,
        conformer_enc_kernel_size: int = 7, conformer_dec_kernel_size: int = 31,
        duration_predictor_layers: int = 2, duration_predictor_chans: int = 384,
        duration_predictor_kernel_size: int = 3, duration_predictor_dropout_rate: float = 0.1,
        energy_predictor_layers: int = 2, energy_predictor_chans: int = 384, energy_predictor_kernel_size: int = 3,
        energy_predictor_dropout: float = 0.5, energy_embed_kernel_size: int = 9, energy_embed_dropout: float = 0.5,
        stop_gradient_from_energy_predictor: bool = False,
        pitch_predictor_layers: int = 2, pitch_predictor_chans: int = 384, pitch_predictor_kernel_size: int = 3,
        pitch_predictor_dropout: float = 0.5, pitch_embed_kernel_size: int = 9, pitch_embed_dropout: float = 0.5,
        stop_gradient_from_pitch_predictor: bool = False,
        spks: Optional[int] = None, spk_embed_dim: Optional[int] = None, spk_embed_integration_type: str = "add",
        use_gst: bool = False, gst_tokens: int = 10, gst_heads: int = 4, gst_conv_layers: int = 6,
        gst_conv_chans_list: Sequence[int] = (32, 32, 64, 64, 128, 128), gst_conv_kernel_size: int = 3,
        gst_conv_stride: int = 2, gst_gru_layers: int = 1, gst_gru_units: int = 128,
        init_type: str = "xavier_uniform", init_enc_alpha: float = 1.0, init_dec_alpha: float = 1.0,
        use_masking: bool = False, use_weighted_masking: bool = False,
        max_len: int = 2048,
    ):
        super().__init__()
        # ---- what the CUDA path implements = what every shipped recipe uses (SURVEY.md finding 5);
        #      everything else fails loudly instead of silently taking another path
        def need(cond, what):
            if not cond:
                raise NotImplementedError(f"jatts_b200.FastSpeech2: {what} is not supported by the B200 path")

        need(encoder_type == "conformer" and decoder_type == "conformer",
             "encoder_type/decoder_type other than 'conformer' (the reference's 'transformer' branch is dead code)")
        need(conformer_rel_pos_type == "legacy" and conformer_pos_enc_layer_type in ("rel_pos", "legacy_rel_pos")
             and conformer_self_attn_layer_type in ("rel_selfattn", "legacy_rel_selfattn"),
             "non-legacy relative position attention")
        need(positionwise_layer_type == "conv1d", "positionwise_layer_type != 'conv1d'")
        need(positionwise_conv_kernel_size % 2 == 1, "even positionwise_conv_kernel_size")
        need(use_macaron_style_in_conformer and use_cnn_in_conformer, "conformer without macaron FFN / CNN module")
        need(conformer_activation_type == "swish", "conformer_activation_type != 'swish'")
        need(encoder_normalize_before and decoder_normalize_before, "normalize_before=False")
        need(not encoder_concat_after and not decoder_concat_after, "concat_after=True")
        need(reduction_factor == 1, "reduction_factor != 1")
        need(not zero_triu, "zero_triu=True")
        need(use_batch_norm and postnet_layers >= 1, "postnet without batch norm / postnet_layers == 0")
        need(pitch_embed_kernel_size == 1 and energy_embed_kernel_size == 1, "pitch/energy embed kernel size != 1")
        need(not use_gst, "GST")
        need(spks is None or spks <= 1, "speaker-id embeddings (spks)")
        need(spk_embed_dim is None or spk_embed_dim <= 0 or spk_embed_integration_type == "add",
             "spk_embed_integration_type != 'add'")
        need(aheads > 0 and adim % aheads == 0 and adim // aheads in (64, 128, 192, 256),
             f"adim/aheads = {adim}/{aheads} (the tcgen05 attention kernel implements head sizes 64, 128, 192 and 256; there is "
             "no other attention path)")
        need(0 < max_len <= _pack.PE_MAX_LEN, "max_len outside (0, 5000] (positional table is rebuilt above 5000)")
        # the packed layout separates utterances by _pack.GAP_ROWS zero rows: a "same"-padded convolution must be
        # odd (so that it is centred like the reference's padding=(k-1)//2) and its half width must fit the gap,
        # else batch row i would read its neighbour (the conformer depthwise conv is bounded per utterance instead)
        for what, k in (("positionwise_conv_kernel_size", positionwise_conv_kernel_size),
                        ("duration_predictor_kernel_size", duration_predictor_kernel_size),
                        ("pitch_predictor_kernel_size", pitch_predictor_kernel_size),
                        ("energy_predictor_kernel_size", energy_predictor_kernel_size),
                        ("postnet_filts", postnet_filts)):
            need(k % 2 == 1 and (k - 1) // 2 <= _pack.GAP_ROWS, f"{what}={k} (must be odd and <= {2 * _pack.GAP_ROWS + 1})")
        need(conformer_enc_kernel_size % 2 == 1 and conformer_dec_kernel_size % 2 == 1, "even conformer kernel size")

        self.idim, self.odim = idim, odim
        self.eos = idim - 1
        self.reduction_factor = reduction_factor
        self.spk_embed_dim = spk_embed_dim if spk_embed_dim and spk_embed_dim > 0 else None
        self.max_len = max_len
        self._cfg = dict(
            idim=idim, odim=odim, adim=adim, aheads=aheads, elayers=elayers, eunits=eunits, dlayers=dlayers,
            dunits=dunits, positionwise_conv_kernel_size=positionwise_conv_kernel_size,
            conformer_enc_kernel_size=conformer_enc_kernel_size, conformer_dec_kernel_size=conformer_dec_kernel_size,
            duration_predictor_layers=duration_predictor_layers, duration_predictor_chans=duration_predictor_chans,
            duration_predictor_kernel_size=duration_predictor_kernel_size,
            pitch_predictor_layers=pitch_predictor_layers, pitch_predictor_chans=pitch_predictor_chans,
            pitch_predictor_kernel_size=pitch_predictor_kernel_size, pitch_embed_kernel_size=1,
            energy_predictor_layers=energy_predictor_layers, energy_predictor_chans=energy_predictor_chans,
            energy_predictor_kernel_size=energy_predictor_kernel_size, energy_embed_kernel_size=1,
            postnet_layers=postnet_layers, postnet_chans=postnet_chans, postnet_filts=postnet_filts,
            spk_embed_dim=self.spk_embed_dim or 0,
        )
        for name, shape in _state_shapes(self._cfg).items():
            _register(self, name, shape)
        self._engine = None          # (handle, weight keep-alive, device)
        self._engine_version = None  # tuple of parameter versions the engine was built from

    # ------------------------------------------------------------------ engine lifetime
    def _drop_engine(self):
        if self._engine is not None:
            _lib.lib.jatts_fs2_destroy(self._engine[0])
            self._engine = None

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

    def _get_engine(self):
        dev = self._parameters_device()
        if dev.type != "cuda":
            raise RuntimeError("jatts_b200.FastSpeech2 runs on CUDA only (there is no CPU fallback); call .to('cuda')")
        if self._engine is not None and self._engine[2] == dev:
            return self._engine[0]
        self._drop_engine()
        sd = {k: v.detach().float().cpu() for k, v in self.state_dict().items()}
        packed = _pack.pack_fs2(sd, self._cfg, self.max_len)
        with torch.cuda.device(dev):
            table = {k: v.to(dev) for k, v in packed.items()}
            arr, keep = _lib.tensor_table(table)
            c = self._cfg
            cfg = _lib.Fs2Config(
                idim=c["idim"], odim=c["odim"], adim=c["adim"], aheads=c["aheads"], elayers=c["elayers"],
                eunits=c["eunits"], dlayers=c["dlayers"], dunits=c["dunits"],
                ffn_kernel=c["positionwise_conv_kernel_size"], enc_cnn_kernel=c["conformer_enc_kernel_size"],
                dec_cnn_kernel=c["conformer_dec_kernel_size"], dur_layers=c["duration_predictor_layers"],
                dur_chans=c["duration_predictor_chans"], dur_kernel=c["duration_predictor_kernel_size"],
                pitch_layers=c["pitch_predictor_layers"], pitch_chans=c["pitch_predictor_chans"],
                pitch_kernel=c["pitch_predictor_kernel_size"], energy_layers=c["energy_predictor_layers"],
                energy_chans=c["energy_predictor_chans"], energy_kernel=c["energy_predictor_kernel_size"],
                postnet_layers=c["postnet_layers"], postnet_chans=c["postnet_chans"],
                postnet_filts=c["postnet_filts"], spk_embed_dim=c["spk_embed_dim"], max_len=self.max_len)
            handle = C.c_void_p()
            torch.cuda.synchronize(dev)
            _lib.check(_lib.lib.jatts_fs2_create(C.byref(cfg), arr, len(table), C.byref(handle)), "fs2_create")
        self._engine = (handle, (table, keep), dev)
        return handle

    def _parameters_device(self) -> torch.device:
        return self.feat_out.weight.device

    # ------------------------------------------------------------------ inference
    def forward(self, *args, **kwargs):
        raise NotImplementedError(
            "jatts_b200.FastSpeech2 implements the inference path only (fastspeech2.py:655-735); "
            "training (forward, :473-564) stays with the reference implementation")

    @torch.no_grad()
    def inference_batch(self, texts: Sequence[torch.Tensor], spembs: Optional[torch.Tensor] = None,
                        alpha: float = 1.0, return_lr_index: bool = False) -> List[Dict[str, torch.Tensor]]:
        """Batched ``inference``: ``texts`` is a list of LongTensor (T_i,), ``spembs`` (B, spk_embed_dim)."""
        handle = self._get_engine()
        dev = self._parameters_device()
        n = len(texts)
        if n == 0:
            return []
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
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            h_lens = (C.c_int32 * n)(*lens)
            h_frames = (C.c_int32 * n)()
            rc = _lib.lib.jatts_fs2_plan(handle, tok.data_ptr(), h_lens, n, sp_ptr, float(alpha), h_frames, stream)
            if bool(bad_ids):
                raise IndexError("token id out of range for the embedding table")  # what nn.Embedding raises
            _lib.check(rc, "fs2_plan")
            frames = list(h_frames)
            tot_f, tot_t = sum(frames), sum(lens)
            mel = torch.empty(tot_f, self.odim, device=dev, dtype=torch.float32)
            dur = torch.empty(tot_t, device=dev, dtype=torch.long)
            pitch = torch.empty(tot_t, device=dev, dtype=torch.float32)
            energy = torch.empty(tot_t, device=dev, dtype=torch.float32)
            lr = torch.empty(tot_f, device=dev, dtype=torch.int32) if return_lr_index else None
            _lib.check(_lib.lib.jatts_fs2_run(handle, mel.data_ptr(), dur.data_ptr(), pitch.data_ptr(),
                                              energy.data_ptr(), lr.data_ptr() if lr is not None else None, stream),
                       "fs2_run")
        outs = []
        fo = to = 0
        for i in range(n):
            d = dict(feat_gen=mel[fo:fo + frames[i]], duration=dur[to:to + lens[i]],
                     pitch=pitch[to:to + lens[i]].unsqueeze(-1), energy=energy[to:to + lens[i]].unsqueeze(-1))
            if return_lr_index:
                d["lr_index"] = lr[fo:fo + frames[i]]
            outs.append(d)
            fo += frames[i]
            to += lens[i]
        return outs

    @torch.no_grad()
    def inference(self, text: torch.Tensor, feats: Optional[torch.Tensor] = None,
                  durations: Optional[torch.Tensor] = None, spembs: torch.Tensor = None,
                  sids: Optional[torch.Tensor] = None, lids: Optional[torch.Tensor] = None,
                  pitch: Optional[torch.Tensor] = None, energy: Optional[torch.Tensor] = None,
                  alpha: float = 1.0, use_teacher_forcing: bool = False) -> Dict[str, torch.Tensor]:
        """Reference signature (fastspeech2.py:655-667): one utterance in, dict of tensors out."""
        if use_teacher_forcing or durations is not None or pitch is not None or energy is not None:
            raise NotImplementedError("teacher forcing is not on the shipped inference path")
        if sids is not None or lids is not None:
            raise NotImplementedError("sids / lids conditioning is not supported")
        sp = None if spembs is None else spembs.reshape(1, -1)
        return self.inference_batch([text], spembs=sp, alpha=alpha)[0]
