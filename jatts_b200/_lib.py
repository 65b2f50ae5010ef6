"""ctypes binding of include/jatts_b200.h (the C-ABI drop-in boundary).

There is NO fallback: if the shared library has not been built (``python -c "import __graft_entry__ as
g; g.build()"`` or ``make -C jatts_b200/csrc``) importing this module raises, and every call that
returns a negative JATTS_E_* code raises ``RuntimeError`` / ``ValueError`` with the library's message.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# JATTS_B200_LIB overrides the location (used to A/B kernel build variants on the GPU box)
LIB_PATH = os.environ.get("JATTS_B200_LIB") or os.path.join(_HERE, "lib", "libjatts_b200.so")

JATTS_F32, JATTS_BF16, JATTS_I64, JATTS_I32, JATTS_F16 = 0, 1, 2, 3, 4
ACT_NONE, ACT_RELU, ACT_LRELU, ACT_TANH, ACT_GLU, ACT_SNAKE = 0, 1, 2, 3, 4, 5


class Tensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("d_ptr", C.c_void_p), ("numel", C.c_int64), ("dtype", C.c_int32)]


class Fs2Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "idim", "odim", "adim", "aheads", "elayers", "eunits", "dlayers", "dunits", "ffn_kernel",
        "enc_cnn_kernel", "dec_cnn_kernel", "dur_layers", "dur_chans", "dur_kernel",
        "pitch_layers", "pitch_chans", "pitch_kernel", "energy_layers", "energy_chans", "energy_kernel",
        "postnet_layers", "postnet_chans", "postnet_filts", "spk_embed_dim", "max_len")]


class MatchaConfig(C.Structure):
    _fields_ = [("text", Fs2Config), ("n_channels", C.c_int32), ("channels", C.c_int32 * 4), ("n_blocks", C.c_int32),
                ("n_mid_blocks", C.c_int32), ("n_heads", C.c_int32), ("head_dim", C.c_int32)]


class HifiganConfig(C.Structure):
    _fields_ = [("in_channels", C.c_int32), ("out_channels", C.c_int32), ("channels", C.c_int32),
                ("kernel_size", C.c_int32), ("n_upsamples", C.c_int32), ("upsample_scales", C.c_int32 * 8),
                ("n_resblocks", C.c_int32), ("resblock_kernels", C.c_int32 * 8), ("n_dilations", C.c_int32),
                ("resblock_dilations", (C.c_int32 * 8) * 8), ("lrelu_slope", C.c_float)]


class ConvGemmArgs(C.Structure):
    _fields_ = [
        ("d_a_hi", C.c_void_p), ("d_a_lo", C.c_void_p), ("a_rows", C.c_int32), ("a_ld", C.c_int32),
        ("a_cols", C.c_int32),
        ("d_w_hi", C.c_void_p), ("d_w_lo", C.c_void_p),
        ("taps", C.c_int32), ("n_pad", C.c_int32), ("k_pad", C.c_int32), ("tap_off0", C.c_int32),
        ("tap_stride", C.c_int32), ("n", C.c_int32), ("m_rows", C.c_int32), ("block_n", C.c_int32),
        ("d_frame_mask", C.c_void_p), ("rate", C.c_int32), ("out_rows", C.c_int32),
        ("up_s", C.c_int32), ("up_p", C.c_int32), ("up_cout", C.c_int32),
        ("d_bias", C.c_void_p), ("act", C.c_int32), ("slope", C.c_float), ("scale", C.c_float),
        ("d_res_f32", C.c_void_p), ("d_res_bf16", C.c_void_p), ("res_ld", C.c_int32),
        ("d_accum_in", C.c_void_p), ("d_accum_bf16", C.c_void_p), ("post_scale", C.c_float),
        ("d_out_f32", C.c_void_p), ("out_f32_ld", C.c_int32),
        ("d_out_hi", C.c_void_p), ("d_out_lo", C.c_void_p), ("out_bf_ld", C.c_int32),
        ("d_out_act", C.c_void_p), ("out_act_slope", C.c_float), ("out_act_ld", C.c_int32),
        ("d_snake_a", C.c_void_p), ("d_snake_ib", C.c_void_p),
    ]


class MrfPairArgs(C.Structure):
    _fields_ = [
        ("d_xa", C.c_void_p), ("rows", C.c_int32), ("ld", C.c_int32), ("c", C.c_int32),
        ("d_w1", C.c_void_p), ("d_w2", C.c_void_p),
        ("taps", C.c_int32), ("n_pad", C.c_int32), ("k_pad", C.c_int32), ("dilation", C.c_int32),
        ("d_b1", C.c_void_p), ("d_b2", C.c_void_p), ("slope", C.c_float),
        ("d_frame_mask", C.c_void_p), ("rate", C.c_int32),
        ("d_accum", C.c_void_p), ("accum_ld", C.c_int32),
        ("post_scale", C.c_float), ("out_slope", C.c_float),
        ("d_out", C.c_void_p), ("out_ld", C.c_int32),
    ]


class RelposAttentionArgs(C.Structure):
    _fields_ = [
        ("d_x_hi", C.c_void_p), ("d_x_lo", C.c_void_p), ("x_rows", C.c_int64),
        ("d_pos_hi", C.c_void_p), ("d_pos_lo", C.c_void_p), ("pos_rows", C.c_int32),
        ("n_head", C.c_int32), ("d_model", C.c_int32),
        ("d_seg_start", C.c_void_p), ("d_seg_len", C.c_void_p), ("nseg", C.c_int32), ("max_len", C.c_int32),
        ("d_out_hi", C.c_void_p), ("d_out_lo", C.c_void_p), ("out_ld", C.c_int32),
    ]


#: every symbol include/jatts_b200.h declares (tests/test_cabi.py checks the library exports them all)
EXPORTS = (
    "jatts_abi_version", "jatts_last_error", "jatts_launch_count",
    "jatts_fs2_create", "jatts_fs2_destroy", "jatts_fs2_plan", "jatts_fs2_run",
    "jatts_matcha_create", "jatts_matcha_destroy", "jatts_matcha_n_resnets", "jatts_matcha_plan", "jatts_matcha_run",
    "jatts_hifigan_create", "jatts_hifigan_destroy", "jatts_hifigan_run", "jatts_hifigan_run_pcm16", "jatts_op_conv_gemm", "jatts_op_mrf_pair",
    "jatts_op_relpos_attention",
    "jatts_profile_begin", "jatts_profile_end", "jatts_profile_end_classes", "jatts_debug_set_trace",
)


def _load():
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"jatts_b200: CUDA library not built ({LIB_PATH} missing). Build it with "
            "`make -C jatts_b200/csrc` (nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.jatts_abi_version.restype = C.c_int
    lib.jatts_last_error.restype = C.c_char_p
    lib.jatts_launch_count.restype = C.c_int64
    lib.jatts_fs2_create.argtypes = [C.POINTER(Fs2Config), C.POINTER(Tensor), C.c_int32, C.POINTER(C.c_void_p)]
    lib.jatts_fs2_destroy.argtypes = [C.c_void_p]
    lib.jatts_fs2_destroy.restype = None
    lib.jatts_fs2_plan.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_int32, C.c_void_p,
                                   C.c_float, C.POINTER(C.c_int32), C.c_void_p]
    lib.jatts_fs2_run.argtypes = [C.c_void_p] * 7
    lib.jatts_matcha_create.argtypes = [C.POINTER(MatchaConfig), C.POINTER(Tensor), C.c_int32, C.POINTER(C.c_void_p)]
    lib.jatts_matcha_destroy.argtypes = [C.c_void_p]
    lib.jatts_matcha_destroy.restype = None
    lib.jatts_matcha_n_resnets.argtypes = [C.c_void_p]
    lib.jatts_matcha_n_resnets.restype = C.c_int32
    lib.jatts_matcha_plan.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_int32, C.c_void_p,
                                      C.POINTER(C.c_int32), C.c_void_p]
    lib.jatts_matcha_run.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.POINTER(C.c_float), C.c_int32,
                                     C.c_void_p, C.c_void_p, C.c_void_p]
    lib.jatts_hifigan_create.argtypes = [C.POINTER(HifiganConfig), C.POINTER(Tensor), C.c_int32,
                                         C.POINTER(C.c_void_p)]
    lib.jatts_hifigan_destroy.argtypes = [C.c_void_p]
    lib.jatts_hifigan_destroy.restype = None
    lib.jatts_hifigan_run.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_int32, C.c_void_p,
                                      C.c_void_p]
    lib.jatts_hifigan_run_pcm16.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_int32, C.c_void_p,
                                            C.c_void_p]
    lib.jatts_op_conv_gemm.argtypes = [C.POINTER(ConvGemmArgs), C.c_int32, C.c_void_p]
    lib.jatts_op_mrf_pair.argtypes = [C.POINTER(MrfPairArgs), C.c_void_p]
    lib.jatts_op_relpos_attention.argtypes = [C.POINTER(RelposAttentionArgs), C.c_void_p]
    lib.jatts_debug_set_trace.argtypes = [C.c_void_p]
    lib.jatts_profile_end_classes.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int32]
    lib.jatts_profile_end.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_double),
                                      C.POINTER(C.c_int64)]
    if lib.jatts_abi_version() != 1:
        raise RuntimeError("jatts_b200: ABI version mismatch between the Python host side and the library")
    return lib


lib = _load()


def check(rc: int, what: str = "") -> None:
    """Map a negative JATTS_E_* return code to a Python exception (vocoder.py:22-24 style: fail loudly)."""
    if rc == 0:
        return
    msg = (lib.jatts_last_error() or b"").decode(errors="replace")
    text = f"jatts_b200 {what} failed (code {rc}): {msg}"
    if rc == -2:
        raise ValueError(text)
    if rc == -1:
        raise NotImplementedError(text)
    raise RuntimeError(text)


_DT = None


def tensor_table(named):
    """dict name -> torch tensor (device, contiguous)  ->  (ctypes array, keep-alive list)."""
    import torch

    global _DT
    if _DT is None:
        _DT = {torch.float32: JATTS_F32, torch.bfloat16: JATTS_BF16, torch.int64: JATTS_I64, torch.int32: JATTS_I32,
               torch.float16: JATTS_F16}
    arr = (Tensor * len(named))()
    keep = []
    for i, (k, t) in enumerate(named.items()):
        assert t.is_contiguous() and t.is_cuda, k
        kb = k.encode()
        keep.append((kb, t))
        arr[i] = Tensor(kb, t.data_ptr(), t.numel(), _DT[t.dtype])
    return arr, keep


def launch_count() -> int:
    return int(lib.jatts_launch_count())


def profile_begin() -> None:
    check(lib.jatts_profile_begin(), "profile_begin")


def profile_end():
    """-> dict(ms_bf16, n_bf16, ms_split, n_split): summed device time of the tcgen05 conv launches"""
    a, b, c, d = C.c_double(), C.c_int64(), C.c_double(), C.c_int64()
    check(lib.jatts_profile_end(C.byref(a), C.byref(b), C.byref(c), C.byref(d)), "profile_end")
    return dict(ms_bf16=a.value, n_bf16=b.value, ms_split=c.value, n_split=d.value)


PROFILE_CLASSES = ("bf16_conv", "split_gemm", "attention", "layernorm", "dwconv_swish", "length_regulate", "output_conv")


def profile_end_classes():
    """-> {class: (ms, launches)} for every kernel class of include/jatts_b200.h::jatts_profile_end_classes"""
    n = len(PROFILE_CLASSES)
    ms, cnt = (C.c_double * n)(), (C.c_int64 * n)()
    check(lib.jatts_profile_end_classes(ms, cnt, n), "profile_end_classes")
    return {name: (ms[i], cnt[i]) for i, name in enumerate(PROFILE_CLASSES)}
