"""GPU parity of the fused HiFi-GAN residual-unit kernel (jatts_op_mrf_pair, csrc/mrf_pair.cu) against an
fp64 torch statement of the same unit (oracle/hifigan.py resblock inner step:
x + conv(k,1)(lrelu(conv(k,d)(lrelu(x))))) on the packed-with-gaps layout's masking rules."""
import ctypes as C
import zlib

import pytest
import torch
import torch.nn.functional as F

from jatts_b200 import _lib

SENTINEL = 768.0


def lrelu(x, s):
    return torch.where(x >= 0, x, x * s)


def run_case(c, k, d, rows, mask_rate=0, accum=False, post_scale=1.0, out_slope=0.1, slope=0.1, seed=0, ld_extra=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(rows, c, generator=g)
    mask = None
    if mask_rate:
        nm = (rows + mask_rate - 1) // mask_rate
        mask = (torch.rand(nm, generator=g) > 0.2).to(torch.uint8)
        rowmask = mask[torch.arange(rows) // mask_rate].bool()
        x = x * rowmask[:, None]          # gap rows of the layout are zero
    else:
        rowmask = torch.ones(rows, dtype=torch.bool)
    xa = lrelu(x, slope).to(torch.bfloat16)
    w1 = (torch.randn(k, c, c, generator=g) / (k * c) ** 0.5).to(torch.bfloat16)   # [tap][out][in]
    w2 = (torch.randn(k, c, c, generator=g) / (k * c) ** 0.5).to(torch.bfloat16)
    b1 = torch.randn(c, generator=g) * 0.3
    b2 = torch.randn(c, generator=g) * 0.3
    acc = torch.randn(rows, c, generator=g).to(torch.bfloat16) if accum else None

    # ---- fp64 reference (t rounded to bf16 as the kernel keeps it) ----
    xa64 = xa.double()
    xr = torch.where(xa64 >= 0, xa64, xa64 * float(torch.tensor(1.0 / slope, dtype=torch.float32)))
    wt1 = w1.double().permute(1, 2, 0)    # [out, in, tap]
    wt2 = w2.double().permute(1, 2, 0)
    t = F.conv1d(xa64.t()[None], wt1, b1.double(), dilation=d, padding=(k - 1) // 2 * d)[0].t()
    t = (lrelu(t, slope) * rowmask[:, None]).to(torch.bfloat16).double()
    v = F.conv1d(t.t()[None], wt2, b2.double(), padding=(k - 1) // 2)[0].t() + xr
    if acc is not None:
        v = v + acc.double()
    v = v * post_scale * rowmask[:, None]
    ref = lrelu(v, out_slope)

    dev = "cuda"
    k_pad = 64
    ld = c + ld_extra

    def pad_w(w):
        buf = torch.zeros(k, c, k_pad, dtype=torch.bfloat16)
        buf[:, :, :c] = w
        return buf.to(dev)

    xa_d = torch.zeros(rows, ld, dtype=torch.bfloat16)
    xa_d[:, :c] = xa
    xa_d = xa_d.to(dev)
    w1_d, w2_d, b1_d, b2_d = pad_w(w1), pad_w(w2), b1.to(dev), b2.to(dev)
    out = torch.full((rows, ld), SENTINEL, dtype=torch.bfloat16, device=dev)
    a = _lib.MrfPairArgs()
    a.d_xa, a.rows, a.ld, a.c = xa_d.data_ptr(), rows, ld, c
    a.d_w1, a.d_w2 = w1_d.data_ptr(), w2_d.data_ptr()
    a.taps, a.n_pad, a.k_pad, a.dilation = k, c, k_pad, d
    a.d_b1, a.d_b2, a.slope = b1_d.data_ptr(), b2_d.data_ptr(), slope
    mask_d = mask.to(dev) if mask is not None else None
    a.d_frame_mask, a.rate = (mask_d.data_ptr() if mask_d is not None else None), max(mask_rate, 1)
    acc_d = None
    if acc is not None:
        acc_d = torch.zeros(rows, ld, dtype=torch.bfloat16)
        acc_d[:, :c] = acc
        acc_d = acc_d.to(dev)
        a.d_accum, a.accum_ld = acc_d.data_ptr(), ld
    a.post_scale, a.out_slope = post_scale, out_slope
    a.d_out, a.out_ld = out.data_ptr(), ld
    _lib.check(_lib.lib.jatts_op_mrf_pair(C.byref(a), torch.cuda.current_stream().cuda_stream), "op_mrf_pair")
    torch.cuda.synchronize()
    got = out.cpu().double()
    if ld_extra:
        assert bool((got[:, c:] == SENTINEL).all()), "columns beyond C were written"
    got = got[:, :c]
    err = float((got - ref).abs().max() / max(1.0, float(ref.abs().max())))
    gaps_zero = bool((got[~rowmask] == 0).all()) if (~rowmask).any() else True
    return err, gaps_zero


CASES = {
    "c32_k3_d1": dict(c=32, k=3, d=1, rows=2000),
    "c32_k3_d5": dict(c=32, k=3, d=5, rows=1000, mask_rate=300),
    "c32_k7_d3": dict(c=32, k=7, d=3, rows=3000, mask_rate=300, accum=True, out_slope=1.0),
    "c32_k11_d5": dict(c=32, k=11, d=5, rows=5000, mask_rate=300, accum=True, post_scale=1 / 3, out_slope=0.01),
    "c32_k11_d1_wrap": dict(c=32, k=11, d=1, rows=118 * 148 * 3 + 55, mask_rate=300),
    "c64_k3_d3": dict(c=64, k=3, d=3, rows=1500, mask_rate=100),
    "c64_k7_d5": dict(c=64, k=7, d=5, rows=2500, mask_rate=100, accum=True, out_slope=1.0),
    "c64_k7_d1_wrap": dict(c=64, k=7, d=1, rows=122 * 148 * 3 + 9, mask_rate=100),
    "c64_k11_d5_stream": dict(c=64, k=11, d=5, rows=4000, mask_rate=100, accum=True, post_scale=1 / 3),
    "c64_k11_d3_stream_wrap": dict(c=64, k=11, d=3, rows=118 * 148 * 4 + 3, mask_rate=100),
    "c64_k3_tiny": dict(c=64, k=3, d=1, rows=7),
    "c32_ld_wider": dict(c=32, k=7, d=1, rows=700, ld_extra=32),
}


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_mrf_pair(name):
    err, gaps_zero = run_case(seed=zlib.crc32(name.encode()) % 1000, **CASES[name])
    assert gaps_zero, f"{name}: rows outside every utterance are not zero"
    assert err < 1e-2, f"{name}: rel err {err:.3e}"
