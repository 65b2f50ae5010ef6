"""Event-timed launches of the six fused residual-unit shapes of the bench workload (best of 5)."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))  # repo root
import ctypes as C
import torch
from jatts_b200 import _lib
import jatts_b200  # noqa

dev = "cuda"
st = torch.cuda.current_stream().cuda_stream
cfgs = [(32, 3, 1), (32, 7, 3), (32, 11, 5), (64, 3, 1), (64, 7, 3), (64, 11, 5)]
frames = 19557 + 64 * 8
for c, k, d in cfgs:
    rows = frames * (300 if c == 32 else 100)
    xa = torch.randn(rows, c, device=dev).to(torch.bfloat16)
    w1 = (torch.randn(k, c, 64, device=dev) * 0.05).to(torch.bfloat16)
    w2 = (torch.randn(k, c, 64, device=dev) * 0.05).to(torch.bfloat16)
    b = torch.zeros(c, device=dev)
    out = torch.empty_like(xa)
    a = _lib.MrfPairArgs()
    a.d_xa, a.rows, a.ld, a.c = xa.data_ptr(), rows, c, c
    a.d_w1, a.d_w2, a.taps, a.n_pad, a.k_pad, a.dilation = w1.data_ptr(), w2.data_ptr(), k, c, 64, d
    a.d_b1, a.d_b2, a.slope, a.rate = b.data_ptr(), b.data_ptr(), 0.1, 1
    a.post_scale, a.out_slope, a.d_out, a.out_ld = 1.0, 0.1, out.data_ptr(), c
    best = 1e9
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(_lib.lib.jatts_op_mrf_pair(C.byref(a), st))
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    flop = 2.0 * rows * c * c * k * 2
    print(f"C={c} k={k} d={d}: {best * 1e3:7.1f} us  {flop / best / 1e9:7.1f} TFLOP/s")
    del xa, out
    torch.cuda.empty_cache()
