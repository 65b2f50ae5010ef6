"""Debug: per-role timeline of CTA 0 for FS2-like split GEMM launches + event timing of each shape."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))  # repo root
import sys, torch
sys.path.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "tests"))
from gemm_ref import Case
from jatts_b200 import _lib
A = _lib
shapes = {
    "dec_qkv": dict(m=20069, c_in=384, n=1536, taps=1, out=("hi", "lo")),
    "dec_out": dict(m=20069, c_in=384, n=384, taps=1, res="f32", out=("f32",)),
    "dec_w1": dict(m=20069, c_in=384, n=1536, taps=3, act=A.ACT_RELU, out=("hi", "lo")),
    "dec_w2": dict(m=20069, c_in=1536, n=384, taps=3, res="f32", scale=0.5, out=("f32",)),
    "enc_qkv": dict(m=3712, c_in=384, n=1536, taps=1, out=("hi", "lo")),
    "enc_w1": dict(m=3712, c_in=384, n=1536, taps=3, act=A.ACT_RELU, out=("hi", "lo")),
    "enc_w2": dict(m=3712, c_in=1536, n=384, taps=3, res="f32", scale=0.5, out=("f32",)),
}
names = {(0,0): "prod.tile0", (1,0): "mma.begin", (1,2): "mma.end", (4,3): "epi.begin", (4,0): "epi.drained", (4,2): "epi.stored"}
for kind in sys.argv[1:]:
    case = Case(split_mode=True, seed=1, **shapes[kind])
    case.run(impl=0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    trace = torch.zeros(5 * 8 * 64, dtype=torch.int64, device="cuda")
    _lib.lib.jatts_debug_set_trace(trace.data_ptr())
    case.run(impl=0)
    _lib.lib.jatts_debug_set_trace(None)
    t = trace.cpu().view(5, 8, 64)
    base = int(t[t > 0].min())
    print(kind, "tile " + " ".join(f"{v:>12s}" for v in names.values()))
    for i in range(0, 8):
        print(f"{i:4d} " + " ".join(f"{int(t[r, e, i]) - base if int(t[r, e, i]) else -1:12d}" for (r, e) in names))
