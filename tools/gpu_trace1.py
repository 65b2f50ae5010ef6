"""Debug: per-role timeline of CTA 0 for one FS2-like split GEMM launch (v1 kernel)."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))  # repo root
import sys, torch
sys.path.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "tests"))
from gemm_ref import Case
from jatts_b200 import _lib
A = _lib
kind = sys.argv[1]
m = 20069
if kind == "w1":
    case = Case(m=m, c_in=384, n=1536, taps=3, split_mode=True, act=A.ACT_RELU, out=("hi", "lo"), seed=1)
elif kind == "w2":
    case = Case(m=m, c_in=1536, n=384, taps=3, split_mode=True, res="f32", scale=0.5, out=("f32",), seed=1)
else:
    case = Case(m=m, c_in=384, n=1152, taps=1, split_mode=True, out=("f32",), seed=1)
trace = torch.zeros(5 * 8 * 64, dtype=torch.int64, device="cuda")
_lib.lib.jatts_debug_set_trace(trace.data_ptr())
case.run(impl=0)
_lib.lib.jatts_debug_set_trace(None)
t = trace.cpu().view(5, 8, 64)
base = int(t[t > 0].min())
names = {(0,0): "prod.tile0", (1,0): "mma.begin", (1,2): "mma.end", (4,3): "epi.begin", (4,0): "epi.drained", (4,2): "epi.stored"}
print(kind, "tile " + " ".join(f"{v:>12s}" for v in names.values()))
for i in range(0, 10):
    print(f"{i:4d} " + " ".join(f"{int(t[r, e, i]) - base:12d}" for (r, e) in names))
