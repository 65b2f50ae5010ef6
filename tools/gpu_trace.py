"""Debug: per-role timeline of CTA 0 for one big conv launch (C=32 conv2-like)."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))  # repo root
import sys, torch
sys.path.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "tests"))
from gemm_ref import Case
from jatts_b200 import _lib
c, k, dil = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
res = sys.argv[4] == "res"
m = 128 * 148 * 40
case = Case(m=m, c_in=c, n=c, taps=k, dil=dil, block_n=c if c <= 256 else 256, res="bf16" if res else None, out=("hi", "act") if res else ("hi",), act=0 if res else 2, seed=1)
trace = torch.zeros(5 * 8 * 64, dtype=torch.int64, device="cuda")
_lib.lib.jatts_debug_set_trace(trace.data_ptr())
case.run(impl=0)
_lib.lib.jatts_debug_set_trace(None)
t = trace.cpu().view(5, 8, 64)
base = int(t[t > 0].min())
names = {(0,0): "prod.A_issue", (1,0): "mma.tempty_ok", (1,1): "mma.afull_ok", (1,2): "mma.done", (2,0): "load.epempty_ok", (3,0): "store.ready_ok", (3,1): "store.released", (4,3): "epi.begin", (4,0): "epi.tfull_ok", (4,1): "epi.epfull_ok", (4,2): "epi.arrived"}
print("tile " + " ".join(f"{v:>15s}" for v in names.values()))
for i in range(8, 28):
    print(f"{i:4d} " + " ".join(f"{int(t[r, e, i]) - base:15d}" for (r, e) in names))
d = {n: (t[r, e, 20:60] - t[r, e, 19:59]).float().mean().item() for (r, e), n in names.items()}
print("mean period per tile (clk):", {k: round(v) for k, v in d.items()})
