"""Debug: per-role clock64 timeline of CTA 0 of the fused residual-unit kernel (mrf_pair.cu) on one big launch."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))  # repo root
import ctypes as C
import sys
import torch
from jatts_b200 import _lib

c, k, d = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
rows = (128 - (k - 1)) * 148 * 60
dev = "cuda"
xa = torch.randn(rows, c, device=dev).to(torch.bfloat16)
w1 = torch.randn(k, c, 64, device=dev).to(torch.bfloat16) * 0.05
w2 = torch.randn(k, c, 64, device=dev).to(torch.bfloat16) * 0.05
b = torch.zeros(c, device=dev)
out = torch.empty_like(xa)
a = _lib.MrfPairArgs()
a.d_xa, a.rows, a.ld, a.c = xa.data_ptr(), rows, c, c
a.d_w1, a.d_w2, a.taps, a.n_pad, a.k_pad, a.dilation = w1.data_ptr(), w2.data_ptr(), k, c, 64, d
a.d_b1, a.d_b2, a.slope, a.rate = b.data_ptr(), b.data_ptr(), 0.1, 1
a.post_scale, a.out_slope, a.d_out, a.out_ld = 1.0, 0.1, out.data_ptr(), c
st = torch.cuda.current_stream().cuda_stream
_lib.check(_lib.lib.jatts_op_mrf_pair(C.byref(a), st))
trace = torch.zeros(5 * 8 * 64, dtype=torch.int64, device=dev)
_lib.lib.jatts_debug_set_trace(trace.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
_lib.check(_lib.lib.jatts_op_mrf_pair(C.byref(a), st))
e1.record()
torch.cuda.synchronize()
_lib.lib.jatts_debug_set_trace(None)
print(f"C={c} k={k} d={d}: {e0.elapsed_time(e1) * 1e3:.1f} us for 60 tiles/CTA -> {e0.elapsed_time(e1) * 1e3 / 60:.2f} us per tile")
t = trace.cpu().view(5, 8, 64)
base = int(t[t > 0].min())
names = {(0, 0): "load.issue", (1, 4): "mma.xa_ok", (1, 0): "mma.c1_go", (1, 1): "mma.c1_iss", (2, 0): "e1.T_ok", (2, 1): "e1.t_free", (2, 3): "e1.ld0", (2, 4): "e1.math", (2, 5): "e1.fence",
         (2, 2): "e1.done", (1, 5): "mma.t_ok", (1, 2): "mma.c2_go", (1, 3): "mma.c2_iss", (3, 0): "e2.U_ok", (3, 2): "e2.ld0", (3, 3): "e2.math", (3, 1): "e2.done",
         (4, 0): "st.ready", (4, 1): "st.freed"}
print("tile " + " ".join(f"{v:>10s}" for v in names.values()))
for i in range(20, 28):
    print(f"{i:4d} " + " ".join(f"{int(t[r, e, i]) - base:10d}" for (r, e) in names))
d_ = {n: (t[r, e, 21:58] - t[r, e, 20:57]).float().mean().item() for (r, e), n in names.items()}
print("mean period per tile (clk):", round(d_["mma.c1_go"]))
