"""Debug: per-role clock64 timeline of CTA 0 of the tcgen05 attention kernel (attention_tc.cu)."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))  # repo root
import ctypes as C
import sys
import torch
from jatts_b200 import _lib, _pack

T, B = int(sys.argv[1]), int(sys.argv[2])
H, D = 2, 384
dev = "cuda"
lens = [T] * B
seg_start = [i * (T + 8) for i in range(B)]
rows = B * (T + 8) + 64
x = torch.randn(rows, 4 * D)
pos = torch.randn(max(T, 128), D)
xh, xl = (t.to(dev) for t in _pack.split16(x))
ph, pl = (t.to(dev) for t in _pack.split16(pos))
oh = torch.zeros(rows, D, dtype=torch.float16, device=dev)
ol = torch.zeros(rows, D, dtype=torch.float16, device=dev)
ss = torch.tensor(seg_start, dtype=torch.int32, device=dev)
sl = torch.tensor(lens, dtype=torch.int32, device=dev)
a = _lib.RelposAttentionArgs(d_x_hi=xh.data_ptr(), d_x_lo=xl.data_ptr(), x_rows=rows, d_pos_hi=ph.data_ptr(), d_pos_lo=pl.data_ptr(),
                             pos_rows=pos.shape[0], n_head=H, d_model=D, d_seg_start=ss.data_ptr(), d_seg_len=sl.data_ptr(), nseg=B,
                             max_len=T, d_out_hi=oh.data_ptr(), d_out_lo=ol.data_ptr(), out_ld=D)
st = torch.cuda.current_stream().cuda_stream
_lib.check(_lib.lib.jatts_op_relpos_attention(C.byref(a), st))
best = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(_lib.lib.jatts_op_relpos_attention(C.byref(a), st))
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(f"T={T} B={B}: best of 5 = {best * 1e3:.1f} us (op entry incl. scratch allocation)")
trace = torch.zeros(5 * 8 * 64, dtype=torch.int64, device=dev)
_lib.lib.jatts_debug_set_trace(trace.data_ptr())
_lib.check(_lib.lib.jatts_op_relpos_attention(C.byref(a), st))
torch.cuda.synchronize()
_lib.lib.jatts_debug_set_trace(None)
t = trace.cpu().view(-1, 8)
base = int(t[t > 0].min())
names = {0: "pr.start", 1: "pr.QK_iss", 2: "pr.q_free", 3: "pr.s2done", 4: "pr.V_iss", 8: "mma.start", 9: "mma.q1_ok", 10: "mma.ph1_is",
         11: "mma.q2_ok", 12: "mma.ph2_is", 13: "mma.pv_ok", 14: "mma.ph3_is", 16: "w0.start", 17: "w0.s_ok", 18: "w0.S_done", 19: "w0.bar",
         20: "w0.stats", 21: "w0.P_done", 22: "w0.ctx_ok", 23: "w0.done", 24: "w4.atbar", 25: "w4.bar", 26: "w4.stats", 27: "w4.P_done"}
for it in range(4):
    print(f"--- tile iteration {it} of CTA 0 (clk since first stamp)")
    for k, n in names.items():
        v = int(t[k, it])
        if v:
            print(f"  {n:12s} {v - base:9d}")
