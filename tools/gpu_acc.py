"""Diagnostic: fp32 accumulation behaviour of tcgen05.mma (bias / spread vs K)."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))  # repo root
import sys, zlib
import torch
sys.path.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "tests"))
from gemm_ref import Case

for positive in (False, True):
    for k in (64, 384, 1152, 4608):
        for impl in (0, 1):
            case = Case(m=256, c_in=k, n=128, block_n=128, seed=1)
            if positive:
                case.a_hi = case.a_hi.abs(); case.w_hi = case.w_hi.abs()
                case.a_eff = case.a_hi.double(); case.w_eff = case.w_hi.double()
            case.bias = None
            out = case.run(impl=impl)["f32"].double()
            ref, _ = case.reference()
            rel = (out - ref) / ref.abs().clamp_min(1e-3)
            print(f"positive={positive} K={k:5d} impl={impl} mean_rel={float(rel.mean()):+.3e} std_rel={float(rel.std()):.3e} max_abs={float((out-ref).abs().max()):.3e} ref_rms={float(ref.pow(2).mean().sqrt()):.3f}", flush=True)
