"""Reference semantics of jatts_op_conv_gemm in fp64 torch (test helper) + a ctypes launcher."""
import ctypes as C

import torch

from jatts_b200 import _lib
from jatts_b200._pack import round_up


def split(x):
    """fp16 (hi, lo*2^11) operand pair, jatts_b200/_pack.py::split16"""
    hi = x.to(torch.float16)
    lo = ((x - hi.float()) * 2048.0).to(torch.float16)
    return hi, lo


class Case:
    """One conv-GEMM problem with all optional epilogue features; tensors live on CPU until run()."""

    def __init__(self, m, c_in, n, taps=1, dil=1, block_n=128, split_mode=False, act=0, slope=0.1, scale=1.0,
                 bias=True, res=None, accum=False, post_scale=1.0, out=("f32",), mask_rate=0, up_s=0, up_cout=0,
                 a_ld=None, seed=0, tap_off0=None, tap_stride=None, amp=1.0):
        g = torch.Generator().manual_seed(seed)
        self.m, self.c_in, self.n, self.taps, self.dil, self.block_n = m, c_in, n, taps, dil, block_n
        self.split, self.act, self.slope, self.scale = split_mode, act, slope, scale
        self.res, self.accum, self.post_scale, self.out = res, accum, post_scale, out
        self.up_s, self.up_cout = up_s, up_cout
        self.up_p = up_s // 2 + up_s % 2 if up_s else 0
        self.k_pad = round_up(c_in, 64)
        self.a_ld = a_ld or c_in
        n_cols = 2 * n if act == _lib.ACT_GLU else n
        self.n_cols = n_cols
        self.n_pad = round_up(n_cols, block_n)
        self.tap_off0 = -((taps - 1) // 2) * dil if tap_off0 is None else tap_off0
        self.tap_stride = dil if tap_stride is None else tap_stride
        a = torch.randn(m, c_in, generator=g) * amp
        w = torch.randn(taps, n_cols, c_in, generator=g) / (taps * c_in) ** 0.5
        if split_mode:
            self.a_hi, self.a_lo = split(a)
            self.w_hi, self.w_lo = split(w)
            self.a_eff = self.a_hi.double() + self.a_lo.double() / 2048.0
            self.w_eff = self.w_hi.double() + self.w_lo.double() / 2048.0
        else:
            self.a_hi, self.a_lo = a.to(torch.bfloat16), None
            self.w_hi, self.w_lo = w.to(torch.bfloat16), None
            self.a_eff, self.w_eff = self.a_hi.double(), self.w_hi.double()
        self.bias = torch.randn(n_cols if not up_s else up_cout, generator=g) * 0.3 if bias else None
        self.out_rows = m * up_s if up_s else m
        self.out_cols = up_cout if up_s else n
        self.rate = 1
        self.mask = None
        if mask_rate:
            self.rate = mask_rate
            nm = (self.out_rows + mask_rate - 1) // mask_rate
            self.mask = (torch.rand(nm, generator=g) > 0.25).to(torch.uint8)
        self.res_t = None
        if res == "f32":
            self.res_t = torch.randn(self.out_rows, self.out_cols, generator=g)
        elif res == "bf16":
            self.res_t = torch.randn(self.out_rows, self.out_cols, generator=g).to(torch.bfloat16)
        self.acc_t = torch.randn(self.out_rows, self.out_cols, generator=g) if accum else None
        if accum == "bf16":
            self.acc_t = self.acc_t.to(torch.bfloat16)
        self.sentinel = 768.0  # exactly representable in bf16 / fp16
        # SnakeBeta constants per output column (transformer.py:28-102 with alpha_logscale: a = exp(alpha), ib = 1 / (exp(beta) + 1e-9))
        self.snake_a = torch.exp(0.3 * torch.randn(self.n_pad, generator=g)) if act == _lib.ACT_SNAKE else None
        self.snake_ib = 1.0 / (torch.exp(0.3 * torch.randn(self.n_pad, generator=g)) + 1e-9) if act == _lib.ACT_SNAKE else None

    # ---- fp64 reference --------------------------------------------------------------------------
    def reference(self):
        m = self.m
        acc = torch.zeros(m, self.n_cols, dtype=torch.float64)
        for j in range(self.taps):
            off = self.tap_off0 + j * self.tap_stride
            src = torch.zeros(m, self.c_in, dtype=torch.float64)
            lo, hi = max(0, -off), min(m, m - off)
            if hi > lo:
                src[lo:hi] = self.a_eff[lo + off:hi + off]
            acc += src @ self.w_eff[j].t()
        if self.act == _lib.ACT_GLU:
            half = self.block_n // 2
            d = self.n
            idx_a, idx_b = [], []
            for t in range(d // half):
                idx_a += list(range(t * self.block_n, t * self.block_n + half))
                idx_b += list(range(t * self.block_n + half, (t + 1) * self.block_n))
            b = self.bias.double() if self.bias is not None else torch.zeros(self.n_cols, dtype=torch.float64)
            v = (acc[:, idx_a] + b[idx_a]) * torch.sigmoid(acc[:, idx_b] + b[idx_b]) * self.scale
            rows = torch.arange(m)
            out = v
        elif self.up_s:
            s, p, co = self.up_s, self.up_p, self.up_cout
            out = torch.full((self.out_rows, co), float("nan"), dtype=torch.float64)
            for q in range(s):
                rows = torch.arange(m) * s + q - p
                ok = (rows >= 0) & (rows < self.out_rows)
                out[rows[ok]] = acc[ok][:, q * co:(q + 1) * co]
            # the last p output rows need input row m (the zero gap row in the engine's layout), which
            # this stand-alone problem does not have: they are simply not produced
            self._uncovered = torch.isnan(out).any(1)
            out = torch.nan_to_num(out)
            if self.bias is not None:
                out = out + self.bias.double()
            out = self._act(out) * self.scale
        else:
            out = acc
            if self.bias is not None:
                out = out + self.bias.double()
            out = self._act(out) * self.scale
        if self.res_t is not None:
            out = out + self.res_t.double()
        if self.acc_t is not None:
            out = out + self.acc_t.double()
        out = out * self.post_scale
        valid = torch.ones(self.out_rows, dtype=torch.bool)
        if self.up_s:
            valid &= ~self._uncovered
        if self.mask is not None:
            valid &= self.mask[torch.arange(self.out_rows) // self.rate].bool()
        return out, valid

    def _act(self, x):
        if self.act == _lib.ACT_RELU:
            return torch.relu(x)
        if self.act == _lib.ACT_LRELU:
            return torch.where(x > 0, x, x * self.slope)
        if self.act == _lib.ACT_TANH:
            return torch.tanh(x)
        if self.act == _lib.ACT_SNAKE:
            n = x.shape[1]
            return x + self.snake_ib[:n].double() * torch.sin(x * self.snake_a[:n].double()) ** 2
        return x

    # ---- device run --------------------------------------------------------------------------------
    def run(self, impl=0, dev="cuda"):
        def pad_a(t):
            if t is None:
                return None
            buf = torch.zeros(self.m, self.a_ld, dtype=t.dtype)
            buf[:, :self.c_in] = t
            return buf.to(dev)

        def pad_w(t):
            if t is None:
                return None
            buf = torch.zeros(self.taps, self.n_pad, self.k_pad, dtype=t.dtype)
            buf[:, :self.n_cols, :self.c_in] = t
            return buf.to(dev)

        a_hi, a_lo, w_hi, w_lo = pad_a(self.a_hi), pad_a(self.a_lo), pad_w(self.w_hi), pad_w(self.w_lo)
        args = _lib.ConvGemmArgs()
        args.d_a_hi, args.d_a_lo = a_hi.data_ptr(), (a_lo.data_ptr() if a_lo is not None else None)
        args.a_rows, args.a_ld, args.a_cols = self.m, self.a_ld, self.c_in
        args.d_w_hi, args.d_w_lo = w_hi.data_ptr(), (w_lo.data_ptr() if w_lo is not None else None)
        args.taps, args.n_pad, args.k_pad = self.taps, self.n_pad, self.k_pad
        args.tap_off0, args.tap_stride = self.tap_off0, self.tap_stride
        args.n, args.m_rows, args.block_n = self.n, self.m, self.block_n
        mask = self.mask.to(dev) if self.mask is not None else None
        args.d_frame_mask = mask.data_ptr() if mask is not None else None
        args.rate, args.out_rows = self.rate, self.out_rows
        args.up_s, args.up_p, args.up_cout = self.up_s, self.up_p, self.up_cout
        bias = self.bias.to(dev) if self.bias is not None else None
        args.d_bias = bias.data_ptr() if bias is not None else None
        args.act, args.slope, args.scale = self.act, self.slope, self.scale
        sn_a = self.snake_a.to(dev) if self.snake_a is not None else None
        sn_ib = self.snake_ib.to(dev) if self.snake_ib is not None else None
        args.d_snake_a = sn_a.data_ptr() if sn_a is not None else None
        args.d_snake_ib = sn_ib.data_ptr() if sn_ib is not None else None
        res = self.res_t.to(dev) if self.res_t is not None else None
        if self.res == "f32":
            args.d_res_f32 = res.data_ptr()
        elif self.res == "bf16":
            args.d_res_bf16 = res.data_ptr()
        args.res_ld = self.out_cols
        acc = self.acc_t.to(dev) if self.acc_t is not None else None
        if acc is not None and acc.dtype == torch.bfloat16:
            args.d_accum_bf16 = acc.data_ptr()
        elif acc is not None:
            args.d_accum_in = acc.data_ptr()
        args.post_scale = self.post_scale
        outs = {}
        ld = self.out_cols
        if "f32" in self.out:
            outs["f32"] = torch.full((self.out_rows, ld), self.sentinel, device=dev)
            args.d_out_f32 = outs["f32"].data_ptr()
        args.out_f32_ld = ld
        if "hi" in self.out:
            outs["hi"] = torch.full((self.out_rows, ld), self.sentinel, device=dev,
                                    dtype=torch.float16 if "lo" in self.out else torch.bfloat16)
            args.d_out_hi = outs["hi"].data_ptr()
        if "lo" in self.out:
            outs["lo"] = torch.full((self.out_rows, ld), self.sentinel, device=dev, dtype=torch.float16)
            args.d_out_lo = outs["lo"].data_ptr()
        args.out_bf_ld = ld
        if "act" in self.out:
            outs["act"] = torch.full((self.out_rows, ld), self.sentinel, device=dev, dtype=torch.bfloat16)
            args.d_out_act = outs["act"].data_ptr()
        args.out_act_slope, args.out_act_ld = 0.1, ld
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib.jatts_op_conv_gemm(C.byref(args), impl, stream), "op_conv_gemm")
        torch.cuda.synchronize()
        return {k: v.cpu() for k, v in outs.items()}

    def compare(self, outs):
        """returns dict of max abs errors (valid rows) and whether masked rows kept the sentinel."""
        ref, valid = self.reference()
        res = {}
        inv = ~valid
        for k, v in outs.items():
            vd = v.double()
            if k == "f32":
                res[k] = float((vd[valid] - ref[valid]).abs().max()) if valid.any() else 0.0
            elif k == "hi":
                res[k] = float((vd[valid] - ref[valid]).abs().max() / max(1.0, float(ref[valid].abs().max())))
            elif k == "lo":
                tot = outs["hi"].double() + vd / 2048.0
                res["hi+lo"] = float((tot[valid] - ref[valid]).abs().max())
            elif k == "act":
                r = torch.where(ref > 0, ref, ref * 0.1)
                res[k] = float((vd[valid] - r[valid]).abs().max() / max(1.0, float(r[valid].abs().max())))
            if inv.any():
                # rows outside every utterance are either never written (register epilogue) or written
                # as zeros (TMA epilogue): both keep the layout's gap rows zero
                res[k + "_masked_untouched"] = bool(((vd[inv] == self.sentinel) | (vd[inv] == 0)).all())
        return res
