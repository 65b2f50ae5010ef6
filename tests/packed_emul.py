"""CPU emulation of the CUDA engines' dataflow FROM THE PACKED WEIGHT TABLES (test helper).

Lets the `-m "not gpu"` suite verify the host-side repacking (BatchNorm folding, GLU interleave,
tap-major layout, polyphase transposed conv, precomputed positional tables, closed-form rel-shift)
against the oracle without a GPU.  Never imported by the product.
"""
import math

import torch
import torch.nn.functional as F

from jatts_b200._pack import pick_block_n, round_up


def W(p, name, n, k):
    """[taps, n_pad, k_pad] hi(+lo) -> fp32 [taps, n, k]"""
    w = p[name + ".hi"].float()
    if name + ".lo" in p:
        w = w + p[name + ".lo"].float() / 2048.0
    return w[:, :n, :k]


def conv(p, name, x, n, dil=1):
    """x: [T, C_in] -> [T, n] 'same' conv from the packed taps"""
    k = x.shape[1]
    w = W(p, name, n, k)
    taps = w.shape[0]
    pad = (taps - 1) // 2 * dil
    xp = F.pad(x, (0, 0, pad, pad))
    out = torch.zeros(x.shape[0], n)
    for j in range(taps):
        out += xp[j * dil:j * dil + x.shape[0]] @ w[j].t()
    if name + ".b" in p:
        out = out + p[name + ".b"][:n]
    return out


def ln(p, name, x):
    return F.layer_norm(x, (x.shape[-1],), p[name + ".g"], p[name + ".b"], 1e-12)


def rel_attention(p, pre, x4, n_head):
    """x4 = projection output [q + bias_u | q + bias_v | k | v]; pos table as an fp16 (hi, lo * 2^11) pair"""
    t, d4 = x4.shape
    d = d4 // 4
    dk = d // n_head
    q_u, q_v, k, v = x4[:, :d], x4[:, d:2 * d], x4[:, 2 * d:3 * d], x4[:, 3 * d:]
    pos = (p[pre + "pos.hi"].float() + p[pre + "pos.lo"].float() / 2048.0)[:t]
    out = torch.zeros(t, d)
    for h in range(n_head):
        sl = slice(h * dk, (h + 1) * dk)
        qu = q_u[:, sl]
        qv = q_v[:, sl]
        ac = qu @ k[:, sl].t()
        bd = qv @ pos[:, sl].t()
        sh = torch.zeros(t, t)
        for a in range(t):  # closed form of the legacy rel_shift
            for b in range(t):
                if b <= a:
                    sh[a, b] = bd[a, t - 1 - a + b]
                elif b > a + 1:
                    sh[a, b] = bd[a + 1, b - a - 2]
        att = torch.softmax((ac + sh) / math.sqrt(dk), dim=-1)
        out[:, sl] = att @ v[:, sl]
    return out


def conformer(p, pre, x, n_layers, n_head, units):
    d = x.shape[1]
    for i in range(n_layers):
        q = f"{pre}.{i}."
        h = ln(p, q + "ln_ffm", x)
        x = x + 0.5 * conv(p, q + "ffm_w2", torch.relu(conv(p, q + "ffm_w1", h, units)), d)
        h = ln(p, q + "ln_mha", x)
        x = x + conv(p, q + "out", rel_attention(p, q, conv(p, q + "qkv", h, 4 * d), n_head), d)
        h = ln(p, q + "ln_conv", x)
        g2 = conv(p, q + "pw1", h, 2 * d)  # interleaved [64 a | 64 gate] per 128 tile
        g2 = g2.view(-1, d // 64, 2, 64)
        g = (g2[:, :, 0] * torch.sigmoid(g2[:, :, 1])).reshape(-1, d)
        wT, bb = p[q + "dw.wT"], p[q + "dw.b"]
        kk = wT.shape[0]
        gp = F.pad(g, (0, 0, (kk - 1) // 2, (kk - 1) // 2))
        y = bb.expand(g.shape[0], d).clone()
        for j in range(kk):
            y = y + gp[j:j + g.shape[0]] * wT[j]
        y = y * torch.sigmoid(y)
        x = x + conv(p, q + "pw2", y, d)
        h = ln(p, q + "ln_ff", x)
        x = x + 0.5 * conv(p, q + "ff_w2", torch.relu(conv(p, q + "ff_w1", h, units)), d)
        x = ln(p, q + "ln_final", x)
    return ln(p, pre + ".after_norm", x)


def predictor(p, pre, hs, n_layers, chans):
    h = hs
    for i in range(n_layers):
        h = ln(p, f"{pre}.ln{i}", torch.relu(conv(p, f"{pre}.conv{i}", h, chans)))
    return h @ p[pre + ".lin_w"] + p[pre + ".lin_b"]


def emul_fs2(p, cfg, text, spemb=None, alpha=1.0):
    d = cfg["adim"]
    x = p["emb"][text] * math.sqrt(d)
    hs = conformer(p, "enc", x, cfg["elayers"], cfg["aheads"], cfg["eunits"])
    if spemb is not None:
        hs = hs + (F.normalize(spemb.unsqueeze(0)).squeeze(0) @ p["spk.w"].t() + p["spk.b"])
    pit = predictor(p, "pitch", hs, cfg["pitch_predictor_layers"], cfg["pitch_predictor_chans"])
    en = predictor(p, "energy", hs, cfg["energy_predictor_layers"], cfg["energy_predictor_chans"])
    logd = predictor(p, "dur", hs, cfg["duration_predictor_layers"], cfg["duration_predictor_chans"])
    dur = torch.clamp(torch.round(logd.exp() - 1.0), min=0).long()
    dl = dur if alpha == 1.0 else torch.round(dur.float() * alpha).long()
    dret = dur
    if int(dl.sum()) == 0:
        dl = torch.ones_like(dl)
        if alpha == 1.0:
            dret = dl
    cum = torch.cumsum(dl, 0)
    idx = torch.searchsorted(cum, torch.arange(int(cum[-1])), right=True)
    hs2 = (hs + (en[:, None] * p["energy_embed.w"] + p["energy_embed.b"])) + (pit[:, None] * p["pitch_embed.w"] + p["pitch_embed.b"])
    z = conformer(p, "dec", hs2[idx] * math.sqrt(d), cfg["dlayers"], cfg["aheads"], cfg["dunits"])
    before = conv(p, "feat_out", z, cfg["odim"])
    h = before
    nl = cfg["postnet_layers"]
    for i in range(nl):
        co = cfg["odim"] if i == nl - 1 else cfg["postnet_chans"]
        h = conv(p, f"postnet{i}", h, co)
        if i != nl - 1:
            h = torch.tanh(h)
    return dict(feat_gen=before + h, duration=dret, pitch=pit.unsqueeze(-1), energy=en.unsqueeze(-1), lr_index=idx)


def emul_hifigan(p, cfg, mel):
    """mel (T, 80) -> (T*hop,) using the packed taps incl. the polyphase transposed conv"""
    lrelu = lambda t, s=cfg["slope"]: torch.where(t > 0, t, t * s)
    x = mel * p["mel_scale"] + p["mel_shift"]
    x = conv(p, "input_conv", x, cfg["channels"])
    nb = len(cfg["resblock_kernel_sizes"])
    for i, s in enumerate(cfg["upsample_scales"]):
        co = cfg["channels"] >> (i + 1)
        ci = cfg["channels"] >> i
        xa = lrelu(x)
        w = W(p, f"ups{i}", s * co, ci)  # [2, s*co, ci]
        L = xa.shape[0]
        xpad = torch.cat([xa, torch.zeros(1, ci)], 0)            # row j = L reads the zero gap row
        xprev = torch.cat([torch.zeros(1, ci), xa], 0)           # x[j-1]
        acc = xpad @ w[0].t() + xprev @ w[1].t()                  # [L+1, s*co]
        pp = s // 2 + s % 2
        out = torch.zeros(L * s, co)
        for q in range(s):
            rows = torch.arange(L + 1) * s + q - pp
            ok = (rows >= 0) & (rows < L * s)
            out[rows[ok]] = acc[ok][:, q * co:(q + 1) * co]
        x0 = out + p[f"ups{i}.b"]
        total = 0
        for j in range(nb):
            xr = x0
            for d, dil in enumerate(cfg["resblock_dilations"][j]):
                t = lrelu(conv(p, f"rb{i}_{j}.c1_{d}", lrelu(xr), co, dil))
                xr = conv(p, f"rb{i}_{j}.c2_{d}", t, co) + xr
            total = total + xr
        x = total / nb
    xa = torch.where(x > 0, x, x * 0.01)
    wo = p["output.w"]  # [k, C]
    k = wo.shape[0]
    xp = F.pad(xa, (0, 0, (k - 1) // 2, (k - 1) // 2))
    y = p["output.b"].expand(x.shape[0]).clone()
    for j in range(k):
        y = y + xp[j:j + x.shape[0]] @ wo[j]
    return torch.tanh(y)


# --------------------------------------------------------------------------------------------------------------
# Matcha-TTS: the dataflow of jatts_b200/csrc/engine_matcha.cu from the table of _pack.py::pack_matcha
# --------------------------------------------------------------------------------------------------------------
def conv_off(p, name, x, n, tap_off0):
    """convolution with explicit first tap offset (the paired-row views): out[r] = sum_j x[r + tap_off0 + j] W_j"""
    k = x.shape[1]
    w = W(p, name, n, k)
    taps = w.shape[0]
    lo, hi = max(0, -tap_off0), max(0, tap_off0 + taps - 1)
    xp = F.pad(x, (0, 0, lo, hi))
    out = torch.zeros(x.shape[0], n)
    for j in range(taps):
        s0 = lo + tap_off0 + j
        out += xp[s0:s0 + x.shape[0]] @ w[j].t()
    return out + p[name + ".b"][:n]


def gn_mish(p, name, x, add=None):
    y = F.mish(F.group_norm(x.t().unsqueeze(0), 8, p[name + ".g"], p[name + ".b"], 1e-5)[0].t())
    return y if add is None else y + add


def resnet(p, r, x, temb_r, c):
    h = gn_mish(p, f"dec.res{r}.gn1", conv(p, f"dec.res{r}.c1", x, c), temb_r)
    h = gn_mish(p, f"dec.res{r}.gn2", conv(p, f"dec.res{r}.c2", h, c))
    return h + conv(p, f"dec.res{r}.res", x, c)


def transformer(p, name, x, heads, inner):
    c = x.shape[1]
    h = F.layer_norm(x, (c,), p[name + ".ln1.g"], p[name + ".ln1.b"], 1e-5)
    qkv = conv(p, name + ".qkv", h, 3 * inner)
    dh = inner // heads
    ctx = torch.zeros(x.shape[0], inner)
    for hd in range(heads):
        sl = slice(hd * dh, (hd + 1) * dh)
        att = torch.softmax(qkv[:, sl] @ qkv[:, inner + hd * dh:inner + (hd + 1) * dh].t() / math.sqrt(dh), dim=-1)
        ctx[:, sl] = att @ qkv[:, 2 * inner + hd * dh:2 * inner + (hd + 1) * dh]
    x = x + conv(p, name + ".out", ctx, c)
    h = F.layer_norm(x, (c,), p[name + ".ln3.g"], p[name + ".ln3.b"], 1e-5)
    y = conv(p, name + ".ff1", h, 4 * c)
    y = y + p[name + ".snake.ib"] * torch.sin(y * p[name + ".snake.a"]) ** 2
    return x + conv(p, name + ".ff2", y, c)


def emul_matcha(p, cfg, text, z, temb, dts):
    """text (T_text,), z (T, odim) noise already scaled, temb [steps, blocks, C], dts [steps] -> dict(feat_gen, duration)"""
    d, od, c = cfg["adim"], cfg["odim"], cfg["decoder_channels"][0]
    heads, inner = cfg["decoder_num_heads"], cfg["decoder_num_heads"] * cfg["decoder_attention_head_dim"]
    nb, nm = cfg["decoder_n_blocks"], cfg["decoder_num_mid_blocks"]
    hs = conformer(p, "enc", p["emb"][text] * math.sqrt(d), cfg["elayers"], cfg["aheads"], cfg["eunits"])
    logd = predictor(p, "dur", hs, cfg["duration_predictor_layers"], cfg["duration_predictor_chans"])
    dur = torch.clamp(torch.round(logd.exp() - 1.0), min=0).long()
    dl = dur if int(dur.sum()) > 0 else torch.ones_like(dur)
    cum = torch.cumsum(dl, 0)
    olen = int(cum[-1]) - int(cum[-1]) % 2
    idx = torch.searchsorted(cum, torch.arange(olen), right=True)
    mu = conv(p, "enc_proj", hs[idx], od)
    x = z[:olen].clone()

    def tr(r, h):
        for j in range(nb):
            h = transformer(p, f"dec.tr{r}_{j}", h, heads, inner)
        return h

    for step in range(temb.shape[0]):
        te = temb[step]
        r = 0
        h0 = tr(r, resnet(p, r, torch.cat([x, mu], 1), te[r], c))                      # down 0 (skip 0)
        pairs = h0.reshape(olen // 2, 2 * c)                                           # two frames per row
        h = conv_off(p, "dec.down0", pairs, c, -1)                                     # stride-2 conv as a 2-tap conv
        r += 1
        h1 = tr(r, resnet(p, r, h, te[r], c))                                          # down 1 (skip 1)
        h = conv(p, "dec.down1", h1, c)
        r += 1
        for _ in range(nm):
            h = tr(r, resnet(p, r, h, te[r], c))
            r += 1
        h = tr(r, resnet(p, r, torch.cat([h, h1], 1), te[r], c))                       # up 0
        h = conv(p, "dec.up0", h, 2 * c).reshape(olen, c)                              # ConvTranspose1d as a 3-tap conv
        r += 1
        h = tr(r, resnet(p, r, torch.cat([h, h0], 1), te[r], c))                       # up 1
        h = conv(p, "dec.up1", h, c)
        h = gn_mish(p, "dec.final.gn", conv(p, "dec.final", h, c))
        x = x + dts[step] * conv(p, "dec.proj", h, od)
    return dict(feat_gen=x, duration=dur)
