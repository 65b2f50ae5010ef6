#!/usr/bin/env python3
"""bench.py -- synthesized audio-seconds per second of the JATTS batched synthesis path on B200.

Workload (BASELINE.json configs[1]): FastSpeech2 (JSUT tts1 config) + HiFi-GAN V1 (hop 300, 24 kHz),
random-init seeded weights (oracle/recipes.py, duration recipe A), batch of 64 synthetic 50-phoneme
utterances (~300 mel frames each).  One "step" = text -> waveform for the whole batch through the
public API (``FastSpeech2.inference_batch`` -> ``Vocoder.decode_batch``).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference ...                     # the reference's CPU arithmetic (oracle port)
  python bench.py --workload matcha64 ...                  # BASELINE config 5: Matcha-TTS (10 Euler steps) + HiFi-GAN, weak scaling
  python bench.py --workload voc10k ...                    # BASELINE config 4: 10 k mel clips, vocoder only, STRONG scaling

Under torchrun (N > 1) every rank owns one GPU and its own 64-utterance batch (weak scaling, no
collective on the data path; with ``--workload voc10k`` the fixed 10 k-clip list is sharded over the ranks
instead); rank 0 prints ONE JSON line.  After the timed region rank 0 checks two rows of the timed batch
against the CPU oracle ("parity_checked") and reports the bandwidth-bound kernels against the measured HBM
peak ("bandwidth").  See DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 64
T_TEXT = 50
METRIC = "synthesized_audio_seconds_per_second"
UNIT = "audio-s/s"


# --------------------------------------------------------------------------------------------
# algorithmic work (2 x MAC of the reference's dense ops; padding, halo and split passes excluded)
# --------------------------------------------------------------------------------------------
def hifigan_flops_per_frame(cfg) -> float:
    ch, k = cfg["channels"], cfg["kernel_size"]
    mac = cfg["in_channels"] * ch * k
    rate = 1
    for i, s in enumerate(cfg["upsample_scales"]):
        ci, co = ch >> i, ch >> (i + 1)
        rate *= s
        mac += rate * ci * co * 2  # transposed conv: 2 taps per output sample
        for rk, dils in zip(cfg["resblock_kernel_sizes"], cfg["resblock_dilations"]):
            mac += rate * co * co * rk * 2 * len(dils)
    mac += rate * (ch >> len(cfg["upsample_scales"])) * cfg["out_channels"] * k
    return 2.0 * mac


def fs2_flops(cfg, t_text: int, t_feats: int) -> float:
    d, k = cfg["adim"], cfg["positionwise_conv_kernel_size"]

    def conformer(t, units, n_layers):
        per_tok = 2 * (d * units * k * 2) + 4 * d * d + d * d + 2 * d * d + d * d  # 2 FFN, qkv+out, pos, pw1, pw2
        attn = 3 * t * d  # QK^T, BD, PV per token
        return n_layers * t * (per_tok + attn)

    mac = conformer(t_text, cfg["eunits"], cfg["elayers"]) + conformer(t_feats, cfg["dunits"], cfg["dlayers"])
    for name in ("duration", "pitch", "energy"):
        nl, c, kk = cfg[f"{name}_predictor_layers"], cfg[f"{name}_predictor_chans"], cfg[f"{name}_predictor_kernel_size"]
        mac += t_text * (d * c * kk + (nl - 1) * c * c * kk + c)
    mac += t_feats * d * cfg["odim"]
    pc, od, pk, pl = cfg["postnet_chans"], cfg["odim"], cfg["postnet_filts"], cfg["postnet_layers"]
    mac += t_feats * pk * (od * pc * 2 + (pl - 2) * pc * pc) if pl >= 2 else t_feats * pk * od * od
    return 2.0 * mac


# --------------------------------------------------------------------------------------------
# clocks sampled DURING the timed region (B200_PROFILING.md recipe)
# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        """index of the next sample line: brackets the timed region"""
        return len(self.lines)

    def stop(self, first=0, last=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        window = self.lines[first:last] or self.lines[-3:]   # samples taken while the timed steps ran
        for ln in window:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# CPU arm: the reference's arithmetic (oracle port) on the host cores
# --------------------------------------------------------------------------------------------
def cpu_synthesize(sd_fs2, cfg_fs2, sd_hg, cfg_hg, texts, keep=None):
    from oracle import fs2 as ofs2
    from oracle import hifigan as ohg

    frames = 0
    for x in texts:  # per-utterance loop exactly as jatts/bin/tts_decode.py:203-255 (no batching exists)
        out = ofs2.fs2_inference(sd_fs2, cfg_fs2, x)
        wav = ohg.hifigan_forward(sd_hg, cfg_hg, out["feat_gen"])
        frames += out["feat_gen"].shape[0]
        if keep is not None:
            keep.append((out, wav.reshape(-1)))
    return frames


def time_cpu(n_utt: int, reps: int, warm: int, seed0: int = 0, keep=None):
    """oracle port on the host cores; ``keep`` (a list) receives (fs2 outputs, waveform) of the first repetition"""
    from oracle import recipes

    cfg_fs2, cfg_hg = recipes.JSUT_FS2, recipes.HIFIGAN_V1_HOP300
    sd_fs2 = recipes.make_fs2_state_dict(cfg_fs2, seed=0, duration_recipe="A")
    sd_hg = recipes.make_hifigan_state_dict(cfg_hg, seed=0)
    texts = [recipes.make_phonemes(T_TEXT, seed0 + i, cfg_fs2["idim"]) for i in range(n_utt)]
    for _ in range(warm):
        cpu_synthesize(sd_fs2, cfg_fs2, sd_hg, cfg_hg, texts[:1])
    times, frames = [], 0
    for r in range(reps):
        t0 = time.perf_counter()
        frames = cpu_synthesize(sd_fs2, cfg_fs2, sd_hg, cfg_hg, texts, keep if r == 0 else None)
        times.append(time.perf_counter() - t0)
    audio_s = frames * recipes.HOP_SIZE / recipes.SAMPLING_RATE
    return audio_s, times


def matcha_decoder_flops(cfg, frames) -> float:
    """algorithmic FLOP (2 per multiply-add, each product counted ONCE although the split GEMM issues three MMAs) of one
    Euler step of the flow-matching U-Net (jatts/modules/matchatts/decoder.py:413-487) for utterances of ``frames`` frames"""
    c, od = cfg["decoder_channels"][0], cfg["odim"]
    inner = cfg["decoder_num_heads"] * cfg["decoder_attention_head_dim"]
    nb, nm = cfg["decoder_n_blocks"], cfg["decoder_num_mid_blocks"]
    total = 0.0
    for t in frames:
        t -= t % 2
        th = t // 2
        res = lambda rows, cin: 2.0 * rows * (3 * cin * c + 3 * c * c + cin * c)
        tr = lambda rows: nb * (2.0 * rows * (c * 3 * inner + inner * c + c * 4 * c + 4 * c * c) + 4.0 * rows * rows * inner)
        total += res(t, 2 * od) + tr(t) + 2.0 * th * 3 * c * c                       # down 0 + stride-2 conv
        total += res(th, c) + tr(th) + 2.0 * th * 3 * c * c                          # down 1 + conv
        total += nm * (res(th, c) + tr(th))                                          # mid
        total += res(th, 2 * c) + tr(th) + 2.0 * th * 4 * c * c                      # up 0 + ConvTranspose1d(4, 2, 1)
        total += res(t, 2 * c) + tr(t) + 2.0 * t * 3 * c * c                         # up 1 + conv
        total += 2.0 * t * 3 * c * c + 2.0 * t * c * od                              # final block + projection
    return total


def time_cpu_matcha(n_utt: int, reps: int, warm: int, keep=None):
    """config 5 on the host cores: oracle port of MatchaTTS.inference + the restated vocoder, per-utterance loop"""
    from oracle import hifigan as ohg
    from oracle import matcha as om
    from oracle import recipes

    cfg, cfg_hg = recipes.JSUT_MATCHA, recipes.HIFIGAN_V1_HOP300
    sd = recipes.make_matcha_state_dict(cfg, seed=0, duration_recipe="A")
    sd_hg = recipes.make_hifigan_state_dict(cfg_hg, seed=0)
    texts = [recipes.make_phonemes(T_TEXT, i, cfg["idim"]) for i in range(n_utt)]

    def one(i, x):
        z = recipes.make_noise(1024, cfg["odim"], i).t()
        out = om.matcha_inference(sd, cfg, x, z, recipes.MATCHA_ODE_STEPS, recipes.MATCHA_TEMPERATURE)
        return out, ohg.hifigan_forward(sd_hg, cfg_hg, out["feat_gen"]).reshape(-1)

    for _ in range(warm):
        one(0, texts[0])
    times, frames = [], 0
    for r in range(reps):
        t0 = time.perf_counter()
        frames = 0
        for i, x in enumerate(texts):
            out, wav = one(i, x)
            frames += out["feat_gen"].shape[0]
            if keep is not None and r == 0:
                keep.append((out, wav))
        times.append(time.perf_counter() - t0)
    return frames * recipes.HOP_SIZE / recipes.SAMPLING_RATE, times


def time_cpu_vocoder(n_clips: int, reps: int, warm: int):
    """config 4 on the host cores: the restated generator on the first clips of the 10 k list"""
    from oracle import hifigan as ohg
    from oracle import recipes

    cfg_hg = recipes.HIFIGAN_V1_HOP300
    sd_hg = recipes.make_hifigan_state_dict(cfg_hg, seed=0)
    lens = voc10k_lengths()
    mels = [recipes.make_mel(lens[i], i) for i in range(n_clips)]
    for _ in range(warm):
        ohg.hifigan_forward(sd_hg, cfg_hg, mels[0])
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        for m in mels:
            ohg.hifigan_forward(sd_hg, cfg_hg, m)
        times.append(time.perf_counter() - t0)
    return sum(lens[:n_clips]) * recipes.HOP_SIZE / recipes.SAMPLING_RATE, times


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if args.workload == "voc10k":
        n = 4
        audio_s, times = time_cpu_vocoder(n, reps=args.steps, warm=min(args.warmup, 1))
        sample = (f"{n} of the 10000 clips per step, per-clip loop as vocoder.py:56-67; restated HiFi-GAN V1 generator "
                  f"(oracle port, fp32 torch CPU)")
        extra = {"sample_clips_per_step": n}
    elif args.workload == "matcha64":
        n = 1
        audio_s, times = time_cpu_matcha(n, reps=args.steps, warm=min(args.warmup, 1))
        sample = (f"{n} of the {BATCH} utterances per step, per-utterance loop as tts_decode.py; oracle port of MatchaTTS.inference "
                  f"(10 Euler steps) + restated HiFi-GAN V1 (fp32 torch CPU)")
        extra = {"sample_utterances_per_step": n}
    else:
        n = 2  # bounded sample of the 64-utterance step
        audio_s, times = time_cpu(n, reps=args.steps, warm=min(args.warmup, 2))
        sample = (f"{n} of the {BATCH} utterances per step, per-utterance loop as tts_decode.py, "
                  f"oracle port of the reference arithmetic (fp32 torch CPU)")
        extra = {"sample_utterances_per_step": n}
    total = sum(times)
    value = audio_s * len(times) / total
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "strong" if args.workload == "voc10k" else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, world, extra),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample,
                         "host_cpus": cores},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def voc10k_lengths():
    g = torch.Generator().manual_seed(4)
    return torch.randint(200, 401, (10000,), generator=g).tolist()   # T ~ U{200..400}


class Timer:
    def __init__(self, dev, world, flush):
        self.dev, self.world, self.flush = dev, world, flush

    def barrier(self):
        if self.world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(self.dev)

    def __call__(self, fn, steps):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        self.barrier()
        for a, b in ev:
            self.flush.fill_(1)  # L2 flush between timed iterations (outside the event pair)
            a.record()
            fn()
            b.record()
        self.barrier()
        return sum(a.elapsed_time(b) for a, b in ev)  # ms

    def region(self, fn, steps, finish):
        """ONE event pair around `steps` back-to-back calls (+ `finish`, which joins the copy stream): the end-to-end leg,
        where step i's device->host copy runs under step i+1 and every copy completes inside the region"""
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        a.record()
        for _ in range(steps):
            fn()
        finish()
        b.record()
        self.barrier()
        return a.elapsed_time(b)


class HostSink:
    """Double-buffered pinned host memory fed by a copy stream: what jatts_b200/decode.py does with its PinnedRing --
    the waveform of step i goes to the host while step i+1 computes."""

    def __init__(self, dev):
        self.dev, self.stream, self.bufs, self.i = dev, torch.cuda.Stream(device=dev), [None, None], 0

    def push(self, flat):
        k = self.i & 1
        self.i += 1
        if self.bufs[k] is None or self.bufs[k].numel() != flat.numel() or self.bufs[k].dtype != flat.dtype:
            self.bufs[k] = torch.empty(flat.numel(), dtype=flat.dtype).pin_memory()
        self.stream.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(self.stream):
            self.bufs[k].copy_(flat, non_blocking=True)   # in-order on the copy stream: buffer k's previous copy has finished
        flat.record_stream(self.stream)

    def finish(self):
        torch.cuda.current_stream(self.dev).wait_stream(self.stream)


def start_sampler(rank, local_rank, dev, warm_fn):
    sampler = ClockSampler(local_rank if "CUDA_VISIBLE_DEVICES" not in os.environ else
                           int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank]))
    if rank == 0:
        # nvidia-smi needs a moment to produce its first line: keep the GPU under the same load until it does
        sampler.start()
        t_wait = time.time()
        while sampler.mark() == 0 and time.time() - t_wait < 5.0:
            warm_fn()
            torch.cuda.synchronize(dev)
    return sampler


def reduce_times(world, dev, times_ms, audio_s):
    t = torch.tensor(times_ms, dtype=torch.float64, device=dev)
    a = torch.tensor([audio_s], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)   # max over ranks
        torch.distributed.all_reduce(a, op=torch.distributed.ReduceOp.SUM)   # whole-job audio
    return [float(x) for x in t], float(a[0])


def hifigan_roofline(_lib, voc_fn, frames, cfg_hg, seconds=1.0):
    """Per-launch CUDA-event timing of every HiFi-GAN convolution launch, repeated back to back for ~`seconds` so the
    power-capped (sustained) clock applies; returns (ms per decode of the conv family, launches, ms of output conv)."""
    reps, t0 = 0, time.time()
    ms_conv = ms_out = 0.0
    n_conv = 0
    while reps < 2 or (time.time() - t0 < seconds and reps < 200):
        _lib.profile_begin()
        voc_fn()
        cls = _lib.profile_end_classes()
        ms_conv += cls["bf16_conv"][0]
        n_conv = cls["bf16_conv"][1]
        ms_out += cls["output_conv"][0]
        reps += 1
    return ms_conv / reps, int(n_conv), ms_out / reps, reps


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            j = json.load(f)
        return {"sustained": float(j["bf16_tflops_sustained"]), "burst": float(j["bf16_tflops"]), "hbm_gbs": float(j["hbm_gbs"]),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"sustained": 1400.0, "burst": 1650.0, "hbm_gbs": 6550.0, "source": "fallback (B200_PROFILING.md)"}


def measured_traffic():
    """DRAM bytes (read + write) of the dominant kernel family per step from the newest committed ncu capture
    (profiles/r0N_traffic.json, written by tools/step_metrics.py); None when absent.  NOT measured in this run."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.isfile(p):
            with open(p) as f:
                j = json.load(f)
            j["file"] = "profiles/" + name
            return j
    return None


def workload_config(workload: str, world: int, extra: dict) -> dict:
    if workload == "voc10k":
        cfg = {"workload": "HiFi-GAN V1 hop 300 vocoder-only bulk synthesis of 10000 synthetic 80-bin mel clips (T ~ U{200..400} "
                           "frames, N(0,1), seed = clip index), utterance-sharded over the GPUs (BASELINE config 4), seeded "
                           "random-init weights", "clips": 10000, "clips_per_launch": VOC_BATCH,
               "parallelism": f"clips sharded by greedy LPT over {world} rank(s), no collective on the data path"}
    elif workload == "matcha64":
        cfg = {"workload": "Matcha-TTS (JSUT tts1 matcha_tts.v1.prior.steplr.large: 512-wide U-Net, 2 heads x 256, predicted durations) "
                           "with 10 Euler steps at temperature 0.667 + HiFi-GAN V1 hop 300, batch 64 x 50 phonemes (~300 frames each) "
                           "per GPU (BASELINE config 5), seeded random-init weights (duration recipe A)",
               "batch_per_gpu": BATCH, "t_text": T_TEXT, "ode_steps": 10, "parallelism": f"utterance-sharded replicas x{world}"}
    else:
        cfg = {"workload": "FastSpeech2 (JSUT tts1) + HiFi-GAN V1 hop 300, batch 64 x 50 phonemes (~300 frames each) per GPU, "
                           "seeded random-init weights (duration recipe A)",
               "batch_per_gpu": BATCH, "t_text": T_TEXT, "parallelism": f"utterance-sharded replicas x{world}"}
    cfg.update(extra)
    return cfg


VOC_BATCH = 128
E2E_TIMING = ("one CUDA-event pair around the K back-to-back steps (barrier + synchronize on both sides, max over ranks): every step copies "
              "its tokens from pinned host memory and its waveform to pinned host memory; the device->host copy of step i runs on a copy "
              "stream under step i+1 (double-buffered, as jatts_b200/decode.py does) and the last copy completes inside the region")


def build_models(dev):
    import jatts_b200
    from oracle import recipes  # weights / inputs only (seeded synthetic recipes); no oracle compute here

    cfg_fs2, cfg_hg = recipes.JSUT_FS2, recipes.HIFIGAN_V1_HOP300
    model = jatts_b200.FastSpeech2(**cfg_fs2)
    model.load_state_dict(recipes.make_fs2_state_dict(cfg_fs2, seed=0, duration_recipe="A"))
    model = model.eval().to(dev)
    stats = {"mean": torch.zeros(80), "scale": torch.ones(80)}
    voc = jatts_b200.Vocoder(recipes.make_hifigan_state_dict(cfg_hg, seed=0),
                             {"generator_type": "HiFiGANGenerator", "generator_params": dict(cfg_hg),
                              "sampling_rate": recipes.SAMPLING_RATE}, stats, dev, trg_stats=stats)
    return model, voc, cfg_fs2, cfg_hg


def run_b200(args, rank: int, world: int, local_rank: int):
    from jatts_b200 import _lib
    from oracle import recipes

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    model, voc, cfg_fs2, cfg_hg = build_models(dev)
    texts_cpu = [recipes.make_phonemes(T_TEXT, 1000 * rank + i, cfg_fs2["idim"]) for i in range(BATCH)]
    tok_host = torch.cat(texts_cpu).pin_memory()
    texts_dev = [t.to(dev) for t in texts_cpu]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    timed = Timer(dev, world, flush)

    def step_device():
        outs = model.inference_batch(texts_dev)
        waves = voc.decode_batch([o["feat_gen"] for o in outs])
        return outs, waves

    sink = HostSink(dev)

    def step_e2e():
        tok = tok_host.to(dev, non_blocking=True)
        outs = model.inference_batch(list(tok.split(T_TEXT)))
        waves = voc.decode_batch([o["feat_gen"] for o in outs])
        flat = torch.cat(waves)
        sink.push(flat)
        return outs, flat

    for _ in range(max(args.warmup, 3)):
        outs, waves = step_device()
    torch.cuda.synchronize(dev)
    frames = sum(int(o["feat_gen"].shape[0]) for o in outs)
    audio_s = frames * recipes.HOP_SIZE / recipes.SAMPLING_RATE

    sampler = start_sampler(rank, local_rank, dev, step_device)
    l0 = _lib.launch_count()
    s0 = sampler.mark()
    ms_dev = timed(step_device, args.steps)
    launches = (_lib.launch_count() - l0) // args.steps
    for _ in range(2):
        step_e2e()
    ms_e2e = timed.region(step_e2e, args.steps, sink.finish)
    clocks = sampler.stop(s0, sampler.mark()) if rank == 0 else None
    # a longer back-to-back region (no L2 flush, >= 3 s) for the power-capped regime: reported next to the K-step number
    sus_steps, t0 = 0, time.time()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    a.record()
    while time.time() - t0 < 3.0:
        step_device()
        sus_steps += 1
    b.record()
    torch.cuda.synchronize(dev)
    ms_sus = a.elapsed_time(b) / sus_steps

    # roofline leg: device time of the dominant kernel family and of the bandwidth-bound kernels (per-launch events)
    mels = [o["feat_gen"] for o in outs]
    torch.cuda.synchronize(dev)
    ms_conv, n_conv, ms_outconv, roof_reps = hifigan_roofline(_lib, lambda: voc.decode_batch(mels), frames, cfg_hg)
    _lib.profile_begin()
    model.inference_batch(texts_dev)
    fs2_cls = _lib.profile_end_classes()

    (ms_dev, ms_e2e), total_audio = reduce_times(world, dev, [ms_dev, ms_e2e], audio_s)
    if rank != 0:
        return
    pk = peaks()
    traffic = measured_traffic()
    hg_flops = hifigan_flops_per_frame(cfg_hg) * frames
    conv_share = 1.0 - 2.0 * (cfg_hg["channels"] >> 4) * 7 * 300 / hifigan_flops_per_frame(cfg_hg)  # minus output_conv
    achieved = hg_flops * conv_share / (ms_conv * 1e-3) / 1e12
    fs2_fl = sum(fs2_flops(cfg_fs2, T_TEXT, int(o["feat_gen"].shape[0])) for o in outs)
    # ---- bandwidth-bound kernels: algorithmic bytes (one read of each input + one write of each output at the stored
    #      dtype) / summed event time / measured HBM copy bandwidth
    d = cfg_fs2["adim"]
    t_rows, f_rows = BATCH * T_TEXT, frames
    n_samp = frames * recipes.HOP_SIZE
    el, dl = cfg_fs2["elayers"], cfg_fs2["dlayers"]
    ln_bytes = (5 * el * t_rows + 5 * dl * f_rows) * d * 8 + (t_rows + f_rows) * d * 8
    pred = [(cfg_fs2[f"{n}_predictor_layers"], cfg_fs2[f"{n}_predictor_chans"]) for n in ("duration", "pitch", "energy")]
    ln_bytes += sum(t_rows * c * 8 * (nl - 1) + t_rows * c * 4 for nl, c in pred)
    bw_alg = {"layernorm": ln_bytes, "dwconv_swish": (el * t_rows + dl * f_rows) * d * 8,
              "length_regulate": (t_rows + f_rows) * d * 4 + f_rows * 4, "output_conv": n_samp * (32 * 2 + 4)}
    bw_ms = {k: fs2_cls[k][0] for k in ("layernorm", "dwconv_swish", "length_regulate")}
    bw_ms["output_conv"] = ms_outconv
    bandwidth = {k: {"algorithmic_bytes": int(bw_alg[k]), "ms": bw_ms[k], "launches": int(fs2_cls[k][1]) if k in fs2_cls and k != "output_conv" else 1,
                     "achieved_gbs": bw_alg[k] / (bw_ms[k] * 1e-3) / 1e9 if bw_ms[k] > 0 else None,
                     "frac_of_hbm_peak": bw_alg[k] / (bw_ms[k] * 1e-3) / 1e9 / pk["hbm_gbs"] if bw_ms[k] > 0 else None}
                 for k in bw_alg}
    # ---- CPU baseline (oracle port) on the first utterances of the timed batch; its outputs double as the parity check
    cpu_cores = os.cpu_count() or 1
    torch.set_num_threads(cpu_cores)
    kept = []
    cpu_audio, cpu_times = time_cpu(4, reps=20, warm=1, seed0=0, keep=kept)   # ~10 s of CPU work
    cpu_value = cpu_audio * len(cpu_times) / sum(cpu_times)
    # the recipe's own setting is ONE thread (egs/*/tts1/path.sh:15 exports OMP_NUM_THREADS=1): one utterance, 3 repetitions
    torch.set_num_threads(1)
    st_audio, st_times = time_cpu(1, reps=3, warm=1, seed0=0)
    torch.set_num_threads(cpu_cores)
    med = lambda v: sorted(v)[len(v) // 2]
    from oracle import hifigan as ohg
    parity = {"rows": 2, "durations_equal": True, "mel_max_abs": 0.0, "wave_ac_snr_db": 1e9}
    for i in range(2):   # rows 0 and 1 of the TIMED batch against the oracle
        ref, wref = kept[i]
        parity["durations_equal"] &= bool(torch.equal(ref["duration"], outs[i]["duration"].cpu()))
        if ref["feat_gen"].shape == outs[i]["feat_gen"].shape:
            parity["mel_max_abs"] = max(parity["mel_max_abs"], float((ref["feat_gen"] - outs[i]["feat_gen"].cpu()).abs().max()))
            parity["wave_ac_snr_db"] = min(parity["wave_ac_snr_db"], ohg.ac_snr_db(wref, waves[i].cpu().reshape(-1)))
        else:
            parity["durations_equal"] = False
    parity_ok = parity["durations_equal"] and parity["mel_max_abs"] < 1e-3 and parity["wave_ac_snr_db"] >= 35.0
    line = {
        "metric": METRIC, "value": total_audio * args.steps / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config("tts64", world, {"mel_frames_per_gpu": frames,
                   "audio_seconds_per_step_per_gpu": audio_s, "sampling_rate": recipes.SAMPLING_RATE,
                   "hop_size": recipes.HOP_SIZE,
                   "l2": "256 MB buffer written between timed steps; per-step activation working set ~2 GB >> 126 MB L2",
                   "precision": "HiFi-GAN: bf16 operands / fp32 TMEM accumulate; FastSpeech2 GEMMs and attention: fp16 hi+lo split (3 MMA) / fp32"}),
        "e2e": {"value": total_audio * args.steps / (ms_e2e * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": int(tok_host.numel() * 8), "d2h_bytes_per_step": int(frames * recipes.HOP_SIZE * 4),
                "ms_per_step": ms_e2e / args.steps, "timing": E2E_TIMING},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "sustained": {"seconds": 1e-3 * ms_sus * sus_steps, "steps": sus_steps, "ms_per_step": ms_sus,
                      "value_per_gpu": audio_s / (ms_sus * 1e-3),
                      "note": "back-to-back steps for >= 3 s on rank 0 (power-capped clocks), no L2 flush; the K-step number above is the contract value"},
        "parity_checked": bool(parity_ok), "parity": parity,
        "roofline": {"bound": "tensor",
                     "kernel": "HiFi-GAN convolution kernels: mrf_pair_kernel<C,k> (fused residual units, C = 32/64) + "
                               "conv_bf16_tma_kernel<*> (all other Conv1d / ConvTranspose1d launches of one step)",
                     "achieved": achieved, "peak": pk["sustained"], "unit": "TFLOP/s", "frac": achieved / pk["sustained"],
                     "frac_of_burst_peak": achieved / pk["burst"], "peak_burst": pk["burst"],
                     "peak_source": pk["source"] + ": bf16_tflops_sustained (the kernels are timed inside ~1 s of back-to-back decodes)",
                     "timed_decodes": roof_reps,
                     "traffic": (traffic or {}).get("hifigan_conv_dram_bytes_per_step"),
                     "traffic_source": ("committed ncu capture " + traffic["file"] + " (not measured in this run): " + str(traffic.get("source")))
                                       if traffic else None,
                     "launches_per_step": n_conv, "kernel_ms_per_step": ms_conv,
                     "algorithmic_flop_per_step": hg_flops * conv_share,
                     "fs2_split_gemm": {"launches_per_step": int(fs2_cls["split_gemm"][1]), "kernel_ms_per_step": fs2_cls["split_gemm"][0],
                                        "algorithmic_flop_per_step": fs2_fl},
                     "fs2_attention": {"launches_per_step": int(fs2_cls["attention"][1]), "kernel_ms_per_step": fs2_cls["attention"][0]}},
        "bandwidth": bandwidth,
        "cpu_baseline": {"value": cpu_value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "host_cpus": cpu_cores,
                         "sample": "4 of the 64 utterances (50 phonemes each) x 20 repetitions after 1 warm-up, per-utterance loop as "
                                   "tts_decode.py; oracle port of the reference arithmetic, fp32 torch CPU, all host cores "
                                   "(torch intra-op threads = os.cpu_count())",
                         "best": cpu_audio / min(cpu_times), "median": cpu_audio / med(cpu_times),
                         "single_thread": {"value": st_audio * len(st_times) / sum(st_times), "best": st_audio / min(st_times),
                                           "median": st_audio / med(st_times), "cores": 1,
                                           "sample": "1 utterance x 3 repetitions after 1 warm-up with torch.set_num_threads(1): the "
                                                     "recipe's own OMP_NUM_THREADS=1 (egs/*/tts1/path.sh:15)"}},
    }
    print(json.dumps(line), flush=True)


def run_matcha64(args, rank: int, world: int, local_rank: int):
    """BASELINE config 5: Matcha-TTS text2mel (10 Euler steps) + HiFi-GAN, batch 64 per GPU, weak scaling."""
    import jatts_b200
    from jatts_b200 import _lib
    from oracle import recipes

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    cfg, cfg_hg = recipes.JSUT_MATCHA, recipes.HIFIGAN_V1_HOP300
    steps_ode, temp = recipes.MATCHA_ODE_STEPS, recipes.MATCHA_TEMPERATURE
    model = jatts_b200.MatchaTTS(**cfg)
    model.load_state_dict(recipes.make_matcha_state_dict(cfg, seed=0, duration_recipe="A"))
    model = model.eval().to(dev)
    stats = {"mean": torch.zeros(80), "scale": torch.ones(80)}
    voc = jatts_b200.Vocoder(recipes.make_hifigan_state_dict(cfg_hg, seed=0),
                             {"generator_type": "HiFiGANGenerator", "generator_params": dict(cfg_hg),
                              "sampling_rate": recipes.SAMPLING_RATE}, stats, dev, trg_stats=stats)
    # rank 0's first utterances are the ones the CPU arm synthesises (seeds 0, 1, ...): the parity check below
    texts_cpu = [recipes.make_phonemes(T_TEXT, 1000 * rank + i, cfg["idim"]) for i in range(BATCH)]
    tok_host = torch.cat(texts_cpu).pin_memory()
    texts_dev = [t.to(dev) for t in texts_cpu]
    noise_dev = {}

    def fixed_noise(frames):   # the seeded z of oracle/recipes.py, resident on the device after the first call
        key = tuple(frames)
        if key not in noise_dev:
            noise_dev[key] = [recipes.make_noise(1024, cfg["odim"], i)[:f].to(dev) for i, f in enumerate(frames)]
        return noise_dev[key]

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    timed = Timer(dev, world, flush)

    def step_device():
        outs = model.inference_batch(texts_dev, n_timesteps=steps_ode, temperature=temp, noise=fixed_noise)
        waves = voc.decode_batch([o["feat_gen"] for o in outs])
        return outs, waves

    sink = HostSink(dev)

    def step_e2e():   # the call a user makes: host tokens in, noise drawn on the device, host waveform out
        tok = tok_host.to(dev, non_blocking=True)
        outs = model.inference_batch(list(tok.split(T_TEXT)), n_timesteps=steps_ode, temperature=temp)
        flat = torch.cat(voc.decode_batch([o["feat_gen"] for o in outs]))
        sink.push(flat)
        return outs, flat

    for _ in range(max(args.warmup, 3)):
        outs, waves = step_device()
    torch.cuda.synchronize(dev)
    frames_l = [int(o["feat_gen"].shape[0]) for o in outs]
    frames = sum(frames_l)
    audio_s = frames * recipes.HOP_SIZE / recipes.SAMPLING_RATE
    sampler = start_sampler(rank, local_rank, dev, step_device)
    l0 = _lib.launch_count()
    s0 = sampler.mark()
    ms_dev = timed(step_device, args.steps)
    launches = (_lib.launch_count() - l0) // args.steps
    for _ in range(2):
        step_e2e()
    ms_e2e = timed.region(step_e2e, args.steps, sink.finish)
    clocks = sampler.stop(s0, sampler.mark()) if rank == 0 else None
    # per-class device time of one text2mel call and one vocoder call (per-launch CUDA events)
    _lib.profile_begin()
    model.inference_batch(texts_dev, n_timesteps=steps_ode, temperature=temp, noise=fixed_noise)
    t2m = _lib.profile_end_classes()
    mels = [o["feat_gen"] for o in outs]
    ms_conv, n_conv, _, _ = hifigan_roofline(_lib, lambda: voc.decode_batch(mels), frames, cfg_hg, seconds=0.5)
    (ms_dev, ms_e2e), total_audio = reduce_times(world, dev, [ms_dev, ms_e2e], audio_s)
    if rank != 0:
        return
    pk = peaks()
    dec_flop = matcha_decoder_flops(cfg, frames_l) * steps_ode
    enc_flop = BATCH * fs2_flops(dict(recipes.JSUT_FS2, dlayers=0, pitch_predictor_chans=0, energy_predictor_chans=0), T_TEXT, 0)
    gemm_ms = t2m["split_gemm"][0]
    achieved = (dec_flop + enc_flop) / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
    cpu_cores = os.cpu_count() or 1
    torch.set_num_threads(cpu_cores)
    kept = []
    cpu_audio, cpu_times = time_cpu_matcha(2, reps=3, warm=1, keep=kept)
    cpu_value = cpu_audio * len(cpu_times) / sum(cpu_times)
    from oracle import hifigan as ohg
    parity = {"rows": 2, "durations_equal": True, "mel_max_abs": 0.0, "wave_ac_snr_db": 1e9}
    for i in range(2):   # rows 0 and 1 of the TIMED batch against the oracle (same tokens, same noise)
        ref, wref = kept[i]
        parity["durations_equal"] &= bool(torch.equal(ref["duration"], outs[i]["duration"].cpu()))
        if ref["feat_gen"].shape == outs[i]["feat_gen"].shape:
            parity["mel_max_abs"] = max(parity["mel_max_abs"], float((ref["feat_gen"] - outs[i]["feat_gen"].cpu()).abs().max()))
            parity["wave_ac_snr_db"] = min(parity["wave_ac_snr_db"], ohg.ac_snr_db(wref, waves[i].cpu().reshape(-1)))
        else:
            parity["durations_equal"] = False
    parity_ok = parity["durations_equal"] and parity["mel_max_abs"] < 1e-3 and parity["wave_ac_snr_db"] >= 35.0
    line = {
        "metric": METRIC, "value": total_audio * args.steps / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16x2-split / bf16", "data": "synthetic",
        "config": workload_config("matcha64", world, {"mel_frames_per_gpu": frames, "audio_seconds_per_step_per_gpu": audio_s,
                   "sampling_rate": recipes.SAMPLING_RATE, "hop_size": recipes.HOP_SIZE,
                   "l2": "256 MB buffer written between timed steps; per-step activation working set >> 126 MB L2",
                   "noise": "device-timed steps use the seeded z of oracle/recipes.py::make_noise resident on the device; the e2e "
                            "steps draw z with torch.randn on the device inside the call, as the reference does",
                   "precision": "Matcha text2mel: fp16 hi+lo split GEMMs and attention (3 MMA) / fp32; HiFi-GAN: bf16 operands / fp32 accumulate"}),
        "e2e": {"value": total_audio * args.steps / (ms_e2e * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": int(tok_host.numel() * 8), "d2h_bytes_per_step": int(frames * recipes.HOP_SIZE * 4),
                "ms_per_step": ms_e2e / args.steps, "timing": E2E_TIMING},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "parity_checked": bool(parity_ok), "parity": parity,
        "roofline": {"bound": "tensor",
                     "kernel": "gemm_split_tma_kernel<*>: every Conv1d / Linear / ConvTranspose1d of the Matcha encoder and of the 10 "
                               "evaluations of the flow-matching U-Net (fp16 hi+lo split operands: 3 tcgen05 MMAs per algorithmic product)",
                     "achieved": achieved, "peak": pk["sustained"], "unit": "TFLOP/s",
                     "frac": achieved / pk["sustained"] if achieved else None,
                     "frac_of_issued_mma": 3.0 * achieved / pk["sustained"] if achieved else None,
                     "peak_source": pk["source"] + ": bf16_tflops_sustained",
                     "traffic": None, "launches_per_step": int(t2m["split_gemm"][1]), "kernel_ms_per_step": gemm_ms,
                     "algorithmic_flop_per_step": dec_flop + enc_flop,
                     "note": "achieved counts each product once; the kernel issues 3 MMAs per product (frac_of_issued_mma)",
                     "matcha_attention": {"launches_per_step": int(t2m["attention"][1]), "kernel_ms_per_step": t2m["attention"][0]},
                     "matcha_norm_act": {"launches_per_step": int(t2m["layernorm"][1]), "kernel_ms_per_step": t2m["layernorm"][0],
                                         "what": "LayerNorm, GroupNorm statistics + Mish (SnakeBeta is an epilogue of its GEMM)"},
                     "hifigan_conv": {"launches_per_step": n_conv, "kernel_ms_per_step": ms_conv}},
        "cpu_baseline": {"value": cpu_value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "host_cpus": cpu_cores,
                         "sample": "2 of the 64 utterances (50 phonemes each) x 3 repetitions after 1 warm-up, per-utterance loop as "
                                   "tts_decode.py; oracle port of MatchaTTS.inference (10 Euler steps) + restated HiFi-GAN V1, fp32 torch "
                                   "CPU, all host cores"},
    }
    print(json.dumps(line), flush=True)


def run_voc10k(args, rank: int, world: int, local_rank: int):
    """BASELINE config 4: 10 k clips, STRONG scaling -- the clip list is fixed and sharded over the ranks."""
    import jatts_b200
    from jatts_b200 import _lib
    from oracle import recipes

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    cfg_hg = recipes.HIFIGAN_V1_HOP300
    gen = jatts_b200.HiFiGANGenerator(**cfg_hg)
    gen.load_state_dict(recipes.make_hifigan_state_dict(cfg_hg, seed=0))
    gen = gen.eval().to(dev)
    lens = voc10k_lengths()
    mine = jatts_b200.shard_utterances(lens, world)[rank]
    mine = sorted(mine, key=lambda i: (lens[i], i))                     # length-bucketed launches
    batches = [mine[s:s + VOC_BATCH] for s in range(0, len(mine), VOC_BATCH)]
    host = [torch.cat([recipes.make_mel(lens[i], i) for i in b]).pin_memory() for b in batches]
    blens = [[lens[i] for i in b] for b in batches]
    devm = [h.to(dev) for h in host]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    timed = Timer(dev, world, flush)
    frames = sum(lens[i] for i in mine)
    audio_s = frames * recipes.HOP_SIZE / recipes.SAMPLING_RATE
    n_samp = frames * recipes.HOP_SIZE
    wave_host = torch.empty(max(sum(bl) for bl in blens) * recipes.HOP_SIZE, dtype=torch.float32).pin_memory()

    def step_device():
        for m, bl in zip(devm, blens):
            gen.inference_batch(list(m.split(bl)))

    def step_e2e():
        for h, bl in zip(host, blens):
            m = h.to(dev, non_blocking=True)
            ys = gen.inference_batch(list(m.split(bl)))
            flat = torch.cat(ys).reshape(-1)
            wave_host[:flat.numel()].copy_(flat, non_blocking=True)

    for _ in range(max(args.warmup, 3) if len(batches) < 4 else 1):
        step_device()
    torch.cuda.synchronize(dev)
    sampler = start_sampler(rank, local_rank, dev, lambda: gen.inference_batch(list(devm[0].split(blens[0]))))
    l0 = _lib.launch_count()
    s0 = sampler.mark()
    ms_dev = timed(step_device, args.steps)
    launches = (_lib.launch_count() - l0) // args.steps
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop(s0, sampler.mark()) if rank == 0 else None
    _lib.profile_begin()
    step_device()
    cls = _lib.profile_end_classes()
    (ms_dev, ms_e2e), total_audio = reduce_times(world, dev, [ms_dev, ms_e2e], audio_s)
    t_all = torch.tensor([frames], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t_all, op=torch.distributed.ReduceOp.MAX)
    if rank != 0:
        return
    pk = peaks()
    hg_flops = hifigan_flops_per_frame(cfg_hg) * frames
    conv_share = 1.0 - 2.0 * (cfg_hg["channels"] >> 4) * 7 * 300 / hifigan_flops_per_frame(cfg_hg)
    achieved = hg_flops * conv_share / (cls["bf16_conv"][0] * 1e-3) / 1e12
    torch.set_num_threads(os.cpu_count() or 1)
    cpu_audio, cpu_times = time_cpu_vocoder(4, reps=3, warm=1)
    cpu_value = cpu_audio * len(cpu_times) / sum(cpu_times)
    line = {
        "metric": METRIC, "value": total_audio * args.steps / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config("voc10k", world, {"clips_rank0": len(mine), "mel_frames_rank0": frames,
                   "max_over_ranks_frames": float(t_all[0]), "total_audio_seconds": total_audio,
                   "l2": "256 MB buffer written between timed steps; one step = the rank's whole shard (>= 0.4 M frames)"}),
        "e2e": {"value": total_audio * args.steps / (ms_e2e * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": int(frames * 80 * 4), "d2h_bytes_per_step": int(n_samp * 4), "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "HiFi-GAN convolution kernels (mrf_pair_kernel + conv_bf16_tma_kernel), rank 0's shard",
                     "achieved": achieved, "peak": pk["sustained"], "unit": "TFLOP/s", "frac": achieved / pk["sustained"],
                     "frac_of_burst_peak": achieved / pk["burst"], "peak_source": pk["source"], "traffic": None,
                     "launches_per_step": int(cls["bf16_conv"][1]), "kernel_ms_per_step": cls["bf16_conv"][0],
                     "algorithmic_flop_per_step": hg_flops * conv_share},
        "cpu_baseline": {"value": cpu_value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": "the first 4 of the 10000 clips x 3 repetitions, restated generator (oracle port), fp32 torch CPU"},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="tts64", choices=["tts64", "voc10k", "matcha64"],
                    help="tts64 = BASELINE config 2 (the headline, weak scaling); voc10k = config 4 (strong scaling); "
                         "matcha64 = config 5 (Matcha-TTS + HiFi-GAN, weak scaling)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback); use --impl reference for the CPU arm")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        if args.workload == "voc10k":
            run_voc10k(args, rank, world, local_rank)
        elif args.workload == "matcha64":
            run_matcha64(args, rank, world, local_rank)
        else:
            run_b200(args, rank, world, local_rank)
    finally:
        if world > 1:
            torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
