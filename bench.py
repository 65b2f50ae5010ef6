#!/usr/bin/env python3
"""bench.py -- synthesized audio-seconds per second of the JATTS batched synthesis path on B200.

Workload (BASELINE.json configs[1]): FastSpeech2 (JSUT tts1 config) + HiFi-GAN V1 (hop 300, 24 kHz),
random-init seeded weights (oracle/recipes.py, duration recipe A), batch of 64 synthetic 50-phoneme
utterances (~300 mel frames each).  One "step" = text -> waveform for the whole batch through the
public API (``FastSpeech2.inference_batch`` -> ``Vocoder.decode_batch``).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference ...                     # the reference's CPU arithmetic (oracle port)

Under torchrun (N > 1) every rank owns one GPU and its own 64-utterance batch (weak scaling, no
collective on the data path); rank 0 prints ONE JSON line.  See DESIGN.md "Measurement".
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


def measured_traffic():
    """DRAM bytes (read + write) of the dominant kernel family per step, from the committed ncu capture
    (profiles/r01_traffic.json, written by tools/step_metrics.py); None when the file is absent."""
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.isfile(p):
        with open(p) as f:
            return json.load(f)
    return None


def workload_config(world: int, extra: dict) -> dict:
    cfg = {"workload": "FastSpeech2 (JSUT tts1) + HiFi-GAN V1 hop 300, batch 64 x 50 phonemes (~300 frames each) per GPU, "
                       "seeded random-init weights (duration recipe A)",
           "batch_per_gpu": BATCH, "t_text": T_TEXT, "parallelism": f"utterance-sharded replicas x{world}"}
    cfg.update(extra)
    return cfg


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            j = json.load(f)
        return float(j["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    return 1400.0, "fallback (B200_PROFILING.md sustained ~1.4 PFLOP/s)"


# --------------------------------------------------------------------------------------------
# CPU arm: the reference's arithmetic (oracle port) on the host cores
# --------------------------------------------------------------------------------------------
def cpu_synthesize(sd_fs2, cfg_fs2, sd_hg, cfg_hg, texts):
    from oracle import fs2 as ofs2
    from oracle import hifigan as ohg

    frames = 0
    for x in texts:  # per-utterance loop exactly as jatts/bin/tts_decode.py:203-255 (no batching exists)
        out = ofs2.fs2_inference(sd_fs2, cfg_fs2, x)
        ohg.hifigan_forward(sd_hg, cfg_hg, out["feat_gen"])
        frames += out["feat_gen"].shape[0]
    return frames


def time_cpu(n_utt: int, reps: int, warm: int, seed0: int = 0):
    from oracle import recipes

    cfg_fs2, cfg_hg = recipes.JSUT_FS2, recipes.HIFIGAN_V1_HOP300
    sd_fs2 = recipes.make_fs2_state_dict(cfg_fs2, seed=0, duration_recipe="A")
    sd_hg = recipes.make_hifigan_state_dict(cfg_hg, seed=0)
    texts = [recipes.make_phonemes(T_TEXT, seed0 + i, cfg_fs2["idim"]) for i in range(n_utt)]
    for _ in range(warm):
        cpu_synthesize(sd_fs2, cfg_fs2, sd_hg, cfg_hg, texts[:1])
    times, frames = [], 0
    for _ in range(reps):
        t0 = time.perf_counter()
        frames = cpu_synthesize(sd_fs2, cfg_fs2, sd_hg, cfg_hg, texts)
        times.append(time.perf_counter() - t0)
    audio_s = frames * recipes.HOP_SIZE / recipes.SAMPLING_RATE
    return audio_s, times


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_utt = 2  # bounded sample of the 64-utterance step
    audio_s, times = time_cpu(n_utt, reps=args.steps, warm=min(args.warmup, 2))
    total = sum(times)
    value = audio_s * len(times) / total
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(world, {"sample_utterances_per_step": n_utt}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{n_utt} of the {BATCH} utterances per step, per-utterance loop as tts_decode.py, "
                                   f"oracle port of the reference arithmetic (fp32 torch CPU)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def run_b200(args, rank: int, world: int, local_rank: int):
    import jatts_b200
    from jatts_b200 import _lib
    from oracle import recipes  # weights / inputs only (seeded synthetic recipes); no oracle compute here

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    cfg_fs2, cfg_hg = recipes.JSUT_FS2, recipes.HIFIGAN_V1_HOP300
    model = jatts_b200.FastSpeech2(**cfg_fs2)
    model.load_state_dict(recipes.make_fs2_state_dict(cfg_fs2, seed=0, duration_recipe="A"))
    model = model.eval().to(dev)
    stats = {"mean": torch.zeros(80), "scale": torch.ones(80)}
    voc = jatts_b200.Vocoder(recipes.make_hifigan_state_dict(cfg_hg, seed=0),
                             {"generator_type": "HiFiGANGenerator", "generator_params": dict(cfg_hg),
                              "sampling_rate": recipes.SAMPLING_RATE}, stats, dev, trg_stats=stats)
    texts_cpu = [recipes.make_phonemes(T_TEXT, 1000 * rank + i, cfg_fs2["idim"]) for i in range(BATCH)]
    tok_host = torch.cat(texts_cpu).pin_memory()
    texts_dev = [t.to(dev) for t in texts_cpu]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step_device():
        outs = model.inference_batch(texts_dev)
        waves = voc.decode_batch([o["feat_gen"] for o in outs])
        return outs, waves

    wave_host = None

    def step_e2e():
        nonlocal wave_host
        tok = tok_host.to(dev, non_blocking=True)
        outs = model.inference_batch(list(tok.split(T_TEXT)))
        waves = voc.decode_batch([o["feat_gen"] for o in outs])
        flat = torch.cat(waves)
        if wave_host is None or wave_host.numel() != flat.numel():
            wave_host = torch.empty(flat.numel(), dtype=torch.float32).pin_memory()
        wave_host.copy_(flat, non_blocking=True)
        return outs, flat

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        outs, waves = step_device()
    torch.cuda.synchronize(dev)
    frames = sum(int(o["feat_gen"].shape[0]) for o in outs)
    audio_s = frames * recipes.HOP_SIZE / recipes.SAMPLING_RATE

    def timed(fn, steps):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for a, b in ev:
            flush.fill_(1)  # L2 flush between timed iterations (outside the event pair)
            a.record()
            fn()
            b.record()
        barrier()
        return sum(a.elapsed_time(b) for a, b in ev)  # ms

    sampler = ClockSampler(local_rank if "CUDA_VISIBLE_DEVICES" not in os.environ else
                           int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank]))
    if rank == 0:
        # nvidia-smi needs a moment to produce its first line: keep the GPU under the same load until it does
        sampler.start()
        t_wait = time.time()
        while sampler.mark() == 0 and time.time() - t_wait < 5.0:
            step_device()
            torch.cuda.synchronize(dev)
    l0 = _lib.launch_count()
    s0 = sampler.mark()
    ms_dev = timed(step_device, args.steps)
    launches = (_lib.launch_count() - l0) // args.steps
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop(s0, sampler.mark()) if rank == 0 else None

    # roofline leg: device time of the dominant kernel (tcgen05 conv GEMM, bf16 = HiFi-GAN convolutions)
    mels = [o["feat_gen"] for o in outs]
    torch.cuda.synchronize(dev)
    _lib.profile_begin()
    voc.decode_batch(mels)
    prof = _lib.profile_end()
    _lib.profile_begin()
    model.inference_batch(texts_dev)
    prof_fs2 = _lib.profile_end()

    t = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device=dev)
    a = torch.tensor([audio_s], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(a, op=torch.distributed.ReduceOp.SUM)
    ms_dev, ms_e2e = float(t[0]), float(t[1])
    total_audio = float(a[0])
    if rank != 0:
        return
    peak, peak_src = measured_peaks()
    traffic = measured_traffic()
    hg_flops = hifigan_flops_per_frame(cfg_hg) * frames
    conv_share = 1.0 - 2.0 * (cfg_hg["channels"] >> 4) * 7 * 300 / hifigan_flops_per_frame(cfg_hg)  # minus output_conv
    achieved = hg_flops * conv_share / (prof["ms_bf16"] * 1e-3) / 1e12
    fs2_fl = sum(fs2_flops(cfg_fs2, T_TEXT, int(o["feat_gen"].shape[0])) for o in outs)
    cpu_cores = os.cpu_count() or 1
    torch.set_num_threads(cpu_cores)
    cpu_audio, cpu_times = time_cpu(4, reps=20, warm=1)   # ~10 s of CPU work
    cpu_value = cpu_audio * len(cpu_times) / sum(cpu_times)
    line = {
        "metric": METRIC, "value": total_audio * args.steps / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(world, {"mel_frames_per_gpu": frames,
                   "audio_seconds_per_step_per_gpu": audio_s, "sampling_rate": recipes.SAMPLING_RATE,
                   "hop_size": recipes.HOP_SIZE,
                   "l2": "256 MB buffer written between timed steps; per-step activation working set ~2 GB >> 126 MB L2",
                   "precision": "HiFi-GAN: bf16 operands / fp32 TMEM accumulate; FastSpeech2 GEMMs: fp16 hi+lo split (3 MMA) / fp32"}),
        "e2e": {"value": total_audio * args.steps / (ms_e2e * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": int(tok_host.numel() * 8), "d2h_bytes_per_step": int(frames * recipes.HOP_SIZE * 4),
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor",
                     "kernel": "HiFi-GAN convolution kernels: mrf_pair_kernel<C,k> (fused residual units, C = 32/64) + "
                               "conv_bf16_tma_kernel<*> (all other Conv1d / ConvTranspose1d launches of one step)",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "peak_source": peak_src,
                     "traffic": (traffic or {}).get("hifigan_conv_dram_bytes_per_step"),
                     "traffic_source": (traffic or {}).get("source"),
                     "launches_per_step": int(prof["n_bf16"]), "kernel_ms_per_step": prof["ms_bf16"],
                     "algorithmic_flop_per_step": hg_flops * conv_share,
                     "fs2_split_gemm": {"launches_per_step": int(prof_fs2["n_split"]), "kernel_ms_per_step": prof_fs2["ms_split"],
                                        "algorithmic_flop_per_step": fs2_fl}},
        "cpu_baseline": {"value": cpu_value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": "4 of the 64 utterances (50 phonemes each) x 20 repetitions after 1 warm-up, per-utterance loop as "
                                   "tts_decode.py; oracle port of the reference arithmetic, fp32 torch CPU"},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
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
        run_b200(args, rank, world, local_rank)
    finally:
        if world > 1:
            torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
