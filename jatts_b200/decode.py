"""Batched stage-4 decoding: the drop-in for ``jatts/bin/tts_decode.py``.

The reference loop (tts_decode.py:203-255) synthesises ONE utterance per iteration: tokens -> device,
``model.inference``, a matplotlib PNG, ``vocoder.decode``, a device->host copy and ``sf.write(..., "PCM_16")``.
With the B200 path at ~10^4 x real time that host loop is the whole run time, so this front-end

  * reads the same inputs (the recipe's csv with ``sample_id`` + the token column, ``tokens.txt``, the text2mel
    ``stats`` file, the checkpoint's ``config.yml``) with the same command-line arguments,
  * plans length-bucketed batches (utterances sorted by token count, a batch closes at ``max_utts`` utterances or
    ``max_tokens`` tokens -- padding never exists on the device, the bound only keeps a batch's memory flat),
  * runs ``FastSpeech2.inference_batch`` -> ``Vocoder.decode_batch(pcm16=True)``: the float -> PCM_16 conversion
    libsndfile would do on the host (``lrintf(y * 0x7FFF)``) is fused into the generator's output convolution,
  * copies the int16 samples to pinned host memory asynchronously and writes the ``.wav`` files from a writer
    thread while the next batch is on the GPU.

Row *i* of a batch equals the reference's per-utterance result, so the wav files hold the
samples stage 4 would have written.  Plots (``outs/*.png``) are not produced: matplotlib is not part of this path.

    python -m jatts_b200.decode --csv data/dev.csv --stats stats.h5 --token-list tokens.txt \\
        --token-column phonemes --outdir exp/out --checkpoint exp/checkpoint-50000steps.pkl
"""
from __future__ import annotations

import argparse
import csv
import logging
import os
import queue
import struct
import threading
import time
from typing import Dict, Iterable, List, Optional, Sequence, Tuple


# --------------------------------------------------------------------------------------------------------------
# host-side pieces (no GPU needed: covered by the CPU test suite)
# --------------------------------------------------------------------------------------------------------------
class TokenIDConverter:
    """jatts/utils/token_id_converter.py:20-80: one token per line, unknown tokens map to ``<unk>``."""

    def __init__(self, token_list, unk_symbol: str = "<unk>"):
        if isinstance(token_list, (str, os.PathLike)):
            with open(token_list, "r", encoding="utf-8") as f:
                self.token_list = [line.rstrip() for line in f]
        else:
            self.token_list = list(token_list)
        self.token2id: Dict[str, int] = {}
        for i, t in enumerate(self.token_list):
            if t in self.token2id:
                raise RuntimeError(f'Symbol "{t}" is duplicated')
            self.token2id[t] = i
        if unk_symbol not in self.token2id:
            raise RuntimeError(f"Unknown symbol '{unk_symbol}' doesn't exist in the token_list")
        self.unk_id = self.token2id[unk_symbol]

    def tokens2ids(self, tokens: Iterable[str]) -> List[int]:
        return [self.token2id.get(t, self.unk_id) for t in tokens]


def read_items(csv_path: str, token_column: str, converter: TokenIDConverter) -> List[dict]:
    """Rows of the recipe csv -> dicts with ``sample_id`` and ``token_indices`` (tts_dataset.py:93-116)."""
    items = []
    with open(csv_path, newline="", encoding="utf-8") as f:
        for row in csv.DictReader(f):
            if "sample_id" not in row or token_column not in row:
                raise KeyError(f"csv needs the columns 'sample_id' and '{token_column}'")
            tokens = [p for p in row[token_column].split(" ") if p != ""]
            item = dict(row)
            item["tokens"] = tokens
            item["token_indices"] = converter.tokens2ids(tokens)
            items.append(item)
    return items


def plan_batches(lengths: Sequence[int], max_utts: int = 64, max_tokens: int = 8192) -> List[List[int]]:
    """Length-bucketed batches: indices sorted by (length, index), cut at ``max_utts`` utterances or ``max_tokens``
    summed tokens.  Every index appears exactly once; empty utterances are rejected (the model needs >= 1 token)."""
    if max_utts < 1 or max_tokens < 1:
        raise ValueError("max_utts and max_tokens must be >= 1")
    for i, n in enumerate(lengths):
        if int(n) <= 0:
            raise ValueError(f"utterance {i} has no tokens")
    order = sorted(range(len(lengths)), key=lambda j: (int(lengths[j]), j))
    batches, cur, tok = [], [], 0
    for j in order:
        n = int(lengths[j])
        if cur and (len(cur) >= max_utts or tok + n > max_tokens):
            batches.append(cur)
            cur, tok = [], 0
        cur.append(j)
        tok += n
    if cur:
        batches.append(cur)
    return batches


def write_wav_pcm16(path: str, samples, sampling_rate: int) -> None:
    """Mono PCM_16 RIFF/WAVE file from int16 samples (numpy array or bytes-like); the sample data is what
    ``sf.write(path, y, sr, "PCM_16")`` stores for ``y`` in [-1, 1]."""
    data = samples.tobytes() if hasattr(samples, "tobytes") else bytes(samples)
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE")
        f.write(b"fmt " + struct.pack("<IHHIIHH", 16, 1, 1, sampling_rate, sampling_rate * 2, 2, 16))
        f.write(b"data" + struct.pack("<I", len(data)))
        f.write(data)


def read_stats(path, prefix: str) -> Tuple["object", "object"]:
    """``{prefix}_mean`` / ``{prefix}_scale`` of the text2mel stats file (tts_decode.py:160-164): the recipe's
    ``stats.h5`` (needs h5py, as the reference does) or an ``.npz`` with the same keys."""
    if isinstance(path, dict):
        return path[f"{prefix}_mean"], path[f"{prefix}_scale"]
    if str(path).endswith(".npz"):
        import numpy as np

        z = np.load(path)
        return z[f"{prefix}_mean"], z[f"{prefix}_scale"]
    try:
        import h5py
    except ImportError as e:
        raise RuntimeError("reading stats.h5 needs h5py; convert it to an .npz with the same keys") from e
    with h5py.File(path, "r") as f:
        return f[f"{prefix}_mean"][()], f[f"{prefix}_scale"][()]


class WavWriter:
    """Writer thread: takes (pinned int16 buffer, CUDA event, [(path, begin, end)], sampling rate) jobs, waits for
    the copy to land and writes the files while the next batch runs."""

    def __init__(self):
        self.q: "queue.Queue" = queue.Queue(maxsize=4)
        self.error: Optional[BaseException] = None
        self.files = 0
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        while True:
            job = self.q.get()
            if job is None:
                return
            try:
                host, event, entries, sr = job
                if event is not None:
                    event.synchronize()
                arr = host.numpy()
                for path, a, b in entries:
                    write_wav_pcm16(path, arr[a:b], sr)
                    self.files += 1
            except BaseException as e:  # surfaced by close()
                self.error = e

    def submit(self, host, event, entries, sr):
        if self.error is not None:
            raise self.error
        self.q.put((host, event, entries, sr))

    def close(self):
        self.q.put(None)
        self.thread.join()
        if self.error is not None:
            raise self.error


# --------------------------------------------------------------------------------------------------------------
# the decode loop
# --------------------------------------------------------------------------------------------------------------
def decode_items(model, vocoder, items: Sequence[dict], outdir: str, sampling_rate: int, device, max_utts: int = 64,
                 max_tokens: int = 8192, spembs: Optional[Dict[str, "object"]] = None) -> dict:
    """Synthesise ``items`` (dicts with ``sample_id`` and ``token_indices``) into ``outdir/wav/<sample_id>.wav``.
    ``spembs`` maps ``sample_id`` (or the item's ``ref_wav_path``) to a speaker embedding for multi-speaker models.
    Returns counters (utterances, batches, frames, audio seconds, wall seconds)."""
    import torch

    wav_dir = os.path.join(outdir, "wav")
    os.makedirs(wav_dir, exist_ok=True)
    lengths = [len(it["token_indices"]) for it in items]
    batches = plan_batches(lengths, max_utts, max_tokens)
    writer = WavWriter()
    hop = vocoder.model.hop
    frames_total, t0 = 0, time.time()
    copy_stream = torch.cuda.Stream(device=device)
    try:
        for batch in batches:
            tok_host = torch.tensor([t for j in batch for t in items[j]["token_indices"]], dtype=torch.long).pin_memory()
            tok = tok_host.to(device, non_blocking=True)
            texts = list(tok.split([lengths[j] for j in batch]))
            sp = None
            if model.spk_embed_dim is not None:
                if spembs is None:
                    raise ValueError("the model is speaker conditioned: pass speaker embeddings (--spkemb-npz)")
                keys = [items[j]["sample_id"] if items[j]["sample_id"] in spembs else items[j].get("ref_wav_path") for j in batch]
                sp = torch.stack([torch.as_tensor(spembs[k], dtype=torch.float32).reshape(-1) for k in keys]).to(device)
            outs = model.inference_batch(texts, spembs=sp)
            pcm = vocoder.decode_batch([o["feat_gen"] for o in outs], pcm16=True)
            flat = torch.cat(pcm)
            host = torch.empty(flat.numel(), dtype=torch.int16).pin_memory()
            done = torch.cuda.Event()
            copy_stream.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(copy_stream):
                host.copy_(flat, non_blocking=True)
                flat.record_stream(copy_stream)
                done.record(copy_stream)
            entries, o = [], 0
            for j, y in zip(batch, pcm):
                n = int(y.numel())
                entries.append((os.path.join(wav_dir, f"{items[j]['sample_id']}.wav"), o, o + n))
                o += n
                frames_total += n // hop
            writer.submit(host, done, entries, sampling_rate)
    finally:
        writer.close()
    wall = time.time() - t0
    audio_s = frames_total * hop / float(sampling_rate)
    return {"utterances": len(items), "batches": len(batches), "frames": frames_total, "audio_seconds": audio_s,
            "wall_seconds": wall, "files": writer.files}


def main(argv=None) -> int:
    import torch
    import yaml

    import jatts_b200

    ap = argparse.ArgumentParser(description="Batched decoding with a trained FastSpeech2 + HiFi-GAN on B200 "
                                             "(same arguments as jatts/bin/tts_decode.py)")
    ap.add_argument("--csv", required=True, type=str)
    ap.add_argument("--stats", required=True, type=str, help="text2mel stats file (.h5 or .npz)")
    ap.add_argument("--token-list", required=True, type=str)
    ap.add_argument("--token-column", required=True, type=str)
    ap.add_argument("--outdir", required=True, type=str)
    ap.add_argument("--checkpoint", required=True, type=str)
    ap.add_argument("--config", default=None, type=str)
    ap.add_argument("--verbose", type=int, default=1)
    ap.add_argument("--max-utts", type=int, default=64, help="utterances per batch")
    ap.add_argument("--max-tokens", type=int, default=8192, help="summed tokens per batch")
    ap.add_argument("--spkemb-npz", default=None, type=str,
                    help="precomputed speaker embeddings keyed by sample_id or ref_wav_path (multi-speaker models)")
    args = ap.parse_args(argv)
    logging.basicConfig(level=logging.DEBUG if args.verbose > 1 else logging.INFO if args.verbose > 0 else logging.WARN,
                        format="%(asctime)s (%(module)s:%(lineno)d) %(levelname)s: %(message)s")
    os.makedirs(args.outdir, exist_ok=True)
    if args.config is None:
        args.config = os.path.join(os.path.dirname(args.checkpoint), "config.yml")
    with open(args.config) as f:
        config = yaml.load(f, Loader=yaml.Loader)
    config.update(vars(args))
    if config["model_type"] != "FastSpeech2":
        raise NotImplementedError(f"model_type {config['model_type']}: only FastSpeech2 has a B200 path")
    if not torch.cuda.is_available():
        raise RuntimeError("jatts_b200.decode needs a CUDA device (there is no CPU fallback)")
    device = torch.device("cuda")
    items = read_items(args.csv, args.token_column, TokenIDConverter(args.token_list))
    logging.info(f"Dataset size = {len(items)}.")
    model = jatts_b200.FastSpeech2(**config["model_params"])
    model.load_state_dict(torch.load(args.checkpoint, map_location="cpu")["model"])
    model = model.eval().to(device)
    logging.info(f"Loaded model parameters from {args.checkpoint}.")
    mean, scale = read_stats(args.stats, config["out_feat_type"])
    stats = {"mean": mean, "scale": scale}
    if not config.get("vocoder", False):
        raise NotImplementedError("Griffin-Lim decoding is not part of the B200 path: configure a HiFi-GAN vocoder")
    vocoder = jatts_b200.Vocoder(config["vocoder"]["checkpoint"], config["vocoder"]["config"], config["vocoder"]["stats"],
                                 device, trg_stats=stats)
    spembs = None
    if args.spkemb_npz:
        import numpy as np

        spembs = dict(np.load(args.spkemb_npz))
    res = decode_items(model, vocoder, items, args.outdir, vocoder.config["sampling_rate"], device, args.max_utts,
                       args.max_tokens, spembs)
    logging.info("decoded %d utterances in %d batches: %.1f s of audio in %.2f s (%.0f x real time)" % (
        res["utterances"], res["batches"], res["audio_seconds"], res["wall_seconds"],
        res["audio_seconds"] / max(res["wall_seconds"], 1e-9)))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
