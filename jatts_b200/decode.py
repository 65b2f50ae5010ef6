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
  * copies the int16 samples to a ring of pinned host buffers asynchronously and writes the ``.wav`` files from a
    writer thread while the next batch is on the GPU,
  * shards the utterances over the GPUs of the box (one process per GPU, ``torchrun`` or ``--rank/--world-size``;
    deterministic greedy partition by length, no communication: jatts_b200/shard.py) -- the reference pins stage 4 to
    ``n_gpus=1`` (egs/jsut/tts1/run.sh:246),
  * looks speaker embeddings up per ``ref_wav_path`` through a cache, so a multi-speaker run extracts (or loads) each
    reference once instead of once per utterance (tts_decode.py:209-212 calls the extractor inside the loop).

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
    ``stats.h5`` (h5py when installed, else the built-in reader jatts_b200/_h5lite.py) or an ``.npz`` with the same keys."""
    if isinstance(path, dict):
        return path[f"{prefix}_mean"], path[f"{prefix}_scale"]
    if str(path).endswith(".npz"):
        import numpy as np

        z = np.load(path)
        return z[f"{prefix}_mean"], z[f"{prefix}_scale"]
    try:
        import h5py
    except ImportError:
        from ._h5lite import read_hdf5

        return read_hdf5(path, f"{prefix}_mean"), read_hdf5(path, f"{prefix}_scale")
    with h5py.File(path, "r") as f:
        return f[f"{prefix}_mean"][()], f[f"{prefix}_scale"][()]


class SpeakerEmbeddingCache:
    """Speaker embeddings for the decode loop (tts_decode.py:146-152, 209-212).  The reference runs its extractor
    (``SpeechBrainSpkEmbExtractor.forward(ref_wav_path)``, spkemb_speechbrain.py:14-28, an ECAPA-TDNN on the CPU) once
    PER UTTERANCE inside the loop; the embedding only depends on the reference wav, so it is looked up here by key:

      * ``table``: precomputed embeddings keyed by ``sample_id`` or ``ref_wav_path`` (an ``.npz``), and / or
      * ``extractor``: any callable ``path -> 1-D embedding`` (e.g. the reference's extractor object's ``forward``) that
        is called once per distinct ``ref_wav_path`` and memoised.

    The extractor itself is a third-party model (speechbrain, downloaded at run time) and is out of scope."""

    def __init__(self, table=None, extractor=None):
        self.table = dict(table) if table is not None else {}
        self.extractor = extractor
        self.calls = 0

    def __call__(self, item: dict):
        import torch

        for key in (item.get("sample_id"), item.get("ref_wav_path")):
            if key is not None and key in self.table:
                return torch.as_tensor(self.table[key], dtype=torch.float32).reshape(-1)
        ref = item.get("ref_wav_path")
        if self.extractor is None or ref is None:
            raise KeyError(f"no speaker embedding for {item.get('sample_id')!r} (neither a table entry nor an extractor / ref_wav_path)")
        self.calls += 1
        emb = torch.as_tensor(self.extractor(ref), dtype=torch.float32).reshape(-1)
        self.table[ref] = emb
        return emb


class PinnedRing:
    """A few reusable pinned host buffers: a fresh ``pin_memory()`` per batch is a page-locking system call (and a
    synchronisation) each time.  A buffer is reused when the consumer has released it (``event`` / writer done)."""

    def __init__(self, dtype, slots: int = 4):
        import torch

        self.dtype, self.slots = dtype, slots
        self.bufs = [torch.empty(0, dtype=dtype).pin_memory() for _ in range(slots)]
        self.free = [threading.Event() for _ in range(slots)]
        for e in self.free:
            e.set()
        self.i = 0

    def take(self, n: int):
        """-> (slot index, pinned 1-D view of n elements); blocks until the slot's previous user released it"""
        import torch

        k = self.i
        self.i = (self.i + 1) % self.slots
        self.free[k].wait()
        self.free[k].clear()
        if self.bufs[k].numel() < n:
            self.bufs[k] = torch.empty(max(n, 2 * self.bufs[k].numel()), dtype=self.dtype).pin_memory()
        return k, self.bufs[k][:n]

    def release(self, k: int):
        self.free[k].set()


class WavWriter:
    """Writer threads: take (pinned int16 buffer, CUDA event, [(path, begin, end)], sampling rate) jobs, wait for the copy
    to land and write the files while the next batches run.  File writes release the GIL, so a few threads keep up with
    the GPU where one (measured: ~180 MB/s of PCM_16 through Python file objects) does not."""

    def __init__(self, threads: int = 4):
        self.q: "queue.Queue" = queue.Queue(maxsize=2 * threads)
        self.error: Optional[BaseException] = None
        self.files = 0
        self._lock = threading.Lock()
        self.threads = [threading.Thread(target=self._run, daemon=True) for _ in range(max(1, threads))]
        for t in self.threads:
            t.start()

    def _run(self):
        while True:
            job = self.q.get()
            if job is None:
                return
            try:
                host, event, entries, sr, done = job
                try:
                    if event is not None:
                        event.synchronize()
                    arr = host.numpy()
                    for path, a, b in entries:
                        write_wav_pcm16(path, arr[a:b], sr)
                    with self._lock:
                        self.files += len(entries)
                finally:
                    if done is not None:
                        done()      # the pinned buffer goes back to the ring
            except BaseException as e:  # surfaced by close()
                self.error = e

    def submit(self, host, event, entries, sr, done=None):
        if self.error is not None:
            raise self.error
        self.q.put((host, event, entries, sr, done))

    def close(self):
        for _ in self.threads:
            self.q.put(None)
        for t in self.threads:
            t.join()
        if self.error is not None:
            raise self.error


# --------------------------------------------------------------------------------------------------------------
# the decode loop
# --------------------------------------------------------------------------------------------------------------
def decode_items(model, vocoder, items: Sequence[dict], outdir: str, sampling_rate: int, device, max_utts: int = 64,
                 max_tokens: int = 8192, spembs=None, rank: int = 0, world_size: int = 1, writer_threads: int = 4,
                 inference_kwargs: Optional[dict] = None) -> dict:
    """Synthesise this rank's share of ``items`` (dicts with ``sample_id`` and ``token_indices``) into
    ``outdir/wav/<sample_id>.wav``.  ``spembs``: a ``SpeakerEmbeddingCache`` (or a plain dict keyed by ``sample_id`` /
    ``ref_wav_path``) for multi-speaker models.  With ``world_size > 1`` every rank calls this with the same ``items``
    and decodes the utterances ``shard_utterances`` assigns to it (no communication; each file is written by exactly
    one rank).  ``inference_kwargs``: extra keyword arguments of ``model.inference_batch`` (Matcha-TTS: ``n_timesteps`` and
    ``temperature``, tts_decode.py:216-226).  Returns counters (utterances, batches, frames, audio seconds, wall seconds)
    of this rank."""
    import torch

    from .shard import shard_utterances

    wav_dir = os.path.join(outdir, "wav")
    os.makedirs(wav_dir, exist_ok=True)
    all_lengths = [len(it["token_indices"]) for it in items]
    mine = shard_utterances(all_lengths, world_size)[rank] if world_size > 1 else list(range(len(items)))
    lengths = [all_lengths[j] for j in mine]
    batches = [[mine[k] for k in b] for b in plan_batches(lengths, max_utts, max_tokens)]
    if spembs is not None and not callable(spembs):
        spembs = SpeakerEmbeddingCache(table=spembs)
    writer = WavWriter(writer_threads)
    hop = vocoder.model.hop
    frames_total, t0 = 0, time.time()
    copy_stream = torch.cuda.Stream(device=device)
    tok_ring, pcm_ring = PinnedRing(torch.long, 2), PinnedRing(torch.int16, 8)
    skipped: List[str] = []

    def run_batch(batch):
        nonlocal frames_total
        n_tok = sum(all_lengths[j] for j in batch)
        tk, tok_host = tok_ring.take(n_tok)
        tok_host.copy_(torch.tensor([t for j in batch for t in items[j]["token_indices"]], dtype=torch.long))
        tok = tok_host.to(device, non_blocking=True)
        h2d = torch.cuda.Event()
        h2d.record()
        texts = list(tok.split([all_lengths[j] for j in batch]))
        sp = None
        if model.spk_embed_dim is not None:
            if spembs is None:
                raise ValueError("the model is speaker conditioned: pass speaker embeddings (--spkemb-npz)")
            sp = torch.stack([spembs(items[j]) for j in batch]).to(device)
        try:
            outs = model.inference_batch(texts, spembs=sp, **(inference_kwargs or {}))
        finally:
            h2d.synchronize()
            tok_ring.release(tk)
        pcm = vocoder.decode_batch([o["feat_gen"] for o in outs], pcm16=True)
        flat = torch.cat(pcm)
        pk, host = pcm_ring.take(flat.numel())
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
        writer.submit(host, done, entries, sampling_rate, done=lambda k=pk: pcm_ring.release(k))

    try:
        for batch in batches:
            try:
                run_batch(batch)
            except NotImplementedError as e:
                # an utterance of the batch expands to more frames than the model's max_len (the reference has no such
                # limit: it extends its positional table) or, Matcha-TTS, to fewer than the 2 frames its decoder needs (the
                # reference's output is empty for it): do not lose the rest of the batch -- retry one by one and report
                # the utterances that really do not fit
                fits = lambda err: not any(k in str(err) for k in ("max_len", "fewer than 2 frames"))
                if fits(e):
                    raise
                for j in (batch if len(batch) > 1 else []):
                    try:
                        run_batch([j])
                    except NotImplementedError as e1:
                        if fits(e1):
                            raise
                        skipped.append(items[j]["sample_id"])
                        logging.warning("utterance %s cannot be synthesised: skipped (%s)", skipped[-1], e1)
                if len(batch) == 1:
                    skipped.append(items[batch[0]]["sample_id"])
                    logging.warning("utterance %s cannot be synthesised: skipped (%s)", skipped[-1], e)
    finally:
        writer.close()
    wall = time.time() - t0
    audio_s = frames_total * hop / float(sampling_rate)
    return {"utterances": len(mine) - len(skipped), "batches": len(batches), "frames": frames_total, "audio_seconds": audio_s,
            "wall_seconds": wall, "files": writer.files, "skipped": skipped, "rank": rank, "world_size": world_size,
            "utterances_per_second": (len(mine) - len(skipped)) / max(wall, 1e-9),
            "audio_seconds_per_second": audio_s / max(wall, 1e-9)}


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
    ap.add_argument("--max-len", type=int, default=2048,
                    help="longest utterance in mel frames the engine is built for (<= 5000, the reference's positional table)")
    ap.add_argument("--writer-threads", type=int, default=4, help="threads writing the wav files")
    ap.add_argument("--rank", type=int, default=int(os.environ.get("RANK", "0")),
                    help="this process's shard (default: torchrun's RANK)")
    ap.add_argument("--world-size", type=int, default=int(os.environ.get("WORLD_SIZE", "1")),
                    help="number of processes sharing the csv, one per GPU (default: torchrun's WORLD_SIZE)")
    args = ap.parse_args(argv)
    logging.basicConfig(level=logging.DEBUG if args.verbose > 1 else logging.INFO if args.verbose > 0 else logging.WARN,
                        format="%(asctime)s (%(module)s:%(lineno)d) %(levelname)s: %(message)s")
    os.makedirs(args.outdir, exist_ok=True)
    if args.config is None:
        args.config = os.path.join(os.path.dirname(args.checkpoint), "config.yml")
    with open(args.config) as f:
        config = yaml.load(f, Loader=yaml.Loader)
    config.update(vars(args))
    classes = {"FastSpeech2": jatts_b200.FastSpeech2, "FastSpeech2B200": jatts_b200.FastSpeech2,
               "MatchaTTS": jatts_b200.MatchaTTS, "MatchaTTSB200": jatts_b200.MatchaTTS}
    if config["model_type"] not in classes:
        raise NotImplementedError(f"model_type {config['model_type']}: only FastSpeech2 and MatchaTTS have a B200 path")
    if not torch.cuda.is_available():
        raise RuntimeError("jatts_b200.decode needs a CUDA device (there is no CPU fallback)")
    if not 0 <= args.rank < args.world_size:
        raise ValueError("--rank must be in [0, --world-size)")
    local = int(os.environ.get("LOCAL_RANK", args.rank % max(torch.cuda.device_count(), 1)))
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    items = read_items(args.csv, args.token_column, TokenIDConverter(args.token_list))
    logging.info(f"Dataset size = {len(items)}.")
    model = classes[config["model_type"]](**config["model_params"], max_len=args.max_len)
    # tts_decode.py:216-226: Matcha-TTS takes its solver settings from the top level of the training config
    inference_kwargs = {}
    if isinstance(model, jatts_b200.MatchaTTS):
        inference_kwargs = {"temperature": config["temperature"], "n_timesteps": config["ode_steps"]}
    model.load_state_dict(torch.load(args.checkpoint, map_location="cpu")["model"])
    model = model.eval().to(device)
    logging.info(f"Loaded model parameters from {args.checkpoint}.")
    mean, scale = read_stats(args.stats, config["out_feat_type"])
    stats = {"mean": mean, "scale": scale}
    if not config.get("vocoder", False):
        raise NotImplementedError("Griffin-Lim decoding is not part of the B200 path: configure a HiFi-GAN vocoder")
    vocoder = jatts_b200.Vocoder(config["vocoder"]["checkpoint"], config["vocoder"]["config"], config["vocoder"]["stats"],
                                 device, trg_stats=stats)
    # build both engines now (weight repacking + upload: ~2 s), as part of "loading the model", not of the decode loop
    model._get_engine()
    vocoder.model._get_engine(False)
    torch.cuda.synchronize(device)
    spembs = None
    if args.spkemb_npz:
        import numpy as np

        spembs = SpeakerEmbeddingCache(table=dict(np.load(args.spkemb_npz)))
    res = decode_items(model, vocoder, items, args.outdir, vocoder.config["sampling_rate"], device, args.max_utts,
                       args.max_tokens, spembs, rank=args.rank, world_size=args.world_size,
                       writer_threads=args.writer_threads, inference_kwargs=inference_kwargs)
    logging.info("rank %d/%d decoded %d utterances in %d batches: %.1f s of audio in %.2f s (%.0f x real time, %.0f utterances/s)%s" % (
        args.rank, args.world_size, res["utterances"], res["batches"], res["audio_seconds"], res["wall_seconds"],
        res["audio_seconds_per_second"], res["utterances_per_second"],
        f"; skipped (too long for --max-len, or too short for the Matcha decoder): {res['skipped']}" if res["skipped"] else ""))
    with open(os.path.join(args.outdir, f"decode_stats.rank{args.rank}.json"), "w") as f:
        import json

        json.dump(res, f)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
