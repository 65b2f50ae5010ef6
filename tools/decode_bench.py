"""Stage-4 front-end throughput: a synthetic recipe directory (csv of N 30..70-phoneme utterances, tokens.txt, stats.h5,
FastSpeech2 + HiFi-GAN V1 hop-300 checkpoints with seeded random weights) decoded by `python -m jatts_b200.decode`
on 1 process, then on --gpus processes (torchrun, one per GPU); reports utterances/s and audio-s/s per run.

    python tools/decode_bench.py --utts 10000 --gpus 8 --out gpurun_out/r02_decode_bench.json
"""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))  # repo root
_sys.path.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "tests"))
import argparse
import glob
import json
import shutil
import subprocess
import sys
import time

import torch
import yaml

from h5_writer import write_h5
from oracle import recipes

ap = argparse.ArgumentParser()
ap.add_argument("--utts", type=int, default=10000)
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--out", default="gpurun_out/decode_bench.json")
ap.add_argument("--work", default="/tmp/jatts_decode_bench")
args = ap.parse_args()

w = args.work
shutil.rmtree(w, ignore_errors=True)
_os.makedirs(w)
cfg, hcfg = recipes.JSUT_FS2, recipes.HIFIGAN_V1_HOP300
vocab = ["<blank>", "<unk>"] + [f"p{i}" for i in range(2, cfg["idim"] - 1)] + ["<sos/eos>"]
open(f"{w}/tokens.txt", "w").write("\n".join(vocab) + "\n")
g = torch.Generator().manual_seed(0)
lens = torch.randint(30, 71, (args.utts,), generator=g).tolist()
with open(f"{w}/dev.csv", "w") as f:
    f.write("sample_id,phonemes\n")
    for i, n in enumerate(lens):
        ids = torch.randint(2, cfg["idim"] - 1, (n,), generator=g).tolist()
        f.write(f"utt{i:05d}," + " ".join(vocab[t] for t in ids) + "\n")
st = recipes.make_stats(1)
write_h5(f"{w}/stats.h5", {"mel_mean": st["mean"].numpy(), "mel_scale": st["scale"].numpy()})
write_h5(f"{w}/voc_stats.h5", {"mean": st["mean"].numpy(), "scale": st["scale"].numpy()})
torch.save({"model": {"generator": recipes.make_hifigan_state_dict(hcfg, 0)}}, f"{w}/voc.pkl")
plain = {k: (list(map(list, v)) if k == "resblock_dilations" else list(v) if isinstance(v, tuple) else v) for k, v in hcfg.items()}
yaml.safe_dump({"generator_type": "HiFiGANGenerator", "sampling_rate": 24000, "generator_params": plain}, open(f"{w}/voc_config.yml", "w"))
torch.save({"model": recipes.make_fs2_state_dict(cfg, seed=0, duration_recipe="A")}, f"{w}/checkpoint-1steps.pkl")
yaml.safe_dump({"model_type": "FastSpeech2", "model_params": dict(cfg), "out_feat_type": "mel", "feat_list": ["mel"], "sampling_rate": 24000,
                "vocoder": {"checkpoint": f"{w}/voc.pkl", "config": f"{w}/voc_config.yml", "stats": f"{w}/voc_stats.h5"}},
               open(f"{w}/config.yml", "w"))

results = []
for n in sorted({1, args.gpus}):
    out = f"{w}/out_n{n}"
    base = ["--csv", f"{w}/dev.csv", "--stats", f"{w}/stats.h5", "--token-list", f"{w}/tokens.txt", "--token-column", "phonemes",
            "--outdir", out, "--checkpoint", f"{w}/checkpoint-1steps.pkl", "--verbose", "0"]
    cmd = ([sys.executable, "-m", "jatts_b200.decode"] if n == 1 else
           [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
            "--master-port", "29533", "-m", "jatts_b200.decode"]) + base
    t0 = time.time()
    subprocess.run(cmd, check=True, cwd=_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
    wall_all = time.time() - t0
    ranks = [json.load(open(p)) for p in sorted(glob.glob(f"{out}/decode_stats.rank*.json"))]
    audio = sum(r["audio_seconds"] for r in ranks)
    wall = max(r["wall_seconds"] for r in ranks)
    files = len(glob.glob(f"{out}/wav/*.wav"))
    results.append({"gpus": n, "utterances": sum(r["utterances"] for r in ranks), "wav_files": files, "audio_seconds": audio,
                    "decode_loop_wall_s_max_over_ranks": wall, "process_wall_s_incl_model_load": wall_all,
                    "utterances_per_second": sum(r["utterances"] for r in ranks) / wall, "audio_seconds_per_second": audio / wall,
                    "per_rank_wall_s": [round(r["wall_seconds"], 3) for r in ranks]})
    print(json.dumps(results[-1]), flush=True)
    shutil.rmtree(out, ignore_errors=True)
_os.makedirs(_os.path.dirname(args.out) or ".", exist_ok=True)
json.dump({"utts": args.utts, "runs": results,
           "note": "decode loop wall = tokenised csv -> wav files on local disk (model load and process start excluded); "
                   "reference loop: 1 utterance per iteration, n_gpus=1 (egs/jsut/tts1/run.sh:246)"}, open(args.out, "w"), indent=1)
