"""First-contact end-to-end diagnostic (GPU box): CUDA path vs the CPU oracle, verbose."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))  # repo root
import sys
import time

import torch

import jatts_b200
from jatts_b200 import _lib
from oracle import fs2 as ofs2
from oracle import hifigan as ohg
from oracle import recipes

torch.set_num_threads(16)
dev = "cuda"
which = sys.argv[1] if len(sys.argv) > 1 else "all"
import traceback

def run_hifigan():
    for cfgname in (["HIFIGAN_TINY"] if which == "hifigan_tiny" else ["HIFIGAN_TINY", "HIFIGAN_V1_HOP300"]):
        cfg = getattr(recipes, cfgname)
        sd = recipes.make_hifigan_state_dict(cfg, 0)
        gen = jatts_b200.HiFiGANGenerator(**cfg)
        gen.load_state_dict(sd)
        gen = gen.eval().to(dev)
        mels = [recipes.make_mel(t, i) for i, t in enumerate([37, 50, 8])]
        t0 = time.time()
        ys = gen.inference_batch(mels)
        torch.cuda.synchronize()
        print(cfgname, "gpu time", time.time() - t0, "launches", _lib.launch_count(), flush=True)
        for m, y in zip(mels, ys):
            ref = ohg.hifigan_forward(sd, cfg, m)
            yc = y.cpu()
            print("  T", m.shape[0], "shape", tuple(yc.shape), tuple(ref.shape), "snr_db %.2f" % ohg.ac_snr_db(ref, yc),
                  "maxabs %.4f" % float((ref - yc).abs().max()), "ref std %.3f" % float(ref.std()), flush=True)

def run_fs2():
    for cfgname, T in ([("TINY_FS2", [13, 7, 20])] if which == "fs2_tiny" else [("TINY_FS2", [13, 7, 20]), ("JSUT_FS2", [50, 31, 50, 5])]):
        cfg = getattr(recipes, cfgname)
        for recipe in ("A", "B"):
            sd = recipes.make_fs2_state_dict(cfg, seed=1, duration_recipe=recipe)
            kw = {k: v for k, v in cfg.items()}
            model = jatts_b200.FastSpeech2(**kw)
            model.load_state_dict(sd)
            model = model.eval().to(dev)
            texts = [recipes.make_phonemes(t, 100 + i, cfg["idim"]) for i, t in enumerate(T)]
            t0 = time.time()
            outs = model.inference_batch(texts, return_lr_index=True)
            torch.cuda.synchronize()
            print(cfgname, recipe, "gpu time", time.time() - t0, flush=True)
            for x, o in zip(texts, outs):
                ref = ofs2.fs2_inference(sd, cfg, x, return_intermediates=True)
                same_d = torch.equal(ref["duration"], o["duration"].cpu())
                nf = (ref["feat_gen"].shape[0], o["feat_gen"].shape[0])
                line = f"  T {x.shape[0]} dur_equal {same_d} frames {nf}"
                line += " pitch %.2e energy %.2e" % (float((ref["pitch"] - o["pitch"].cpu()).abs().max()),
                                                     float((ref["energy"] - o["energy"].cpu()).abs().max()))
                if nf[0] == nf[1]:
                    line += " mel maxabs %.3e" % float((ref["feat_gen"] - o["feat_gen"].cpu()).abs().max())
                    line += " lr_equal %s" % torch.equal(ref["lr_index"].int(), o["lr_index"].cpu())
                else:
                    line += f" ref_d {ref['duration'].tolist()} got {o['duration'].cpu().tolist()}"
                print(line, flush=True)
for fn, keys in ((run_hifigan, ("all", "hifigan_tiny", "hifigan")), (run_fs2, ("all", "fs2_tiny", "fs2"))):
    if which in keys:
        try:
            fn()
        except Exception:
            traceback.print_exc()
print("done", flush=True)
