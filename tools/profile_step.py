"""One warm-up + one profiled step of the bench workload (run under ncu; not a bench)."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))  # repo root
import sys
import torch
import jatts_b200
from oracle import recipes

which = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = torch.device("cuda", 0)
cfg_fs2, cfg_hg = recipes.JSUT_FS2, recipes.HIFIGAN_V1_HOP300
model = jatts_b200.FastSpeech2(**cfg_fs2)
model.load_state_dict(recipes.make_fs2_state_dict(cfg_fs2, seed=0, duration_recipe="A"))
model = model.eval().to(dev)
stats = {"mean": torch.zeros(80), "scale": torch.ones(80)}
voc = jatts_b200.Vocoder(recipes.make_hifigan_state_dict(cfg_hg, seed=0),
                         {"generator_type": "HiFiGANGenerator", "generator_params": dict(cfg_hg), "sampling_rate": 24000},
                         stats, dev, trg_stats=stats)
texts = [recipes.make_phonemes(50, i, cfg_fs2["idim"]).to(dev) for i in range(64)]
outs = model.inference_batch(texts)
mels = [o["feat_gen"] for o in outs]
voc.decode_batch(mels)
torch.cuda.synchronize()
torch.cuda.profiler.start()
if which in ("all", "fs2"):
    outs = model.inference_batch(texts)
if which in ("all", "voc"):
    voc.decode_batch(mels)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
