"""One Matcha-TTS text2mel call at BASELINE config 5 size (batch 64 x 50 phonemes, 10 Euler steps), for ncu launch lists."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import torch
import jatts_b200
from oracle import recipes

cfg = recipes.JSUT_MATCHA
m = jatts_b200.MatchaTTS(**cfg)
m.load_state_dict(recipes.make_matcha_state_dict(cfg, 0, "A"))
m = m.eval().to("cuda")
texts = [recipes.make_phonemes(50, i, cfg["idim"]).cuda() for i in range(64)]
steps = int(_sys.argv[1]) if len(_sys.argv) > 1 else 10
for _ in range(2):
    out = m.inference_batch(texts, n_timesteps=steps, temperature=0.667)
torch.cuda.synchronize()
print(sum(int(o["feat_gen"].shape[0]) for o in out), "frames")
