"""Diagnostic: mel / pitch error statistics of the CUDA FS2 path vs the fp32 oracle over several utterances."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))  # repo root
import os, sys, torch
import jatts_b200
from oracle import fs2 as ofs2, recipes
torch.set_num_threads(16)
cfg = recipes.JSUT_FS2
for wseed, recipe in ((1, "A"), (2, "B")):
    sd = recipes.make_fs2_state_dict(cfg, seed=wseed, duration_recipe=recipe)
    model = jatts_b200.FastSpeech2(**cfg); model.load_state_dict(sd); model = model.eval().to("cuda")
    texts = [recipes.make_phonemes(50, 500 + i, cfg["idim"]) for i in range(8)]
    outs = model.inference_batch(texts)
    errs, perr, deq = [], [], []
    for x, o in zip(texts, outs):
        ref = ofs2.fs2_inference(sd, cfg, x)
        deq.append(torch.equal(ref["duration"], o["duration"].cpu()))
        if ref["feat_gen"].shape == o["feat_gen"].shape:
            errs.append(float((ref["feat_gen"] - o["feat_gen"].cpu()).abs().max()))
        perr.append(float((ref["pitch"] - o["pitch"].cpu()).abs().max()))
    print(os.environ.get("JATTS_B200_LIB", "default"), recipe, "dur_equal", all(deq), "mel max %.2e mean %.2e" % (max(errs), sum(errs) / len(errs)), "pitch max %.2e" % max(perr), flush=True)
