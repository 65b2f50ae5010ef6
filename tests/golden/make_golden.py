"""Generate the golden vectors under tests/golden/ by running the REAL reference code.

Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py
It imports the unmodified ``jatts.models.fastspeech2.FastSpeech2`` (oracle/ref_loader.py), loads the
seeded weights of oracle/recipes.py into it and stores what ``inference()`` returns.  The files travel
to the GPU box, the reference does not.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import recipes, ref_loader  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# name: (config name, weight seed, duration recipe, [text lengths], text seed base, alpha)
CASES = {
    "fs2_jsut_A": ("JSUT_FS2", 0, "A", [50], 1234, 1.0),
    "fs2_jsut_B": ("JSUT_FS2", 1, "B", [50, 23], 77, 1.0),
    "fs2_jsut_alpha": ("JSUT_FS2", 0, "A", [31], 5, 1.3),
    "fs2_jsut_Z": ("JSUT_FS2", 0, "Z", [12], 9, 1.0),
    "fs2_jvs_A": ("JVS_FS2", 2, "A", [40, 64], 300, 1.0),
}


def case_inputs(name):
    cfg_name, wseed, recipe, lens, tseed, alpha = CASES[name]
    cfg = getattr(recipes, cfg_name)
    texts = [recipes.make_phonemes(t, tseed + i, cfg["idim"]) for i, t in enumerate(lens)]
    spembs = recipes.make_spembs(len(lens), tseed) if cfg.get("spk_embed_dim") else None
    return cfg, wseed, recipe, texts, spembs, alpha


def main():
    torch.set_num_threads(8)
    for name in CASES:
        cfg, wseed, recipe, texts, spembs, alpha = case_inputs(name)
        sd = recipes.make_fs2_state_dict(cfg, seed=wseed, duration_recipe=recipe)
        model = ref_loader.build_reference_model(cfg, sd)
        out = {}
        with torch.no_grad():
            for i, x in enumerate(texts):
                r = model.inference(x, spembs=None if spembs is None else spembs[i], alpha=alpha)
                out[f"feat_gen_{i}"] = r["feat_gen"].numpy().astype(np.float32)
                out[f"duration_{i}"] = r["duration"].numpy().astype(np.int64)
                out[f"pitch_{i}"] = r["pitch"].numpy().astype(np.float32)
                out[f"energy_{i}"] = r["energy"].numpy().astype(np.float32)
                print(name, i, "T_text", x.shape[0], "frames", r["feat_gen"].shape[0],
                      "dur min/max", int(r["duration"].min()), int(r["duration"].max()))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


if __name__ == "__main__":
    main()
