"""Generate the golden vectors under tests/golden/ by running the REAL reference code.

Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py
It imports the unmodified ``jatts.models.fastspeech2.FastSpeech2`` (and ``jatts.models.matchatts.MatchaTTS``) through
oracle/ref_loader.py, loads the seeded weights of oracle/recipes.py into it and stores what ``inference()`` returns.  The files travel
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


# Matcha-TTS: name: (config name, weight seed, [text lengths], text seed base, Euler steps, temperature, spk)
MATCHA_CASES = {
    "matcha_small": ("SMALL_MATCHA", 3, [9, 17, 4], 500, 10, 0.667, False),
    "matcha_small_spk": ("SMALL_MATCHA", 4, [12, 6], 600, 4, 0.5, True),
    "matcha_jsut": ("JSUT_MATCHA", 5, [14], 700, 10, 0.667, False),
}


def matcha_case_inputs(name):
    cfg_name, wseed, lens, tseed, steps, temp, spk = MATCHA_CASES[name]
    cfg = dict(getattr(recipes, cfg_name))
    if spk:
        cfg.update(spk_embed_dim=24, spk_embed_integration_type="add")
    texts = [recipes.make_phonemes(t, tseed + i, cfg["idim"]) for i, t in enumerate(lens)]
    spembs = recipes.make_spembs(len(lens), tseed, 24) if spk else None
    return cfg, wseed, texts, spembs, steps, temp


def main_matcha():
    """The REAL reference MatchaTTS (matchatts.py, flow_matching.py, decoder.py, transformer.py imported in place; the one
    missing third-party class, diffusers' Attention, is the stand-in of oracle/ref_loader.py) run through its own
    ``inference()``.  The noise it draws inside CFM.inference (``randn_like`` of a permuted view, CPU generator) is
    reproduced with the same call and stored next to the output, so that the GPU box can feed the same z."""
    import logging

    logging.disable(logging.WARNING)
    cls = ref_loader.load_reference_matcha()
    for name in MATCHA_CASES:
        cfg, wseed, texts, spembs, steps, temp = matcha_case_inputs(name)
        sd = recipes.make_matcha_state_dict(cfg, seed=wseed, duration_recipe="A")
        model = cls(**cfg)
        model.load_state_dict(sd, strict=True)
        model.eval()
        out = {}
        with torch.no_grad():
            for i, x in enumerate(texts):
                torch.manual_seed(9000 + i)
                r = model.inference(x, spembs=None if spembs is None else spembs[i], n_timesteps=steps, temperature=temp)
                t_out = r["feat_gen"].shape[0]
                torch.manual_seed(9000 + i)
                z = torch.randn_like(torch.empty(1, t_out, cfg["odim"]).permute(0, 2, 1))[0]      # (odim, T) as drawn
                out[f"feat_gen_{i}"] = r["feat_gen"].numpy().astype(np.float32)
                out[f"duration_{i}"] = r["duration"].numpy().astype(np.int64)
                out[f"noise_{i}"] = z.t().contiguous().numpy().astype(np.float32)                 # (T, odim)
                print(name, i, "T_text", x.shape[0], "frames", t_out, "dur min/max", int(r["duration"].min()),
                      int(r["duration"].max()), "|mel| max", float(r["feat_gen"].abs().max()))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    logging.disable(logging.NOTSET)


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
    if len(sys.argv) < 2 or sys.argv[1] != "matcha":
        main()
    if len(sys.argv) < 2 or sys.argv[1] == "matcha":
        main_matcha()
