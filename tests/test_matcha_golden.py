"""Matcha-TTS against golden vectors produced by the REAL reference (tests/golden/make_golden.py::main_matcha runs the
unmodified jatts.models.matchatts.MatchaTTS.inference; files: tests/golden/matcha_*.npz with the noise the reference drew).
The reference tree does not exist on the GPU box; these files are what ties the oracle AND the CUDA path to it there."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import matcha as om
from oracle import recipes

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
make_golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_golden)

CASES = sorted(make_golden.MATCHA_CASES)


def load(name):
    cfg, wseed, texts, spembs, steps, temp = make_golden.matcha_case_inputs(name)
    gold = np.load(os.path.join(HERE, "golden", name + ".npz"))
    sd = recipes.make_matcha_state_dict(cfg, seed=wseed, duration_recipe="A")
    return cfg, sd, texts, spembs, steps, temp, gold


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_the_golden_vectors(name):
    cfg, sd, texts, spembs, steps, temp, gold = load(name)
    torch.set_num_threads(os.cpu_count() or 1)
    for i, x in enumerate(texts):
        z = torch.from_numpy(gold[f"noise_{i}"]).t()
        got = om.matcha_inference(sd, cfg, x, z, steps, temp, spemb=None if spembs is None else spembs[i])
        assert torch.equal(got["duration"], torch.from_numpy(gold[f"duration_{i}"]))
        want = torch.from_numpy(gold[f"feat_gen_{i}"])
        assert got["feat_gen"].shape == want.shape
        assert float((got["feat_gen"] - want).abs().max()) < 2e-4


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_path_matches_the_golden_vectors(name):
    """durations bit-exact, mel < 1e-3 max-abs against what the reference itself produced (measured: ~1e-5)"""
    import jatts_b200

    cfg, sd, texts, spembs, steps, temp, gold = load(name)
    model = jatts_b200.MatchaTTS(**cfg)
    model.load_state_dict(sd)
    model = model.eval().to("cuda")
    noise = [torch.from_numpy(gold[f"noise_{i}"]) for i in range(len(texts))]
    outs = model.inference_batch(texts, spembs=spembs, n_timesteps=steps, temperature=temp, noise=noise)
    worst = 0.0
    for i, o in enumerate(outs):
        assert torch.equal(o["duration"].cpu(), torch.from_numpy(gold[f"duration_{i}"]))
        want = torch.from_numpy(gold[f"feat_gen_{i}"])
        assert tuple(o["feat_gen"].shape) == tuple(want.shape)
        worst = max(worst, float((o["feat_gen"].cpu() - want).abs().max()))
    print(f"{name}: mel max-abs error vs the reference's own output {worst:.3e}")
    assert worst < 1e-3, worst
