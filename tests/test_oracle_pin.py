"""Pin the CPU oracle: against golden vectors produced by the real reference (always) and against the
real reference code itself (when /root/reference is present, i.e. in the build container)."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import fs2, recipes, ref_loader

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import make_golden  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
needs_ref = pytest.mark.skipif(not ref_loader.reference_available(), reason="/root/reference not present")


@pytest.mark.parametrize("name", list(make_golden.CASES))
def test_oracle_matches_golden(name):
    cfg, wseed, recipe, texts, spembs, alpha = make_golden.case_inputs(name)
    sd = recipes.make_fs2_state_dict(cfg, seed=wseed, duration_recipe=recipe)
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    for i, x in enumerate(texts):
        o = fs2.fs2_inference(sd, cfg, x, spemb=None if spembs is None else spembs[i], alpha=alpha)
        assert o["duration"].tolist() == z[f"duration_{i}"].tolist()           # bit-exact
        assert o["feat_gen"].shape == z[f"feat_gen_{i}"].shape
        assert np.abs(o["feat_gen"].numpy() - z[f"feat_gen_{i}"]).max() < 5e-5  # fp32 noise floor ~4e-6
        assert np.abs(o["pitch"].numpy() - z[f"pitch_{i}"]).max() < 2e-5
        assert np.abs(o["energy"].numpy() - z[f"energy_{i}"]).max() < 2e-5


def test_golden_covers_edge_cases():
    zb = np.load(os.path.join(GOLDEN, "fs2_jsut_B.npz"))
    assert (zb["duration_0"] == 0).any(), "recipe B must contain zero-duration tokens"
    zz = np.load(os.path.join(GOLDEN, "fs2_jsut_Z.npz"))
    assert (zz["duration_0"] == 1).all() and zz["feat_gen_0"].shape[0] == zz["duration_0"].shape[0]


@needs_ref
def test_state_dict_layout_matches_reference():
    for cfg in (recipes.JSUT_FS2, recipes.JVS_FS2):
        ref = ref_loader.load_reference_fastspeech2()(**cfg)
        sd = ref.state_dict()
        shapes = recipes.fs2_state_shapes(cfg)
        assert list(sd.keys()) == list(shapes.keys())
        for k, v in sd.items():
            assert tuple(v.shape) == tuple(shapes[k]), k


@needs_ref
def test_oracle_matches_real_reference_live():
    cfg = recipes.JSUT_FS2
    sd = recipes.make_fs2_state_dict(cfg, seed=3, duration_recipe="B")
    model = ref_loader.build_reference_model(cfg, sd)
    for seed, t in ((11, 17), (12, 40)):
        x = recipes.make_phonemes(t, seed)
        with torch.no_grad():
            r = model.inference(x)
        o = fs2.fs2_inference(sd, cfg, x)
        assert torch.equal(r["duration"], o["duration"])
        assert (r["feat_gen"] - o["feat_gen"]).abs().max() < 5e-5


@needs_ref
def test_reference_transformer_branch_is_dead_code():
    """SURVEY finding 5: encoder_type='transformer' raises NameError in the reference."""
    cls = ref_loader.load_reference_fastspeech2()
    with pytest.raises(NameError):
        cls(**dict(recipes.JSUT_FS2, encoder_type="transformer", decoder_type="transformer"))


def test_rel_shift_closed_form_equals_legacy_view_trick():
    g = torch.Generator().manual_seed(0)
    for t in (1, 2, 3, 13, 31):
        bd = torch.randn(2, t, t, generator=g)
        assert torch.equal(fs2.rel_shift_legacy(bd), fs2.rel_shift_closed_form(bd))


def test_legacy_pe_rows_are_reversed_positions():
    pe = fs2.legacy_rel_pe_table(8, 6)
    pos = torch.tensor([4999.0 - n for n in range(6)]).unsqueeze(1)
    div = torch.exp(torch.arange(0, 8, 2, dtype=torch.float32) * -(np.log(10000.0) / 8))
    assert torch.allclose(pe[:, 0::2], torch.sin(pos * div)) and torch.allclose(pe[:, 1::2], torch.cos(pos * div))
