"""bench.py's roofline numerators are analytic FLOP counts: check them against torch's own FLOP counter run over the ORACLE
(the reference's arithmetic) so that `roofline.achieved` divides the right number."""
import importlib.util
import os

import torch
from torch.utils.flop_counter import FlopCounterMode

from oracle import fs2 as ofs2
from oracle import hifigan as ohg
from oracle import matcha as om
from oracle import recipes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(bench)


def counted(fn):
    with FlopCounterMode(display=False) as fc:
        fn()
    return float(fc.get_total_flops())


def test_matcha_decoder_flops_match_a_flop_counter():
    cfg = recipes.SMALL_MATCHA
    sd = recipes.make_matcha_state_dict(cfg, 0)
    T = 96
    g = torch.Generator().manual_seed(0)
    x, mu = torch.randn(1, cfg["odim"], T, generator=g), torch.randn(1, cfg["odim"], T, generator=g)
    mask = torch.ones(1, 1, T)
    got = counted(lambda: om.decoder_forward(sd, "decoder.estimator.", x, mask, mu, torch.tensor(0.3), tuple(cfg["decoder_channels"]),
                                             cfg["decoder_n_blocks"], cfg["decoder_num_mid_blocks"], cfg["decoder_num_heads"]))
    want = bench.matcha_decoder_flops(cfg, [T])
    # the counter also sees the time-embedding MLP (a per-step constant the engine takes as a table): < 1 % at this length
    assert abs(got - want) / got < 0.01, (got, want)


def test_hifigan_flops_per_frame_match_a_flop_counter():
    cfg = recipes.HIFIGAN_V1_HOP300
    sd = recipes.make_hifigan_state_dict(cfg, 0)
    T = 12
    got = counted(lambda: ohg.hifigan_forward(sd, cfg, recipes.make_mel(T, 0))) / T
    want = bench.hifigan_flops_per_frame(cfg)
    assert abs(got - want) / got < 0.02, (got, want)      # edge effects of the short clip only
    assert abs(want - 378.77e6) / 378.77e6 < 1e-3


def test_fs2_flops_match_a_flop_counter():
    cfg = recipes.JSUT_FS2
    sd = recipes.make_fs2_state_dict(cfg, 0, duration_recipe="A")
    x = recipes.make_phonemes(20, 3, cfg["idim"])
    out = {}
    got = counted(lambda: out.update(ofs2.fs2_inference(sd, cfg, x)))
    want = bench.fs2_flops(cfg, 20, int(out["feat_gen"].shape[0]))
    # the analytic count leaves out the positional projection of the oracle's per-call pos table and the elementwise work
    assert abs(got - want) / got < 0.05, (got, want)
