"""BASELINE.json configs 3 and 4 at their stated sizes, through the public classes -> C ABI:
  config 3  JVS multi-speaker FastSpeech2 (192-d speaker embedding, "add"), batch 128, variable-length utterances
  config 4  HiFi-GAN V1 vocoder-only bulk synthesis of 10 k synthetic 80-bin mel clips sharded across GPUs
Full-size runs are checked through size-independent properties; a few rows are compared with the CPU oracle."""
import os

import pytest
import torch

import jatts_b200
from oracle import fs2 as ofs2
from oracle import hifigan as ohg
from oracle import recipes

MEL_TOL = 1e-3
SNR_DB = 35.0


@pytest.mark.gpu
def test_config3_jvs_batch128_variable_lengths():
    cfg = recipes.JVS_FS2
    sd = recipes.make_fs2_state_dict(cfg, seed=2, duration_recipe="A")
    model = jatts_b200.FastSpeech2(**cfg)
    model.load_state_dict(sd)
    model = model.eval().to("cuda")
    g = torch.Generator().manual_seed(3)
    lens = torch.randint(20, 81, (128,), generator=g).tolist()                 # T_text ~ U{20..80}, SURVEY 8(d)
    texts = [recipes.make_phonemes(t, i, cfg["idim"]) for i, t in enumerate(lens)]
    spembs = torch.cat([recipes.make_spembs(1, 100 + i) for i in range(128)])   # raw x-vectors; the model normalises
    outs = model.inference_batch(texts, spembs=spembs, return_lr_index=True)
    assert len(outs) == 128
    for x, o in zip(texts, outs):
        d = o["duration"].cpu()
        assert d.shape == (len(x),) and d.dtype == torch.int64 and int(d.sum()) == o["feat_gen"].shape[0]
        assert torch.equal(o["lr_index"].cpu(), torch.repeat_interleave(torch.arange(len(x)), d).int())
        assert torch.isfinite(o["feat_gen"]).all()
    # the speaker embedding matters, and only for its own row
    other = spembs.clone()
    other[5] = recipes.make_spembs(1, 999)[0]
    outs2 = model.inference_batch(texts, spembs=other)
    assert not torch.equal(outs[5]["feat_gen"], outs2[5]["feat_gen"]) or not torch.equal(outs[5]["duration"], outs2[5]["duration"])
    assert torch.equal(outs[6]["feat_gen"], outs2[6]["feat_gen"]) and torch.equal(outs[4]["duration"], outs2[4]["duration"])
    # rows against the per-utterance oracle (the shortest, the longest and two in between)
    torch.set_num_threads(os.cpu_count() or 1)
    order = sorted(range(128), key=lambda i: lens[i])
    for i in (order[0], order[40], order[90], order[-1]):
        ref = ofs2.fs2_inference(sd, cfg, texts[i], spemb=spembs[i])
        assert torch.equal(ref["duration"], outs[i]["duration"].cpu())
        assert float((ref["feat_gen"] - outs[i]["feat_gen"].cpu()).abs().max()) < MEL_TOL


@pytest.mark.gpu
def test_config4_vocoder_only_bulk_sharded():
    cfg = recipes.HIFIGAN_V1_HOP300
    sd = recipes.make_hifigan_state_dict(cfg, 0)
    gen = jatts_b200.HiFiGANGenerator(**cfg)
    gen.load_state_dict(sd)
    gen = gen.eval().to("cuda")
    g = torch.Generator().manual_seed(4)
    lens = torch.randint(200, 401, (10000,), generator=g).tolist()             # 10 k clips, T ~ U{200..400}
    shards = jatts_b200.shard_utterances(lens, 8)
    assert sorted(i for s in shards for i in s) == list(range(10000))
    loads = [sum(lens[i] for i in s) for s in shards]
    assert max(loads) - min(loads) <= 400                                       # balanced to within one clip
    # rank 3's first two batches of 64 clips, and the same clips regrouped: bit-identical per clip
    mine = shards[3][:128]
    mels = {i: recipes.make_mel(lens[i], i) for i in mine}
    out = {}
    for b in (mine[:64], mine[64:]):
        for i, y in zip(b, gen.inference_batch([mels[i] for i in b])):
            assert y.shape == (lens[i] * gen.hop, 1) and torch.isfinite(y).all() and float(y.abs().max()) <= 1.0
            out[i] = y
    regroup = mine[32:96]
    for i, y in zip(regroup, gen.inference_batch([mels[i] for i in regroup])):
        assert torch.equal(y, out[i])
    torch.set_num_threads(os.cpu_count() or 1)
    for i in (mine[0], mine[77]):
        ref = ohg.hifigan_forward(sd, cfg, mels[i])
        assert ohg.ac_snr_db(ref, out[i].cpu()) >= SNR_DB


@pytest.mark.gpu
def test_config4_sharded_output_is_bit_identical_across_world_sizes():
    """SURVEY section 4 / VERDICT r1 missing #3: the N-way sharded run returns the bytes of the 1-GPU run.  Ranks are
    independent processes running this same code on their shard (no collective on the data path), so the partitions of
    world sizes 1, 2, 4 and 8 are replayed here rank by rank with bench.py's batching (length-sorted, 128 clips per
    launch) and compared clip by clip."""
    cfg = recipes.HIFIGAN_V1_HOP300
    gen = jatts_b200.HiFiGANGenerator(**cfg)
    gen.load_state_dict(recipes.make_hifigan_state_dict(cfg, 0))
    gen = gen.eval().to("cuda")
    g = torch.Generator().manual_seed(4)
    lens = torch.randint(200, 401, (10000,), generator=g).tolist()[:384]        # the first 384 clips of the config-4 list
    mels = [recipes.make_mel(t, i).cuda() for i, t in enumerate(lens)]

    def run_world(world):
        out = {}
        for shard in jatts_b200.shard_utterances(lens, world):
            order = sorted(shard, key=lambda i: (lens[i], i))
            for s in range(0, len(order), 128):
                b = order[s:s + 128]
                for i, y in zip(b, gen.inference_batch([mels[i] for i in b])):
                    out[i] = y
        return out

    base = run_world(1)
    assert sorted(base) == list(range(len(lens)))
    for world in (2, 4, 8):
        got = run_world(world)
        for i in range(len(lens)):
            assert torch.equal(got[i], base[i]), f"world {world}: clip {i} differs from the 1-GPU bytes"
