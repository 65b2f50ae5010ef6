"""GPU tests of the batched stage-4 front-end (jatts_b200/decode.py): fused PCM_16 output and the command-line
path with the recipe's file formats (csv, tokens.txt, config.yml, checkpoint .pkl, stats), checked against the
per-utterance reference-shaped calls ``model.inference`` -> ``vocoder.decode`` -> PCM_16 rounding."""
import wave

import numpy as np
import pytest
import torch
import yaml

import jatts_b200
from h5_writer import write_h5
from jatts_b200 import decode
from oracle import fs2 as ofs2
from oracle import hifigan as ohg
from oracle import recipes


def pcm16_of(y: torch.Tensor) -> torch.Tensor:
    """what libsndfile stores for float data written as PCM_16: lrintf(y * 0x7FFF) (round half to even)"""
    return torch.round(y.float() * 32767.0).to(torch.int16)


@pytest.mark.gpu
def test_pcm16_is_the_rounded_float_waveform():
    cfg = recipes.HIFIGAN_TINY
    g = jatts_b200.HiFiGANGenerator(**cfg)
    g.load_state_dict(recipes.make_hifigan_state_dict(cfg, 0))
    g = g.eval().to("cuda")
    mels = [recipes.make_mel(t, i) for i, t in enumerate([33, 1, 80])]
    ys = g.inference_batch(mels)
    ps = g.inference_batch(mels, pcm16=True)
    for y, p in zip(ys, ps):
        assert p.dtype == torch.int16 and p.shape == y.shape
        assert torch.equal(p.cpu(), pcm16_of(y.cpu()))          # bit exact: same kernel, same accumulator


@pytest.mark.gpu
def test_cli_writes_the_per_utterance_samples(tmp_path):
    cfg, hcfg = recipes.JSUT_FS2, recipes.HIFIGAN_TINY
    sd = recipes.make_fs2_state_dict(cfg, seed=0, duration_recipe="A")
    hsd = recipes.make_hifigan_state_dict(hcfg, seed=0)
    tstats, vstats = recipes.make_stats(1), recipes.make_stats(2)
    # ---- the recipe's files
    vocab = ["<blank>", "<unk>"] + [f"p{i}" for i in range(2, cfg["idim"] - 1)] + ["<sos/eos>"]
    (tmp_path / "tokens.txt").write_text("\n".join(vocab) + "\n", encoding="utf-8")
    lens = [9, 31, 4, 17, 9, 50, 1]
    rows = ["sample_id,phonemes"]
    texts = []
    for i, n in enumerate(lens):
        ids = recipes.make_phonemes(n, 300 + i, cfg["idim"]).tolist()
        toks = [vocab[t] for t in ids]
        if i == 3:
            toks[2] = "not-in-vocab"        # -> <unk> (id 1)
            ids[2] = 1
        texts.append(torch.tensor(ids, dtype=torch.long))
        rows.append(f"utt{i}," + " ".join(toks))
    (tmp_path / "dev.csv").write_text("\n".join(rows) + "\n", encoding="utf-8")
    # the recipe's statistics files are HDF5 (compute_statistics.py / parallel_wavegan): read without h5py
    write_h5(tmp_path / "stats.h5", {"mel_mean": tstats["mean"].numpy(), "mel_scale": tstats["scale"].numpy()})
    write_h5(tmp_path / "voc_stats.h5", {"mean": vstats["mean"].numpy(), "scale": vstats["scale"].numpy()})
    torch.save({"model": {"generator": hsd}}, tmp_path / "voc.pkl")
    with open(tmp_path / "voc_config.yml", "w") as f:
        yaml.safe_dump({"generator_type": "HiFiGANGenerator", "sampling_rate": 24000,
                        "generator_params": {k: (list(map(list, v)) if k == "resblock_dilations" else list(v) if isinstance(v, tuple) else v)
                                             for k, v in hcfg.items()}}, f)
    torch.save({"model": sd}, tmp_path / "checkpoint-1steps.pkl")
    with open(tmp_path / "config.yml", "w") as f:
        yaml.safe_dump({"model_type": "FastSpeech2", "model_params": dict(cfg), "out_feat_type": "mel", "feat_list": ["mel"],
                        "sampling_rate": 24000,
                        "vocoder": {"checkpoint": str(tmp_path / "voc.pkl"), "config": str(tmp_path / "voc_config.yml"),
                                    "stats": str(tmp_path / "voc_stats.h5")}}, f)
    out = tmp_path / "out"
    rc = decode.main(["--csv", str(tmp_path / "dev.csv"), "--stats", str(tmp_path / "stats.h5"),
                      "--token-list", str(tmp_path / "tokens.txt"), "--token-column", "phonemes", "--outdir", str(out),
                      "--checkpoint", str(tmp_path / "checkpoint-1steps.pkl"), "--max-utts", "3", "--verbose", "0"])
    assert rc == 0
    # ---- the same utterances one at a time through the reference-shaped calls
    model = jatts_b200.FastSpeech2(**cfg)
    model.load_state_dict(sd)
    model = model.eval().to("cuda")
    voc = jatts_b200.Vocoder(hsd, {"generator_type": "HiFiGANGenerator", "generator_params": dict(hcfg), "sampling_rate": 24000},
                             vstats, "cuda", trg_stats=tstats)
    for i, x in enumerate(texts):
        y, sr = voc.decode(model.inference(x.to("cuda"))["feat_gen"])
        with wave.open(str(out / "wav" / f"utt{i}.wav"), "rb") as w:
            assert w.getframerate() == sr == 24000 and w.getnchannels() == 1 and w.getsampwidth() == 2
            got = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2")
        assert np.array_equal(got, pcm16_of(y.cpu()).numpy()), f"utt{i}: wav samples differ from the per-utterance path"

    # ---- and against the CPU oracle (the reference's arithmetic): same number of samples (durations are bit-exact),
    #      waveform within the stated AC-SNR bound; PCM_16 quantisation (-98 dB) is far below the bf16 noise floor
    torch.set_num_threads(__import__("os").cpu_count() or 1)
    for i in (0, 1, 3, 6):
        ref = ofs2.fs2_inference(sd, cfg, texts[i])
        yref = ohg.vocoder_decode(hsd, hcfg, ref["feat_gen"], vstats, tstats)
        with wave.open(str(out / "wav" / f"utt{i}.wav"), "rb") as w:
            got = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").astype(np.float32) / 32767.0
        assert got.shape[0] == yref.shape[0] == int(ref["duration"].sum()) * 6
        snr = ohg.ac_snr_db(yref, torch.from_numpy(got))
        assert snr > 35.0, f"utt{i}: AC-SNR {snr:.1f} dB vs the oracle"

    # ---- two processes' worth of sharding (one per GPU in production: torchrun sets RANK / WORLD_SIZE): every file is
    #      written by exactly one rank and holds the bytes of the single-process run
    out2 = tmp_path / "out2"
    for r in range(2):
        assert decode.main(["--csv", str(tmp_path / "dev.csv"), "--stats", str(tmp_path / "stats.h5"), "--token-list",
                            str(tmp_path / "tokens.txt"), "--token-column", "phonemes", "--outdir", str(out2), "--checkpoint",
                            str(tmp_path / "checkpoint-1steps.pkl"), "--verbose", "0", "--rank", str(r), "--world-size", "2"]) == 0
    import json
    counts = [json.load(open(out2 / f"decode_stats.rank{r}.json"))["utterances"] for r in range(2)]
    assert sum(counts) == len(lens) and min(counts) >= 1
    for i in range(len(lens)):
        assert (out2 / "wav" / f"utt{i}.wav").read_bytes() == (out / "wav" / f"utt{i}.wav").read_bytes()


@pytest.mark.gpu
def test_utterance_longer_than_max_len_is_reported_not_fatal(tmp_path):
    """one over-long utterance used to abort the whole run part-way; now the rest of its batch is decoded
    and the offender is reported (the reference itself has no length limit: --max-len raises ours to 5000 frames)."""
    cfg, hcfg = recipes.JSUT_FS2, recipes.HIFIGAN_TINY
    model = jatts_b200.FastSpeech2(**cfg, max_len=128)
    model.load_state_dict(recipes.make_fs2_state_dict(cfg, seed=0, duration_recipe="A"))
    model = model.eval().to("cuda")
    st = recipes.make_stats(1)
    voc = jatts_b200.Vocoder(recipes.make_hifigan_state_dict(hcfg, 0), {"generator_type": "HiFiGANGenerator",
                             "generator_params": dict(hcfg), "sampling_rate": 24000}, st, "cuda", trg_stats=st)
    items = [{"sample_id": f"u{i}", "token_indices": recipes.make_phonemes(n, 50 + i, cfg["idim"]).tolist()}
             for i, n in enumerate([10, 12, 40, 11])]          # 40 tokens -> ~240 frames > 128
    res = decode.decode_items(model, voc, items, str(tmp_path), 24000, torch.device("cuda"), max_utts=8)
    assert res["skipped"] == ["u2"] and res["files"] == 3
    assert sorted(p.name for p in (tmp_path / "wav").iterdir()) == ["u0.wav", "u1.wav", "u3.wav"]


@pytest.mark.gpu
def test_matcha_through_the_stage4_front_end(tmp_path):
    """BASELINE config 5 through the same front-end (tts_decode.py:216-226 passes ``temperature`` / ``n_timesteps``):
    one wav per utterance, hop x (even-truncated duration sum) samples, and -- the noise being drawn from the seeded CUDA
    generator exactly once per batch -- the samples the ORACLE predicts from that noise."""
    from oracle import matcha as om

    cfg, hcfg = recipes.SMALL_MATCHA, dict(recipes.HIFIGAN_TINY, in_channels=recipes.SMALL_MATCHA["odim"])
    sd = recipes.make_matcha_state_dict(cfg, 2)
    model = jatts_b200.MatchaTTS(**cfg)
    model.load_state_dict(sd)
    model = model.eval().to("cuda")
    tstats, vstats = recipes.make_stats(3, cfg["odim"]), recipes.make_stats(4, cfg["odim"])
    hsd = recipes.make_hifigan_state_dict(hcfg, 1)
    voc = jatts_b200.Vocoder(hsd, {"generator_type": "HiFiGANGenerator", "generator_params": dict(hcfg), "sampling_rate": 24000},
                             vstats, "cuda", trg_stats=tstats)
    lens = [6, 14, 9]
    items = [{"sample_id": f"m{i}", "token_indices": recipes.make_phonemes(n, 70 + i, cfg["idim"]).tolist()} for i, n in enumerate(lens)]
    kw = {"n_timesteps": 4, "temperature": 0.667}
    torch.manual_seed(1234)
    res = decode.decode_items(model, voc, items, str(tmp_path), 24000, torch.device("cuda"), max_utts=8, inference_kwargs=kw)
    assert res["files"] == 3 and not res["skipped"]
    # the front-end sorts by length into ONE batch here: reproduce its noise draw (matchatts.py: torch.randn(sum T, odim))
    order = sorted(range(len(lens)), key=lambda j: lens[j])
    texts = [torch.tensor(items[j]["token_indices"]) for j in order]
    frames = [int(om.matcha_inference(sd, cfg, x, torch.zeros(cfg["odim"], 4096), 1, 0.0)["feat_gen"].shape[0]) for x in texts]
    torch.manual_seed(1234)
    z = torch.randn(sum(frames), cfg["odim"], device="cuda").cpu()
    hop = 6
    o = 0
    for x, j, f in zip(texts, order, frames):
        ref = om.matcha_inference(sd, cfg, x, z[o:o + f].t(), kw["n_timesteps"], kw["temperature"])
        o += f
        yref = ohg.vocoder_decode(hsd, hcfg, ref["feat_gen"], vstats, tstats)
        with wave.open(str(tmp_path / "wav" / f"m{j}.wav"), "rb") as w:
            got = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").astype(np.float32) / 32767.0
        tot = int(ref["duration"].sum())
        assert got.shape[0] == yref.shape[0] == (tot - tot % 2) * hop
        snr = ohg.ac_snr_db(yref, torch.from_numpy(got))
        assert snr > 35.0, f"m{j}: AC-SNR {snr:.1f} dB vs the oracle"
