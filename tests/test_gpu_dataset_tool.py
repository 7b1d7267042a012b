"""f4 on the GPU: tools/infer_dataset.py (the engine-driven equivalent of the reference's `--infer` batch path,
tasks/StyleSinger/stylesinger.py:168-275) over the reference's on-disk formats: a work dir with model_ckpt_steps_*.ckpt,
a HiFi-GAN directory (config.yaml + checkpoint) and a binarised IndexedDataset (<prefix>.idx / .data)."""
import os
import pickle
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _items(n):
    rng = np.random.default_rng(0)
    out = []
    for i in range(n):
        Fr, P = 60 + 17 * i, 5 + i
        f0 = rng.uniform(150, 400, Fr).astype(np.float32)
        f0[rng.random(Fr) < 0.2] = 0.0
        m2p = np.repeat(np.arange(1, P + 1), Fr // P + 1)[:Fr]
        out.append({"item_name": f"utt{i}" if i != 2 else None, "mel": np.clip(rng.normal(-3, 0.8, (Fr, 80)), -6, 0.6).astype(np.float32),
                    "f0": f0, "mel2ph": m2p, "ph_token": rng.integers(3, 60, P), "ep_pitches": rng.integers(48, 72, P),
                    "ep_notedurs": rng.uniform(0.1, 0.6, P), "ep_types": rng.integers(1, 3, P),
                    "spk_embed": rng.normal(size=256).astype(np.float32), "emo_embed": rng.normal(size=256).astype(np.float32)})
    return out


def test_infer_dataset_tool_writes_one_wav_per_item(tmp_path):
    import yaml
    from scipy.io import wavfile

    from stylesinger_b200 import synth
    from stylesinger_b200.hparams import DEFAULT_VOCODER_CONFIG, resolve
    T = 4
    hp = resolve(timesteps=T, K_step=T, f0_timesteps=T)
    exp, voc, data, out = (str(tmp_path / d) for d in ("exp", "hifigan", "binary", "out"))
    for d in (exp, voc, data):
        os.makedirs(d)
    torch.save({"state_dict": {"model": synth.acoustic_state_dict(hp, seed=0)}}, os.path.join(exp, "model_ckpt_steps_100.ckpt"))
    torch.save({"state_dict": {"model": {}}}, os.path.join(exp, "model_ckpt_steps_7.ckpt"))  # older step: must be ignored
    torch.save({"state_dict": {"model_gen": synth.vocoder_state_dict(DEFAULT_VOCODER_CONFIG, seed=0)}},
               os.path.join(voc, "model_ckpt_steps_1.ckpt"))
    yaml.safe_dump(dict(DEFAULT_VOCODER_CONFIG), open(os.path.join(voc, "config.yaml"), "w"))
    items = _items(5)
    prefix = os.path.join(data, "test")
    offs = [0]
    with open(prefix + ".data", "wb") as f:
        for it in items:
            offs.append(offs[-1] + f.write(pickle.dumps(it)))
    np.save(open(prefix + ".idx", "wb"), {"offsets": offs})

    sys.path.insert(0, os.path.join(REPO, "tools"))
    import infer_dataset
    for gt in (True, False):
        o = out + ("_gt" if gt else "_pred")
        argv = ["infer_dataset.py", "--ckpt", exp, "--vocoder", voc, "--data", prefix, "--out", o, "--batch", "3", "--T", str(T)]
        sys.argv = argv + (["--use-gt-dur"] if gt else [])
        infer_dataset.main()
        names = sorted(os.listdir(o))
        assert names == sorted(["utt0.wav", "utt1.wav", "item2.wav", "utt3.wav", "utt4.wav"]), names  # fallback name from the dataset index
        for i, it in enumerate(items):
            sr, w = wavfile.read(os.path.join(o, (it["item_name"] or f"item{i}") + ".wav"))
            assert sr == 48000 and w.dtype == np.float32 and np.isfinite(w).all() and np.abs(w).max() <= 1.0
            if gt:  # ground-truth durations: exactly the item's frame count (reference hparam use_gt_dur)
                assert len(w) == len(it["mel2ph"]) * 256
            else:   # predicted durations (the reference default, egs/stylesinger.yaml use_gt_dur: false)
                assert len(w) % 256 == 0 and len(w) > 0
