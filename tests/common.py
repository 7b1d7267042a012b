"""Shared helpers for the test-suite (oracle side).  Tests are the only importers of oracle/."""
import json
import os

import numpy as np
import torch

from oracle import stylesinger_oracle as O
from stylesinger_b200 import synth
from stylesinger_b200.hparams import DEFAULT_VOCODER_CONFIG, resolve

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_CACHE = {}


def golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    return g, json.loads(str(g["meta"]))


def hp_for(T, f0_T=None):
    return resolve(timesteps=T, K_step=T, f0_timesteps=T if f0_T is None else f0_T)


def acoustic_sd():
    if "sd" not in _CACHE:
        _CACHE["sd"] = synth.acoustic_state_dict(hp_for(4), seed=0)
    return _CACHE["sd"]


def vocoder_sd():
    if "vsd" not in _CACHE:
        _CACHE["vsd"] = synth.vocoder_state_dict(DEFAULT_VOCODER_CONFIG, seed=0)
    return _CACHE["vsd"]


def utt_from_meta(meta):
    return synth.make_utterance(meta["frames"] / 187.5, utt_idx=meta["utt_idx"], ref_frames=meta["ref_frames"],
                                frames=meta["frames"], phones=meta["phones"])


def oracle_forward(u, hp, seed, use_mel2ph=True, **kw):
    ns = O.NoiseSource(seed)
    with torch.no_grad():
        r = O.stylesinger_forward(acoustic_sd(), hp, u["txt_tokens"][None], u["note"][None], u["note_dur"][None],
                                  u["note_type"][None], u["spk_embed"][None], u["emo_embed"][None],
                                  u["ref_mels"][None], u["ref_f0"], ns,
                                  mel2ph=u["mel2ph"][None] if use_mel2ph else None, **kw)
    return r, ns
