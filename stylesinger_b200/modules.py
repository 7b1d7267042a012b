"""Reference-named module facades over libstylesinger_b200.so (inference only).

``StyleSinger`` mirrors ``modules/StyleSinger/stylesinger.py:119-187`` (``forward`` keyword arguments, padded
[B, L, ...] tensors in, the ``ret`` dict the reference's callers read out: inference/StyleSinger.py:54-55,
tasks/StyleSinger/stylesinger.py:122-123,190-195).  ``HifiGAN`` mirrors the registered vocoder class
(tasks/tts/vocoder_infer/hifigan_nsf.py:46-75: ``spec2wav(mel np[T,80], f0=np[T]) -> np[T*hop]``).

Only what the ph -> mel -> wav inference path uses is implemented; everything else raises instead of silently
doing something different (training mode, teacher-forced f0/uv, the `forcing` aligner branch of early training
steps, ProDiff / fft decoders).
"""
from typing import Dict, List, Optional

import numpy as np
import torch

from .engine import AcousticModel, PackedBatch, Vocoder, pack_batch
from .hparams import resolve


def padded_to_utterances(txt_tokens, note, note_dur, note_type, spk_embed, emo_embed, ref_mels, ref_f0,
                         mel2ph=None) -> List[Dict[str, torch.Tensor]]:
    """Split the reference's zero-padded batch tensors into per-utterance true-length CPU tensors.

    Padding conventions of the reference collater (tasks/StyleSinger/dataset.py): token id 0 pads ``txt_tokens``
    (and the note tensors alongside), all-zero frames pad ``ref_mels`` (the reference derives its own mask from
    ``ref_mels[:, :, 0] != 0``, lse.py:104, and so does the kernel), 0 pads ``mel2ph``.
    """
    txt_tokens = torch.as_tensor(txt_tokens).cpu()
    B = txt_tokens.shape[0]
    ref_mels = torch.as_tensor(ref_mels).float().cpu()
    ref_f0 = torch.as_tensor(ref_f0).float().cpu()
    if ref_f0.dim() == 1:  # inference/StyleSinger.py passes [R] at B=1
        ref_f0 = ref_f0[None]
    utts = []
    for b in range(B):
        nz = (txt_tokens[b] != 0).nonzero()
        P = int(nz[-1]) + 1 if len(nz) else 0
        if P == 0:
            raise ValueError(f"utterance {b}: empty phone sequence")
        rnz = (ref_mels[b].abs().sum(-1) > 0).nonzero()
        R = int(rnz[-1]) + 1 if len(rnz) else 0
        if R == 0:
            raise ValueError(f"utterance {b}: empty reference mel")
        u = {"txt_tokens": txt_tokens[b, :P].long(), "note": torch.as_tensor(note)[b, :P].long().cpu(),
             "note_dur": torch.as_tensor(note_dur)[b, :P].float().cpu(),
             "note_type": torch.as_tensor(note_type)[b, :P].long().cpu(),
             "spk_embed": torch.as_tensor(spk_embed)[b].float().reshape(-1).cpu(),
             "emo_embed": torch.as_tensor(emo_embed)[b].float().reshape(-1).cpu(),
             "ref_mels": ref_mels[b, :R], "ref_f0": ref_f0[b, :R]}
        if mel2ph is not None:
            m2p = torch.as_tensor(mel2ph)[b].long().cpu()
            fnz = (m2p != 0).nonzero()
            F = int(fnz[-1]) + 1 if len(fnz) else 0
            if F == 0:
                raise ValueError(f"utterance {b}: empty mel2ph")
            u["mel2ph"] = m2p[:F]
        utts.append(u)
    return utts


def packed_to_padded(x: torch.Tensor, offsets, pad_value=0) -> torch.Tensor:
    """Tight [sum_L, ...] rows + host offsets [B+1] -> zero-padded [B, max_L, ...] (the reference's layout)."""
    B = len(offsets) - 1
    lens = [int(offsets[i + 1] - offsets[i]) for i in range(B)]
    out = x.new_full((B, max(lens) if lens else 0) + tuple(x.shape[1:]), pad_value)
    for i in range(B):
        out[i, :lens[i]] = x[int(offsets[i]):int(offsets[i + 1])]
    return out


class StyleSinger:
    """Inference facade with the reference module's call signature (``model(txt_tokens, ..., infer=True)``)."""

    RET_KEYS = ("mel_out", "f0_denorm", "mel2ph", "decoder_inp", "style", "pitch_pred", "spk_embed", "emo_embed",
                "x_mask", "dur")

    def __init__(self, state_dict=None, hparams=None, device=None, engine: Optional[AcousticModel] = None):
        self.hparams = resolve(hparams)
        self.engine = engine if engine is not None else AcousticModel(state_dict, self.hparams, device)
        self.training = False

    def eval(self):
        return self

    def __call__(self, *args, **kwargs):
        return self.forward(*args, **kwargs)

    def forward(self, txt_tokens, mel2ph=None, spk_embed=None, emo_embed=None, ref_mels=None, ref_f0=None,
                f0=None, uv=None, skip_decoder=False, global_steps=0, infer=False, note=None, note_dur=None,
                note_type=None, seed=0, noise=None, **kwargs):
        hp = self.hparams
        if not infer:
            raise NotImplementedError("stylesinger_b200 implements the inference path only (infer=True)")
        if f0 is not None or uv is not None:
            raise NotImplementedError("teacher-forced f0/uv is a training-time input; the inference path predicts pitch")
        if global_steps < hp.get("forcing", 0):
            raise NotImplementedError("global_steps < hparams['forcing'] selects the forced-alignment branch of "
                                      "ProsodyAligner (training warm-up); pass the checkpoint's step count")
        if spk_embed is None or emo_embed is None or ref_mels is None or ref_f0 is None or note is None:
            raise ValueError("spk_embed, emo_embed, ref_mels, ref_f0 and note/note_dur/note_type are required")
        utts = padded_to_utterances(txt_tokens, note, note_dur, note_type, spk_embed, emo_embed, ref_mels, ref_f0, mel2ph)
        pb: PackedBatch = pack_batch(utts, use_mel2ph=mel2ph is not None).to(self.engine.device)
        ret = {}
        dur = None
        if pb.frame_offsets is None:  # FastSpeech2.add_dur at inference (fs2.py:151-174): predicted durations
            dur, _ = self.engine.predict_durations(pb)
            d = dur.cpu().numpy()
            po = pb.ph_offsets
            lens = [int(d[po[i]:po[i + 1]].sum()) for i in range(pb.B)]
            pb.frame_offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
            ret["dur"] = packed_to_padded(dur, po)
        # the reference runs the shallow-diffusion refinement only once training passed diff_start (stylesinger.py:181)
        run_diff = (not skip_decoder) and global_steps > hp.get("diff_start", 0)
        want = ["f0_denorm", "mel2ph", "decoder_inp", "style", "pitch_pred", "spk_proj", "emo_proj"]
        if not skip_decoder:
            want.append("mel_out" if run_diff else "coarse_mel")
        out = self.engine.forward(pb, noise=noise, seed=seed, skip_mel_diffusion=not run_diff, dur=dur, want=tuple(want))
        fo = pb.frame_offsets
        m2p = packed_to_padded(out["mel2ph"].long(), fo)
        ret["mel2ph"] = m2p
        ret["x_mask"] = (m2p > 0).float()[:, :, None]
        ret["f0_denorm"] = packed_to_padded(out["f0_denorm"], fo)
        ret["decoder_inp"] = packed_to_padded(out["decoder_inp"], fo)
        ret["style"] = packed_to_padded(out["style"], fo)
        ret["pitch_pred"] = packed_to_padded(out["pitch_pred"], fo)
        ret["spk_embed"] = out["spk_proj"][:, None, :]
        ret["emo_embed"] = out["emo_proj"][:, None, :]
        if not skip_decoder:
            ret["mel_out"] = packed_to_padded(out["mel_out" if run_diff else "coarse_mel"], fo)
        # training-only entries the callers index unconditionally
        for k in ("gdiff1", "gdiff2", "mdiff1", "mdiff2", "diff", "rq_loss", "gloss"):
            ret[k] = 0.0
        return ret


class HifiGAN:
    """``spec2wav(mel, f0=...)`` with numpy in / numpy out (tasks/tts/vocoder_infer/hifigan_nsf.py:62-75)."""

    def __init__(self, state_dict=None, config=None, device=None, use_nsf=True, engine: Optional[Vocoder] = None):
        self.v = engine if engine is not None else Vocoder(state_dict, config, device)
        self.use_nsf = use_nsf

    def spec2wav(self, mel, **kwargs):
        f0 = kwargs.get("f0")
        dev = self.v.device
        m = torch.as_tensor(np.ascontiguousarray(mel), dtype=torch.float32).to(dev)
        f = None
        if f0 is not None and self.use_nsf:
            f = torch.as_tensor(np.ascontiguousarray(f0), dtype=torch.float32).to(dev)
        offs = np.array([0, m.shape[0]], np.int32)
        wav = self.v.generate(m, f, offs, seed=int(kwargs.get("seed", 0)))
        return wav.cpu().numpy()
