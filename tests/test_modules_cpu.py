"""Host-side logic of the reference-named facades (stylesinger_b200/modules.py) with a stand-in engine: padded <->
packed conversion, the keys the reference's callers read, and loud refusal of the modes that are out of scope."""
import numpy as np
import pytest
import torch

from stylesinger_b200 import modules as M
from stylesinger_b200 import synth


class FakeEngine:
    """Returns row-index ramps so that the un-packing can be checked exactly."""
    device = torch.device("cpu")

    def __init__(self):
        self.calls = []

    def predict_durations(self, pb):
        P = int(pb.ph_offsets[-1])
        dur = torch.full((P,), 3, dtype=torch.int32)
        return dur, torch.zeros(P)

    def forward(self, pb, noise=None, seed=0, skip_mel_diffusion=False, dur=None, want=()):
        self.calls.append({"skip": skip_mel_diffusion, "want": tuple(want), "dur": dur is not None})
        F = int(pb.frame_offsets[-1])
        ramp = torch.arange(F, dtype=torch.float32)
        out = {}
        for k in want:
            if k in ("mel_out", "coarse_mel"):
                out[k] = ramp[:, None].repeat(1, 80) + (1000.0 if k == "mel_out" else 0.0)
            elif k in ("decoder_inp", "style"):
                out[k] = ramp[:, None].repeat(1, 256)
            elif k == "pitch_pred":
                out[k] = ramp[:, None].repeat(1, 2)
            elif k == "f0_denorm":
                out[k] = ramp + 100.0
            elif k == "mel2ph":
                out[k] = torch.ones(F, dtype=torch.int32)
            elif k in ("spk_proj", "emo_proj"):
                out[k] = torch.ones(pb.B, 256)
            else:
                raise AssertionError(k)
        return out


def _padded_batch(frames=(40, 25), phones=(6, 4), ref=(30, 20)):
    us = [synth.make_utterance(frames[i] / 187.5, utt_idx=10 + i, ref_frames=ref[i], frames=frames[i], phones=phones[i])
          for i in range(len(frames))]
    pad = torch.nn.utils.rnn.pad_sequence
    b = {k: pad([u[k] for u in us], batch_first=True) for k in ("txt_tokens", "note", "note_dur", "note_type", "mel2ph",
                                                                "ref_mels", "ref_f0")}
    b["spk_embed"] = torch.stack([u["spk_embed"] for u in us])
    b["emo_embed"] = torch.stack([u["emo_embed"] for u in us])
    return us, b


def test_padded_to_utterances_recovers_true_lengths():
    us, b = _padded_batch()
    got = M.padded_to_utterances(b["txt_tokens"], b["note"], b["note_dur"], b["note_type"], b["spk_embed"], b["emo_embed"],
                                 b["ref_mels"], b["ref_f0"], b["mel2ph"])
    for u, g in zip(us, got):
        for k in ("txt_tokens", "note", "note_dur", "note_type", "mel2ph", "ref_mels", "ref_f0", "spk_embed", "emo_embed"):
            assert torch.equal(torch.as_tensor(u[k]).to(g[k].dtype), g[k]), k


def test_packed_to_padded_round_trip():
    x = torch.arange(7 * 3, dtype=torch.float32).reshape(7, 3)
    p = M.packed_to_padded(x, np.array([0, 2, 7], np.int32))
    assert p.shape == (2, 5, 3)
    assert torch.equal(p[0, :2], x[:2]) and torch.equal(p[1], x[2:]) and float(p[0, 2:].abs().sum()) == 0.0


def test_forward_returns_the_reference_ret_keys_with_padded_layout():
    us, b = _padded_batch()
    eng = FakeEngine()
    m = M.StyleSinger(engine=eng)
    ret = m(b["txt_tokens"], mel2ph=b["mel2ph"], spk_embed=b["spk_embed"], emo_embed=b["emo_embed"], ref_mels=b["ref_mels"],
            ref_f0=b["ref_f0"], global_steps=320000, infer=True, note=b["note"], note_dur=b["note_dur"], note_type=b["note_type"])
    for k in ("mel_out", "f0_denorm", "mel2ph", "decoder_inp", "style", "pitch_pred", "spk_embed", "emo_embed", "x_mask",
              "gdiff1", "gdiff2", "mdiff1", "mdiff2", "diff", "rq_loss", "gloss"):
        assert k in ret, k
    assert ret["mel_out"].shape == (2, 40, 80) and ret["f0_denorm"].shape == (2, 40)
    assert ret["spk_embed"].shape == (2, 1, 256) and ret["x_mask"].shape == (2, 40, 1)
    # utterance 1 owns packed rows 40..64: ramp values survive, the padding stays zero
    assert float(ret["f0_denorm"][1, 0]) == 140.0 and float(ret["f0_denorm"][1, 24]) == 164.0
    assert float(ret["f0_denorm"][1, 25:].abs().sum()) == 0.0 and float(ret["x_mask"][1, 25:].sum()) == 0.0
    assert float(ret["mel_out"][0, 0, 0]) == 1000.0  # post-diffusion mel when global_steps > diff_start
    assert eng.calls[-1]["skip"] is False and eng.calls[-1]["dur"] is False


def test_forward_before_diff_start_returns_the_coarse_mel_and_predicts_durations():
    us, b = _padded_batch()
    eng = FakeEngine()
    m = M.StyleSinger(engine=eng)
    ret = m(b["txt_tokens"], spk_embed=b["spk_embed"], emo_embed=b["emo_embed"], ref_mels=b["ref_mels"], ref_f0=b["ref_f0"],
            global_steps=50000, infer=True, note=b["note"], note_dur=b["note_dur"], note_type=b["note_type"])
    assert eng.calls[-1]["skip"] is True and eng.calls[-1]["dur"] is True and "coarse_mel" in eng.calls[-1]["want"]
    assert ret["dur"].shape == (2, 6) and int(ret["dur"][1, 4:].sum()) == 0
    assert ret["mel_out"].shape == (2, 18, 80) and float(ret["mel_out"][0, 0, 0]) == 0.0  # 6 phones x 3 frames


def test_out_of_scope_modes_fail_loudly():
    us, b = _padded_batch()
    m = M.StyleSinger(engine=FakeEngine())
    kw = dict(spk_embed=b["spk_embed"], emo_embed=b["emo_embed"], ref_mels=b["ref_mels"], ref_f0=b["ref_f0"], note=b["note"],
              note_dur=b["note_dur"], note_type=b["note_type"])
    with pytest.raises(NotImplementedError):
        m(b["txt_tokens"], infer=False, global_steps=320000, **kw)
    with pytest.raises(NotImplementedError):
        m(b["txt_tokens"], infer=True, global_steps=320000, f0=torch.zeros(2, 40), **kw)
    with pytest.raises(NotImplementedError):
        m(b["txt_tokens"], infer=True, global_steps=100, **kw)
    with pytest.raises(ValueError):
        m(b["txt_tokens"], infer=True, global_steps=320000)


def test_hifigan_facade_passes_numpy_through():
    class FakeVoc:
        device = torch.device("cpu")
        hop = 256

        def generate(self, mel, f0, offs, seed=0, **kw):
            assert mel.shape[1] == 80 and (f0 is None or f0.shape[0] == mel.shape[0]) and list(offs) == [0, mel.shape[0]]
            return torch.zeros(mel.shape[0] * self.hop) + (0.0 if f0 is None else 1.0)

    v = M.HifiGAN(engine=FakeVoc())
    w = v.spec2wav(np.zeros((7, 80), np.float32), f0=np.ones(7, np.float32))
    assert isinstance(w, np.ndarray) and w.shape == (7 * 256,) and float(w[0]) == 1.0
    assert float(M.HifiGAN(engine=FakeVoc(), use_nsf=False).spec2wav(np.zeros((3, 80), np.float32), f0=np.ones(3))[0]) == 0.0
