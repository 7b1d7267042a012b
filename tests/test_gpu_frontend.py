"""f3 (mel half): the CUDA front-end (ssb_melspec_*) against the numpy restatement of librosa 0.8.0 in oracle/frontend_oracle.py.

Tolerance: |log10 mel| differences < 5e-3 everywhere and < 2e-4 on average for signals whose quietest band is >= 50 dB below
the loudest one (the DFT runs as one fp32 GEMM with K = 1024: its rounding is relative to the frame's total energy, the
reference's float64 FFT rounds per bin), exact -6.0 floor on silence.
"""
import numpy as np
import pytest
import torch

from oracle import frontend_oracle as FO

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _signals():
    rng = np.random.default_rng(7)
    sr = 48000
    out = []
    for n in (48000 * 6, 12345, 1024, 255, 70000):  # 6 s reference clip (R = 1125 frames), ragged, shorter than a window
        t = np.arange(n) / sr
        f0 = 180 + 120 * np.sin(2 * np.pi * 0.7 * t)
        y = sum(0.25 / k * np.sin(2 * np.pi * k * np.cumsum(f0) / sr) for k in range(1, 9))
        y = y + 0.01 * rng.standard_normal(n)
        out.append(y.astype(np.float32))
    return out


def test_melspec_matches_oracle_on_a_ragged_batch():
    from stylesinger_b200.engine import MelSpectrogram
    fe = MelSpectrogram(device=DEV)
    wavs = _signals()
    mels = fe(wavs)
    assert [tuple(m.shape) for m in mels] == [(1 + len(w) // 256, 80) for w in wavs]
    worst, mean = 0.0, []
    for w, m in zip(wavs, mels):
        ref = FO.wav2mel(w)
        d = np.abs(m.cpu().numpy() - ref)
        worst = max(worst, float(d.max()))
        mean.append(float(d.mean()))
    print(f"melspec vs oracle: max |d log10 mel| {worst:.3e}, mean {np.mean(mean):.3e}")
    assert worst < 5e-3 and np.mean(mean) < 2e-4
    # batch composition does not matter (B = 1 semantics)
    solo = fe(wavs[1])
    assert torch.equal(solo, mels[1])


def test_melspec_silence_hits_the_floor_and_process_audio_mirrors_the_reference():
    from stylesinger_b200 import synth
    from stylesinger_b200.hparams import DEFAULT_VOCODER_CONFIG, resolve
    from stylesinger_b200.infer import StyleSingerInfer
    hp = resolve(timesteps=4, K_step=4, f0_timesteps=4)
    eng = StyleSingerInfer(hp, DEV, synth.acoustic_state_dict(hp, seed=0), synth.vocoder_state_dict(DEFAULT_VOCODER_CONFIG, seed=0),
                           DEFAULT_VOCODER_CONFIG)
    wav, mel = eng.process_audio(np.zeros(5000, np.float32))
    assert mel.shape == (1 + 5000 // 256, 80) and np.allclose(mel, -6.0) and wav.dtype == np.float16 and len(wav) == mel.shape[0] * 256
    y = _signals()[3]
    wav, mel = eng.process_audio(y)
    assert np.abs(mel - FO.wav2mel(y)).max() < 5e-3
