"""f3 front-end oracle (oracle/frontend_oracle.py) against what can be checked without librosa: scipy's STFT and the defining
properties of the Slaney mel filterbank."""
import numpy as np
import pytest
import scipy.signal

from oracle import frontend_oracle as FO


def test_stft_matches_scipy():
    rng = np.random.default_rng(0)
    for n in (48000, 12345, 1024):  # (scipy refuses signals shorter than the window; librosa pads them)
        y = rng.standard_normal(n).astype(np.float32)
        S = FO.stft(y, 1024, 256, 1024)
        assert S.shape == (513, 1 + n // 256)
        w = FO.hann_periodic(1024)
        _, _, Z = scipy.signal.stft(y.astype(np.float64), window=w, nperseg=1024, noverlap=768, nfft=1024, boundary="zeros",
                                    padded=False, return_onesided=True)
        Z = Z * w.sum()  # scipy scales by 1 / sum(window)
        T = min(Z.shape[1], S.shape[1])  # scipy drops the last partial hop
        assert T >= S.shape[1] - 1
        scale = np.abs(Z[:, :T]).max()
        assert np.abs(Z[:, :T] - S[:, :T]).max() < 2e-6 * scale


def test_mel_basis_is_the_slaney_construction():
    sr, n_fft, n_mels, fmin, fmax = 48000, 1024, 80, 20, 24000
    B = FO.mel_basis(sr, n_fft, n_mels, fmin, fmax)
    assert B.shape == (80, 513) and B.dtype == np.float32 and (B >= 0).all()
    edges = FO.mel_to_hz(np.linspace(FO.hz_to_mel(fmin), FO.hz_to_mel(fmax), n_mels + 2))
    assert abs(edges[0] - fmin) < 1e-9 and abs(edges[-1] - fmax) < 1e-6
    assert abs(float(FO.hz_to_mel(1000.0)) - 15.0) < 1e-12            # Slaney: 1 kHz = 15 mel, linear below
    assert abs(float(FO.hz_to_mel(6400.0)) - 42.0) < 1e-9             # 27 log-spaced steps up to 6.4 kHz
    freqs = np.linspace(0, sr / 2, 513)
    enorm = 2.0 / (edges[2:] - edges[:-2])
    tri = B / enorm[:, None]                                          # un-normalised triangles
    for i in (0, 17, 40, 79):
        inside = (freqs > edges[i]) & (freqs < edges[i + 2])
        assert (tri[i][~inside] == 0).all() and tri[i].max() <= 1.0 + 1e-6
        peak = freqs[np.argmax(tri[i])]
        assert abs(peak - edges[i + 1]) <= sr / n_fft                 # apex at the centre frequency (to one bin)
    # neighbouring triangles sum to one between the first and the last centre frequency
    mid = (freqs >= edges[1]) & (freqs <= edges[-2])
    assert np.abs(tri.sum(0)[mid] - 1.0).max() < 1e-5


def test_wav2mel_shape_and_floor():
    y = np.zeros(5000, np.float32)
    m = FO.wav2mel(y)
    assert m.shape == (1 + 5000 // 256, 80) and np.allclose(m, -6.0)  # log10(eps)


# ---- emotion-encoder half: restatements pinned by the unmodified reference (tests/golden/ref_emotion_encoder.npz) -------------
def _emo_golden():
    import os
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_emotion_encoder.npz"))


def test_lstm_oracle_matches_the_reference_encoder():
    g = _emo_golden()
    sd = FO.emotion_encoder_weights(int(g["seed"]))
    hidden = FO.lstm_hidden(g["frames"], sd)
    assert hidden.shape == g["hidden"].shape == (5, 256)
    assert np.abs(hidden - g["hidden"]).max() < 2e-7          # torch fp32 LSTM vs this float64 restatement
    assert np.abs(FO.emotion_embeds(hidden, sd) - g["embeds"]).max() < 5e-7
    assert np.abs(FO.utterance_embed(g["hidden"]) - g["utt_embed"]).max() < 1e-7
    h2 = FO.lstm_hidden(g["frames2"], sd)                     # 3 sequences of 97 frames (using_partials=False shape)
    assert h2.shape == (3, 256) and np.abs(h2 - g["hidden2"]).max() < 5e-7


def test_partial_slices_match_the_reference_and_the_host_mirror():
    from stylesinger_b200 import emotion
    g = _emo_golden()["slices"]
    for n in np.unique(g[:, 0]):
        want = g[g[:, 0] == n][:, 1:]
        wav, mel = FO.compute_partial_slices(int(n))
        assert np.array_equal(np.array([[a, b, c, d] for (a, b), (c, d) in zip(wav, mel)]), want)
        ws, ms = emotion.compute_partial_slices(int(n))
        assert np.array_equal(np.array([[w.start, w.stop, m.start, m.stop] for w, m in zip(ws, ms)]), want)


def test_reflect_stft_matches_scipy_and_emotion_mel_shape():
    rng = np.random.default_rng(3)
    y = rng.standard_normal(16000).astype(np.float32)
    S = FO.stft(y, 400, 160, 400, pad_mode="reflect")
    w = FO.hann_periodic(400)
    _, _, Z = scipy.signal.stft(y.astype(np.float64), window=w, nperseg=400, noverlap=240, nfft=400, boundary="even", padded=False)
    Z = Z * w.sum()
    T = min(Z.shape[1], S.shape[1])
    assert T >= S.shape[1] - 1 and np.abs(Z[:, :T] - S[:, :T]).max() < 2e-6 * np.abs(Z).max()
    m = FO.emotion_mel(y)
    assert m.shape == (101, 40) and m.dtype == np.float32 and (m >= 0).all()


def test_emotion_module_refuses_to_run_without_a_model():
    from stylesinger_b200 import emotion
    assert not emotion.is_loaded()
    with pytest.raises(Exception, match="Model was not loaded"):
        emotion.embed_frames_batch(np.zeros((1, 160, 40), np.float32))


def test_embed_utterance_glue_with_stand_in_kernels(monkeypatch):
    """emotion.embed_utterance host flow (padding, partial batch, return values) with the oracle standing in for the two CUDA
    objects: must equal the reference's composition of the same pieces (inference.py:110-155)."""
    import torch
    from stylesinger_b200 import emotion
    sd = FO.emotion_encoder_weights(3)

    def fake_mel(wav):
        return torch.from_numpy(FO.emotion_mel(np.asarray(wav, np.float32)))

    def fake_model(frames, utt_offsets=None, want_embeds=False):
        hid = FO.lstm_hidden(np.asarray(frames), sd)
        out = {"hidden": torch.from_numpy(hid)}
        if utt_offsets is not None:
            out["utt_embed"] = torch.from_numpy(np.stack([FO.utterance_embed(hid[a:b]) for a, b in zip(utt_offsets[:-1], utt_offsets[1:])]))
        return out

    monkeypatch.setattr(emotion, "_mel", fake_mel)
    monkeypatch.setattr(emotion, "_model", fake_model)
    rng = np.random.default_rng(5)
    for n in (3000, 30000):
        y = (0.1 * rng.standard_normal(n)).astype(np.float32)
        wav_sl, mel_sl = FO.compute_partial_slices(n)
        fr = FO.emotion_mel(np.pad(y, (0, max(0, wav_sl[-1][1] - n))))
        hid = FO.lstm_hidden(np.stack([fr[a:b] for a, b in mel_sl]), sd)
        emb, partials, wave_slices = emotion.embed_utterance(y, return_partials=True)
        assert np.allclose(emb, FO.utterance_embed(hid), atol=1e-7) and np.allclose(partials, hid, atol=1e-7)
        assert [(s.start, s.stop) for s in wave_slices] == wav_sl
        assert np.allclose(emotion.embed_utterance(y), emb)
        whole, p2, s2 = emotion.embed_utterance(y, using_partials=False, return_partials=True)
        assert p2 is None and s2 is None and np.allclose(whole, FO.lstm_hidden(FO.emotion_mel(y)[None], sd)[0], atol=1e-7)


def test_voice_encoder_mirror_slicing_and_flow_with_stand_in_kernels():
    """stylesinger_b200.voice_encoder (resemblyzer's VoiceEncoder on this package's kernels; third party, unpinned): the slicing
    against the oracle's loop formulation, and embed_utterance's flow with the oracle standing in for the CUDA objects."""
    import torch
    from stylesinger_b200.voice_encoder import VoiceEncoder
    for n in (1, 159, 8000, 25599, 25600, 37919, 37920, 38000, 48000, 160000, 479999):
        for rate, cov in ((1.3, 0.75), (2.0, 0.5), (0.7, 1.0)):
            ws, ms = VoiceEncoder.compute_partial_slices(n, rate, cov)
            want = FO.resemblyzer_partial_slices(n, rate, cov)
            assert ([(s.start, s.stop) for s in ws], [(s.start, s.stop) for s in ms]) == want
    with pytest.raises(AssertionError):
        VoiceEncoder.compute_partial_slices(16000, 0.5, 0.75)    # fewer than one partial per 1.6 s
    sd = FO.emotion_encoder_weights(9)
    ve = VoiceEncoder.__new__(VoiceEncoder)
    ve._mel = lambda wav: torch.from_numpy(FO.emotion_mel(np.asarray(wav, np.float32)))
    ve._net = lambda mels, want_embeds=False: {"embeds": torch.from_numpy(FO.emotion_embeds(FO.lstm_hidden(np.asarray(mels), sd), sd))}
    rng = np.random.default_rng(2)
    y = (0.1 * rng.standard_normal(40000)).astype(np.float32)
    emb, partials, wav_slices = ve.embed_utterance(y, return_partials=True)
    wav_sl, mel_sl = FO.resemblyzer_partial_slices(len(y))
    fr = FO.emotion_mel(np.pad(y, (0, max(0, wav_sl[-1][1] - len(y)))))
    pe = FO.emotion_embeds(FO.lstm_hidden(np.stack([fr[a:b] for a, b in mel_sl]), sd), sd)
    raw = pe.mean(0)
    assert np.allclose(partials, pe, atol=1e-7) and np.allclose(emb, raw / np.linalg.norm(raw), atol=1e-7)
    assert abs(np.linalg.norm(emb) - 1) < 1e-6 and [(s.start, s.stop) for s in wav_slices] == wav_sl
