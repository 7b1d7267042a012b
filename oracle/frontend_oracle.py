"""ORACLE — test infrastructure only.  NOT part of the product path.

f3, mel-spectrogram half of the reference-audio front-end: a numpy restatement of
`librosa_wav2spec` (reference utils/audios/__init__.py:36-84, called by inference/StyleSinger.py:79-92).

The arithmetic of that function lives in a THIRD-PARTY dependency that is not vendored in /root/reference and not installed
in this image: librosa==0.8.0 (requirements.txt:2).  This file restates the published algorithms of librosa 0.8.0
(`librosa.core.spectrum.stft`, `librosa.filters.mel`, `librosa.util.pad_center`) and follows the reference's call site for
everything around them.  PARITY UNPINNED against librosa itself: what can be checked here is checked in
tests/test_frontend_cpu.py - the STFT against scipy.signal.stft (independent implementation, same framing / padding
convention), the filterbank against the defining properties of the Slaney construction (triangles on the Slaney mel scale,
area normalisation 2 / (f[i+2] - f[i]), partition of unity before normalisation).
"""
import numpy as np


def hann_periodic(n):
    """scipy.signal.get_window('hann', n, fftbins=True), which librosa.filters.get_window forwards to."""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def pad_center(w, size):
    """librosa.util.pad_center: zero-pad to `size`, centred (left pad = (size - len) // 2)."""
    lpad = (size - len(w)) // 2
    return np.pad(w, (lpad, size - len(w) - lpad))


def stft(y, n_fft, hop_length, win_length, pad_mode="constant"):
    """librosa.stft(y, n_fft, hop_length, win_length, window='hann', center=True, pad_mode=pad_mode) -> complex64 [1 + n_fft/2, T].
    (librosa_wav2spec passes pad_mode='constant'; librosa's own default, used by melspectrogram, is 'reflect'.)
    librosa multiplies the float64 window into the float32 frames (-> float64), runs the real FFT and stores complex64."""
    y = np.asarray(y, dtype=np.float32)
    w = pad_center(hann_periodic(win_length), n_fft).reshape(-1, 1)
    yp = np.pad(y, n_fft // 2, mode=pad_mode)
    n_frames = 1 + (len(yp) - n_fft) // hop_length
    idx = np.arange(n_fft)[:, None] + hop_length * np.arange(n_frames)[None, :]
    frames = yp[idx]  # [n_fft, T]
    return np.fft.rfft(w * frames, axis=0).astype(np.complex64)


def hz_to_mel(f):
    """librosa.hz_to_mel(htk=False): Slaney's Auditory Toolbox scale (linear below 1 kHz, log above)."""
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz, logstep = 1000.0, np.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)


def mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz, logstep = 1000.0, np.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_basis(sr, n_fft, n_mels, fmin, fmax):
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax, htk=False, norm='slaney', dtype=float32) -> [n_mels, 1 + n_fft/2]."""
    fftfreqs = np.linspace(0, float(sr) / 2, int(1 + n_fft // 2), endpoint=True)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((n_mels, int(1 + n_fft // 2)), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights.astype(np.float32)


def wav2mel(wav, fft_size=1024, hop_size=256, win_length=1024, num_mels=80, fmin=20, fmax=24000, eps=1e-6, sample_rate=48000):
    """The 'mel' entry of librosa_wav2spec (utils/audios/__init__.py:57-77), transposed to [T, num_mels] like its return value.
    loud_norm / trim_long_sil are false in the reference's configuration (egs/egs_bases/tts/base.yaml:39) and not restated."""
    x_stft = stft(wav, fft_size, hop_size, win_length)
    linear_spc = np.abs(x_stft)                                   # :59
    fmin = 0 if fmin == -1 else fmin                              # :62
    fmax = sample_rate / 2 if fmax == -1 else fmax                # :63
    basis = mel_basis(sample_rate, fft_size, num_mels, fmin, fmax)  # :66
    mel = basis @ linear_spc                                      # :69
    mel = np.log10(np.maximum(eps, mel))                          # :70
    return mel.T.astype(np.float32)


# ------------------------------------------------------------------------------------------------------------------------
# f3, emotion-encoder half: the reference's OWN network (data_gen/tts/emotion/model.py:10-77) and the utterance aggregation
# of data_gen/tts/emotion/inference.py:58-155.  PINNED: tests/test_frontend_cpu.py checks these restatements against
# tests/golden/ref_emotion_encoder.npz, generated by tools/make_golden.py from the unmodified reference classes.
# (The 40-channel power mel that feeds it, data_gen/tts/emotion/audio.py:43-55, is librosa.feature.melspectrogram - third
# party, not restated here.)

EMO_HIDDEN = 256        # params_model.py: model_hidden_size
EMO_EMBED = 256         # params_model.py: model_embedding_size
EMO_LAYERS = 3          # params_model.py: model_num_layers
EMO_MELS = 40           # params_data.py: mel_n_channels
EMO_PARTIAL_FRAMES = 160  # params_data.py: partials_n_frames
EMO_SR = 16000          # params_data.py: sampling_rate
EMO_STEP_MS = 10        # params_data.py: mel_window_step


def emotion_encoder_weights(seed, hidden=EMO_HIDDEN, n_in=EMO_MELS, layers=EMO_LAYERS, embed=EMO_EMBED):
    """Seeded synthetic state_dict of the encoder (keys and shapes of model.py:16-21, torch.nn.LSTM layout: gate rows
    i | f | g | o).  numpy's legacy RandomState stream is stable across versions, so fixtures only carry inputs / outputs."""
    rs = np.random.RandomState(seed)
    k = 1.0 / np.sqrt(hidden)
    u = lambda *shape: rs.uniform(-k, k, size=shape).astype(np.float32)
    sd = {}
    for l in range(layers):
        sd["lstm.weight_ih_l%d" % l] = u(4 * hidden, n_in if l == 0 else hidden)
        sd["lstm.weight_hh_l%d" % l] = u(4 * hidden, hidden)
        sd["lstm.bias_ih_l%d" % l] = u(4 * hidden)
        sd["lstm.bias_hh_l%d" % l] = u(4 * hidden)
    sd["linear.weight"] = u(embed, hidden)
    sd["linear.bias"] = u(embed)
    return sd


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def lstm_hidden(frames, sd, layers=EMO_LAYERS, dtype=np.float64):
    """EmotionEncoder.inference (model.py:62-77): batch-first multi-layer LSTM from zero state, returns hidden[-1]
    (the last layer's h at the last frame), [P, hidden].  Gate arithmetic of torch.nn.LSTM: i, f, o = sigmoid, g = tanh,
    c = f c + i g, h = o tanh(c)."""
    x = np.asarray(frames, dtype)
    P, T, _ = x.shape
    for l in range(layers):
        wih = sd["lstm.weight_ih_l%d" % l].astype(dtype)
        whh = sd["lstm.weight_hh_l%d" % l].astype(dtype)
        b = (sd["lstm.bias_ih_l%d" % l].astype(dtype) + sd["lstm.bias_hh_l%d" % l].astype(dtype))
        H = whh.shape[1]
        h = np.zeros((P, H), dtype)
        c = np.zeros((P, H), dtype)
        out = np.empty((P, T, H), dtype)
        xp = x @ wih.T + b
        for t in range(T):
            g = xp[:, t] + h @ whh.T
            i, f, gg, o = g[:, :H], g[:, H:2 * H], g[:, 2 * H:3 * H], g[:, 3 * H:]
            c = _sigmoid(f) * c + _sigmoid(i) * np.tanh(gg)
            h = _sigmoid(o) * np.tanh(c)
            out[:, t] = h
        x = out
    return x[:, -1].astype(np.float32)


def emotion_embeds(hidden, sd):
    """EmotionEncoder.forward after the LSTM (model.py:51-57): relu(linear(hidden[-1])), L2-normalised per row."""
    h = np.asarray(hidden, np.float64)
    e = np.maximum(0.0, h @ sd["linear.weight"].astype(np.float64).T + sd["linear.bias"].astype(np.float64))
    return (e / np.linalg.norm(e, axis=1, keepdims=True)).astype(np.float32)


def utterance_embed(partial_hidden):
    """embed_utterance (inference.py:150-151): mean of the partial embeddings, L2-normalised."""
    raw = np.mean(np.asarray(partial_hidden, np.float32), axis=0)
    return raw / np.linalg.norm(raw, 2)


def emotion_mel(wav):
    """wav_to_mel_spectrogram (data_gen/tts/emotion/audio.py:43-55): librosa.feature.melspectrogram(y, sr=16000, n_fft=400,
    hop_length=160, n_mels=40) with librosa 0.8.0's defaults - win_length = n_fft, hann, center=True, pad_mode='reflect',
    power=2.0, filters.mel(fmin=0, fmax=sr/2, htk=False, norm='slaney') - transposed to [T, 40] float32.  Third-party
    arithmetic restated from the published algorithm: PARITY UNPINNED against librosa (STFT checked against scipy)."""
    n_fft, hop = int(EMO_SR * 25 / 1000), int(EMO_SR * EMO_STEP_MS / 1000)
    S = np.abs(stft(wav, n_fft, hop, n_fft, pad_mode="reflect")) ** 2.0
    return (mel_basis(EMO_SR, n_fft, EMO_MELS, 0.0, EMO_SR / 2) @ S).astype(np.float32).T


def compute_partial_slices(n_samples, partial_utterance_n_frames=EMO_PARTIAL_FRAMES, min_pad_coverage=0.75, overlap=0.5):
    """inference.py:58-107 as (start, stop) pairs: (wav ranges, mel ranges)."""
    assert 0 <= overlap < 1 and 0 < min_pad_coverage <= 1
    spf = int(EMO_SR * EMO_STEP_MS / 1000)
    n_frames = int(np.ceil((n_samples + 1) / spf))
    frame_step = max(int(np.round(partial_utterance_n_frames * (1 - overlap))), 1)
    steps = max(1, n_frames - partial_utterance_n_frames + frame_step + 1)
    mel = [(i, i + partial_utterance_n_frames) for i in range(0, steps, frame_step)]
    wav = [(a * spf, b * spf) for a, b in mel]
    coverage = (n_samples - wav[-1][0]) / (wav[-1][1] - wav[-1][0])
    if coverage < min_pad_coverage and len(mel) > 1:
        mel, wav = mel[:-1], wav[:-1]
    return wav, mel


def resemblyzer_partial_slices(n_samples, rate=1.3, min_coverage=0.75):
    """VoiceEncoder.compute_partial_slices of resemblyzer 0.1.1.dev0 (third party, requirements.txt:14; the code the reference's
    data_gen/tts/emotion/inference.py:58-107 was forked from, with `rate` partials per second instead of an overlap fraction),
    as (start, stop) pairs: (wav ranges, mel ranges).  Restated from the published source; PARITY UNPINNED."""
    assert 0 < min_coverage <= 1
    spf = int(EMO_SR * EMO_STEP_MS / 1000)
    n_frames = int(np.ceil((n_samples + 1) / spf))
    frame_step = int(np.round((EMO_SR / rate) / spf))
    assert 0 < frame_step <= EMO_PARTIAL_FRAMES
    mel, wav = [], []
    steps = max(1, n_frames - EMO_PARTIAL_FRAMES + frame_step + 1)
    for i in range(0, steps, frame_step):
        mel.append((i, i + EMO_PARTIAL_FRAMES))
        wav.append((i * spf, (i + EMO_PARTIAL_FRAMES) * spf))
    coverage = (n_samples - wav[-1][0]) / (wav[-1][1] - wav[-1][0])
    if coverage < min_coverage and len(mel) > 1:
        mel, wav = mel[:-1], wav[:-1]
    return wav, mel
