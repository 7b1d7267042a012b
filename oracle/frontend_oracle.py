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


def stft(y, n_fft, hop_length, win_length):
    """librosa.stft(y, n_fft, hop_length, win_length, window='hann', center=True, pad_mode='constant') -> complex64 [1 + n_fft/2, T].
    librosa multiplies the float64 window into the float32 frames (-> float64), runs the real FFT and stores complex64."""
    y = np.asarray(y, dtype=np.float32)
    w = pad_center(hann_periodic(win_length), n_fft).reshape(-1, 1)
    yp = np.pad(y, n_fft // 2, mode="constant")
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
