"""Host-side mirror of the reference's emotion-encoder module (data_gen/tts/emotion/inference.py) on the CUDA front-end.

Same names and argument meaning as the reference for the part between a PREPROCESSED waveform (16 kHz, volume-normalised,
VAD-trimmed by data_gen/tts/emotion/audio.py:14-40 - librosa / webrtcvad, third party, outside this package) and the
`emo_embed` vector of inference/StyleSinger.py:106:

    load_model(weights)            inference.py:15-37   (a state_dict / checkpoint dict instead of a path is accepted too)
    is_loaded()                    inference.py:39-40
    wav_to_mel_spectrogram(wav)    audio.py:43-55       (ssb_melspec_create_ex: reflect centring, power 2, no log)
    embed_frames_batch(frames)     inference.py:43-55   (ssb_lstm_encoder_forward -> hidden[-1])
    compute_partial_slices(n)      inference.py:58-107  (host integers, restated)
    embed_utterance(wav, ...)      inference.py:110-155

Everything numeric runs in libstylesinger_b200.so; there is no CPU fallback.
"""
import numpy as np
import torch

# data_gen/tts/emotion/params_data.py
mel_window_length = 25
mel_window_step = 10
mel_n_channels = 40
sampling_rate = 16000
partials_n_frames = 160

_model = None
_mel = None
_device = None


def load_model(weights, device=None):
    """weights: path of a reference checkpoint (torch.load -> ['model_state'], inference.py:29-30), that dict, or a bare state_dict."""
    global _model, _mel, _device
    from .engine import LstmEncoder, MelSpectrogram
    if isinstance(weights, (str, bytes)) or hasattr(weights, "__fspath__"):
        weights = torch.load(weights, map_location="cpu")
    if "model_state" in weights:
        weights = weights["model_state"]
    _device = torch.device(device if device is not None else "cuda:0")
    _model = LstmEncoder(weights, _device)
    _mel = MelSpectrogram(dict(audio_sample_rate=sampling_rate, fft_size=int(sampling_rate * mel_window_length / 1000),
                               hop_size=int(sampling_rate * mel_window_step / 1000), win_size=int(sampling_rate * mel_window_length / 1000),
                               audio_num_mel_bins=mel_n_channels, fmin=0, fmax=sampling_rate / 2),
                          _device, pad_reflect=True, power=True, log=False)


def is_loaded():
    return _model is not None


def _need_model():
    if _model is None:
        raise Exception("Model was not loaded. Call load_model() before inference.")  # inference.py:51-52


def wav_to_mel_spectrogram(wav, device_out=False):
    """[frames, 40] float32 power mel of a preprocessed waveform (audio.py:43-55)."""
    _need_model()
    mel = _mel(np.asarray(wav, np.float32))
    return mel if device_out else mel.cpu().numpy()


def embed_frames_batch(frames_batch):
    """(batch, n_frames, 40) float32 -> (batch, 256) float32 numpy (inference.py:43-55)."""
    _need_model()
    return _model(frames_batch)["hidden"].cpu().numpy()


def compute_partial_slices(n_samples, partial_utterance_n_frames=partials_n_frames, min_pad_coverage=0.75, overlap=0.5):
    """inference.py:58-107: (wav_slices, mel_slices) as lists of python slices.  Partials start every
    round(frames * (1 - overlap)) mel frames; the last one is dropped when less than ``min_pad_coverage`` of it lies inside
    the waveform (unless it is the only one)."""
    if not (0 <= overlap < 1 and 0 < min_pad_coverage <= 1):
        raise AssertionError("overlap in [0, 1), min_pad_coverage in (0, 1]")
    hop = sampling_rate * mel_window_step // 1000                       # samples per mel frame
    total_frames = -(-(n_samples + 1) // hop)                           # ceil
    stride = max(int(np.round(partial_utterance_n_frames * (1 - overlap))), 1)
    last_start_bound = max(1, total_frames - partial_utterance_n_frames + stride + 1)
    starts = list(range(0, last_start_bound, stride))
    if len(starts) > 1:
        inside = n_samples - starts[-1] * hop                           # samples of the last partial that exist
        if inside / (partial_utterance_n_frames * hop) < min_pad_coverage:
            starts.pop()
    mel_slices = [slice(f, f + partial_utterance_n_frames) for f in starts]
    wav_slices = [slice(f * hop, (f + partial_utterance_n_frames) * hop) for f in starts]
    return wav_slices, mel_slices


def embed_utterance(wav, using_partials=True, return_partials=False, **kwargs):
    """inference.py:110-155.  The waveform goes to the GPU once; mel frames, partial batch, LSTM and the normalised mean stay
    on the device, only the 256 (+ partials) floats come back."""
    _need_model()
    samples = np.asarray(wav, np.float32).reshape(-1)
    if using_partials:
        wave_slices, mel_slices = compute_partial_slices(len(samples), **kwargs)
        short = wave_slices[-1].stop - len(samples)
        if short >= 0:                                   # :141-142 (zero-pad up to the end of the last partial)
            samples = np.concatenate([samples, np.zeros(short, np.float32)])
        frames = wav_to_mel_spectrogram(samples, device_out=True)
        batch = torch.stack([frames[s] for s in mel_slices])
        res = _model(batch, utt_offsets=[0, len(mel_slices)])   # mean of the partial embeddings, L2-normalised (:150-151)
        embed = res["utt_embed"][0].cpu().numpy()
        return (embed, res["hidden"].cpu().numpy(), wave_slices) if return_partials else embed
    frames = wav_to_mel_spectrogram(samples, device_out=True)    # :129-134: the whole spectrogram as one sequence
    embed = _model(frames[None])["hidden"][0].cpu().numpy()
    return (embed, None, None) if return_partials else embed
