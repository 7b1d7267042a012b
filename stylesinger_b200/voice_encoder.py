"""Host-side mirror of resemblyzer's ``VoiceEncoder`` (third party, resemblyzer==0.1.1.dev0 in the reference's requirements.txt:14)
for the call of inference/StyleSinger.py:100-104 (``spk_embed = VoiceEncoder().cuda().embed_utterance(wav)``) on the CUDA
front-end of this package.

The network is the architecture the reference's own EmotionEncoder was forked from - a 3-layer LSTM 40 -> 256 over 160-frame
partials, ``relu(linear(hidden[-1]))`` L2-normalised per partial, the utterance embedding = normalised mean of the partial
embeddings - so it runs on the same kernels (``ssb_lstm_encoder_forward`` with ``embeds_out``, pinned against the reference's
EmotionEncoder.forward) and the same 40-band power mel (``ssb_melspec_create_ex``).  What differs from
``stylesinger_b200.emotion`` is the slicing (partials start every ``round(16000 / rate / 160)`` frames, rate = 1.3) and that
the mean is taken over ``forward`` outputs.  resemblyzer is not installed in the build image and its pretrained weights are not
part of the reference: this module restates the published algorithm and is PARITY UNPINNED against the package itself; the
weights file (resemblyzer's ``pretrained.pt``, ``['model_state']``) must be supplied by the caller.
"""
import numpy as np
import torch

sampling_rate = 16000
mel_window_length = 25
mel_window_step = 10
mel_n_channels = 40
partials_n_frames = 160


class VoiceEncoder:
    def __init__(self, weights, device=None):
        """weights: path of resemblyzer's checkpoint, the loaded dict (``['model_state']``) or a bare state_dict."""
        from .engine import LstmEncoder, MelSpectrogram
        if isinstance(weights, (str, bytes)) or hasattr(weights, "__fspath__"):
            weights = torch.load(weights, map_location="cpu")
        if "model_state" in weights:
            weights = weights["model_state"]
        self.device = torch.device(device if device is not None else "cuda:0")
        self._net = LstmEncoder(weights, self.device)
        if self._net.embed == 0:
            raise KeyError("VoiceEncoder weights need linear.weight / linear.bias")
        n_fft = int(sampling_rate * mel_window_length / 1000)
        self._mel = MelSpectrogram(dict(audio_sample_rate=sampling_rate, fft_size=n_fft, hop_size=int(sampling_rate * mel_window_step / 1000),
                                        win_size=n_fft, audio_num_mel_bins=mel_n_channels, fmin=0, fmax=sampling_rate / 2),
                                   self.device, pad_reflect=True, power=True, log=False)

    def cuda(self):  # the reference writes VoiceEncoder().cuda()
        return self

    def forward(self, mels):
        """(batch, n_frames, 40) -> (batch, 256) L2-normalised partial embeddings (device tensor)."""
        return self._net(mels, want_embeds=True)["embeds"]

    __call__ = forward

    @staticmethod
    def compute_partial_slices(n_samples, rate, min_coverage):
        """Where to cut an utterance of ``n_samples`` into 160-frame partials, ``rate`` partials per second; the last partial is
        dropped when less than ``min_coverage`` of it lies inside the waveform (unless it is the only one)."""
        if not 0 < min_coverage <= 1:
            raise AssertionError("min_coverage in (0, 1]")
        hop = sampling_rate * mel_window_step // 1000
        total_frames = -(-(n_samples + 1) // hop)
        stride = int(np.round((sampling_rate / rate) / hop))
        if not 0 < stride:
            raise AssertionError("The rate is too high")
        if stride > partials_n_frames:
            raise AssertionError("The rate is too low, it should be %f at least" % (sampling_rate / (hop * partials_n_frames)))
        starts = list(range(0, max(1, total_frames - partials_n_frames + stride + 1), stride))
        if len(starts) > 1 and (n_samples - starts[-1] * hop) / (partials_n_frames * hop) < min_coverage:
            starts.pop()
        mel_slices = [slice(f, f + partials_n_frames) for f in starts]
        wav_slices = [slice(f * hop, (f + partials_n_frames) * hop) for f in starts]
        return wav_slices, mel_slices

    def embed_utterance(self, wav, return_partials=False, rate=1.3, min_coverage=0.75):
        """(256,) float32 utterance embedding of a waveform (16 kHz as resemblyzer defines it; the reference hands over its
        float16 48 kHz array unchanged and so does this mirror)."""
        samples = np.asarray(wav, np.float32).reshape(-1)
        wav_slices, mel_slices = self.compute_partial_slices(len(samples), rate, min_coverage)
        short = wav_slices[-1].stop - len(samples)
        if short >= 0:
            samples = np.concatenate([samples, np.zeros(short, np.float32)])
        mel = self._mel(samples)
        partial_embeds = self.forward(torch.stack([mel[s] for s in mel_slices])).cpu().numpy()
        raw_embed = np.mean(partial_embeds, axis=0)
        embed = raw_embed / np.linalg.norm(raw_embed, 2)
        if return_partials:
            return embed, partial_embeds, wav_slices
        return embed
