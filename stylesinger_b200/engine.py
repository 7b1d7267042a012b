"""Host-side engine: checkpoint packing, ragged batches, workspaces and the calls into the C ABI.

PyTorch is used here only for device memory, streams and host<->device copies; every arithmetic
operation of the path runs in libstylesinger_b200.so (see include/stylesinger_b200.h).
"""
import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib
from ._lib import AcousticInputs, AcousticOutputs, HParams, TensorDesc, VocoderConfig, check, lib
from .hparams import DEFAULT_VOCODER_CONFIG, resolve
from .schedules import multinomial_table, sampler_table


def _require_cuda():
    if not torch.cuda.is_available():
        raise _lib.SsbError("stylesinger_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def _ptr(t: Optional[torch.Tensor]):
    """Device pointer of a tensor handed to the C ABI. The library reads plain row-major memory, so a strided
    (e.g. transposed / Fortran-ordered) view would be silently misread: refuse it loudly."""
    if t is None:
        return None
    if not t.is_contiguous():
        raise ValueError("stylesinger_b200: tensors passed to the C ABI must be contiguous (call .contiguous())")
    return C.c_void_p(t.data_ptr())


def _descs(named: Dict[str, torch.Tensor]):
    """ssb_tensor_desc array over fp32 contiguous CPU copies (kept alive by the returned list)."""
    keep, arr = [], (TensorDesc * len(named))()
    for i, (k, v) in enumerate(named.items()):
        t = v.detach().to("cpu", torch.float32).contiguous()
        if t.dim() == 0:
            t = t.reshape(1)
        if t.dim() > 4:
            raise ValueError(f"{k}: rank > 4")
        keep.append(t)
        kb = k.encode()
        keep.append(kb)
        arr[i].name = kb
        arr[i].data = t.data_ptr()
        arr[i].ndim = t.dim()
        for d in range(t.dim()):
            arr[i].shape[d] = t.shape[d]
    return arr, keep


def sinusoid_table(n, dim=256, padding_idx=0):
    """SinusoidalPositionalEmbedding.get_embedding (reference modules/commons/common_layers.py:111-127),
    computed on the host exactly like the reference does at module construction."""
    half = dim // 2
    e = math.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half, dtype=torch.float) * -e)
    e = torch.arange(n, dtype=torch.float).unsqueeze(1) * e.unsqueeze(0)
    e = torch.cat([torch.sin(e), torch.cos(e)], dim=1).view(n, -1)
    e[padding_idx, :] = 0
    return e


def step_embedding(T, C_):
    """SinusoidalPosEmb(t) for t = 0..T-1 (reference modules/diff/net.py:31-44)."""
    half = C_ // 2
    e = math.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half) * -e)
    e = torch.arange(T)[:, None] * e[None, :]
    return torch.cat((e.sin(), e.cos()), dim=-1).float().contiguous()


@dataclass
class PackedBatch:
    """A ragged batch in the library's tight layout. Offsets are host int32 arrays [B+1]."""
    B: int
    ph_offsets: np.ndarray
    ref_offsets: np.ndarray
    frame_offsets: Optional[np.ndarray]
    t: Dict[str, torch.Tensor] = field(default_factory=dict)  # txt_tokens, note, note_type (int32), note_dur,
    # spk_embed, emo_embed, ref_mels, ref_f0, [mel2ph int32], [f0], [uv]
    may_have_pad_frames: bool = True  # False when the host knows every frame maps to a phone (mel2ph > 0 everywhere)

    def to(self, device, non_blocking=True):
        return PackedBatch(self.B, self.ph_offsets, self.ref_offsets, self.frame_offsets,
                           {k: v.to(device, non_blocking=non_blocking) for k, v in self.t.items()},
                           self.may_have_pad_frames)

    def h2d_bytes(self):
        return int(sum(v.numel() * v.element_size() for v in self.t.values()))

    @property
    def total_frames(self):
        return int(self.frame_offsets[-1]) if self.frame_offsets is not None else 0


def pack_batch(utts: List[dict], use_mel2ph=True, pin=False) -> PackedBatch:
    """Concatenate per-utterance CPU tensors (as produced by synth.make_utterance) into one PackedBatch."""
    def offs(lens):
        return np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)

    B = len(utts)
    po = offs([len(u["txt_tokens"]) for u in utts])
    ro = offs([u["ref_mels"].shape[0] for u in utts])
    fo = offs([len(u["mel2ph"]) for u in utts]) if use_mel2ph else None
    cat = lambda k, dt: torch.cat([u[k].reshape(-1) if u[k].dim() == 1 else u[k] for u in utts]).to(dt).contiguous()
    t = {"txt_tokens": cat("txt_tokens", torch.int32), "note": cat("note", torch.int32),
         "note_type": cat("note_type", torch.int32), "note_dur": cat("note_dur", torch.float32),
         "spk_embed": torch.stack([u["spk_embed"] for u in utts]).float().contiguous(),
         "emo_embed": torch.stack([u["emo_embed"] for u in utts]).float().contiguous(),
         "ref_mels": cat("ref_mels", torch.float32), "ref_f0": cat("ref_f0", torch.float32)}
    pad = False  # predicted durations: the length regulator never emits a zero entry inside an utterance
    if use_mel2ph:
        t["mel2ph"] = cat("mel2ph", torch.int32)
        pad = bool((t["mel2ph"] <= 0).any())
    if pin and torch.cuda.is_available():
        t = {k: v.pin_memory() for k, v in t.items()}
    return PackedBatch(B, po, ro, fo, t, pad)


class _Workspace:
    def __init__(self, device):
        self.device = device
        self.buf = None

    def get(self, nbytes):
        if self.buf is None or self.buf.numel() < nbytes:
            self.buf = None
            self.buf = torch.empty(int(nbytes * 1.05) + 4096, dtype=torch.uint8, device=self.device)
        return self.buf


class AcousticModel:
    """Packed StyleSinger acoustic model on one GPU (ssb_model_t)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], hparams=None, device=None, max_positions=4096):
        _require_cuda()
        self.hp = resolve(hparams)
        self.device = torch.device(device if device is not None else "cuda:0")
        torch.cuda.set_device(self.device)
        sd = {k: v for k, v in state_dict.items() if isinstance(v, torch.Tensor)}
        sd = dict(sd)
        sd["__pos_table"] = sinusoid_table(max_positions, self.hp["hidden_size"])
        if "encoder.embed_tokens.weight" not in sd and "encoder_embed_tokens.weight" in sd:
            sd["encoder.embed_tokens.weight"] = sd["encoder_embed_tokens.weight"]
        hp = self.hp
        h = HParams(hp["hidden_size"], hp["enc_layers"], hp["dec_layers"], hp["enc_ffn_kernel_size"],
                    hp["dec_ffn_kernel_size"], hp["dur_predictor_layers"], hp["dur_predictor_kernel"],
                    int(sd["encoder.embed_tokens.weight"].shape[0]), hp["nRQ"], hp["rq_depth"],
                    hp["residual_channels"], hp["residual_layers"], hp["dilation_cycle_length"],
                    hp["f0_residual_channels"], hp["f0_residual_layers"], hp["f0_dilation_cycle_length"],
                    hp["audio_num_mel_bins"])
        arr, keep = _descs(sd)
        handle = C.c_void_p()
        check(lib.ssb_model_create(C.byref(handle), arr, len(sd), C.byref(h)), "ssb_model_create")
        self._h = handle
        self._ws = _Workspace(self.device)
        self.T = self.f0_T = None
        self.set_timesteps(hp["timesteps"], hp["f0_timesteps"])

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib.ssb_model_free(h)
            self._h = None

    def set_tensor_cores(self, enable=True):
        """tcgen05 (fp16 hi/lo split, 3 MMAs) vs fp32 FFMA for the denoiser layer GEMMs. Returns the mode in effect."""
        return bool(lib.ssb_model_set_tensor_cores(self._h, 1 if enable else 0))

    def set_persistent(self, enable=True):
        """Single-launch persistent sampler kernel for small batches (True) vs one launch per GEMM (False)."""
        return bool(lib.ssb_model_set_persistent(self._h, 1 if enable else 0))

    def set_cond_hoist(self, enable=True):
        """Conditioner projection hoisted out of the T loop (default) vs contracted inside every layer GEMM."""
        return bool(lib.ssb_model_set_cond_hoist(self._h, 1 if enable else 0))

    def set_persistent_groups(self, enable=True):
        """Large batches: run the mel sampler as one persistent launch per group of <= 48 row tiles (default off)."""
        return bool(lib.ssb_model_set_persistent_groups(self._h, 1 if enable else 0))

    def set_fft_tensor_cores(self, enable: bool) -> bool:
        """Decoder FFT-block FFN GEMMs on the tcgen05 kernel for batches of >= 1024 frames (default on)."""
        return bool(lib.ssb_model_set_fft_tensor_cores(self._h, 1 if enable else 0))

    # -- schedules -------------------------------------------------------------------------------
    def set_timesteps(self, T=None, f0_T=None):
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        if T is not None and T != self.T:
            emb = step_embedding(T, self.hp["residual_channels"])
            g = np.ascontiguousarray(sampler_table(T, self.hp["max_beta"]))
            check(lib.ssb_model_set_schedule(self._h, 0, T, C.c_void_p(emb.data_ptr()), g.ctypes.data_as(C.c_void_p),
                                             None, stream), "ssb_model_set_schedule(mel)")
            self.T = T
        if f0_T is not None and f0_T != self.f0_T:
            emb = step_embedding(f0_T, self.hp["f0_residual_channels"])
            g = np.ascontiguousarray(sampler_table(f0_T, self.hp["f0_max_beta"]))
            m = np.ascontiguousarray(multinomial_table(f0_T, self.hp["f0_max_beta"]))
            check(lib.ssb_model_set_schedule(self._h, 1, f0_T, C.c_void_p(emb.data_ptr()),
                                             g.ctypes.data_as(C.c_void_p), m.ctypes.data_as(C.c_void_p), stream),
                  "ssb_model_set_schedule(f0)")
            self.f0_T = f0_T

    # -- helpers ---------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _inputs(self, pb: PackedBatch, noise=None, seed=0, skip_mel=False, dur=None, f0=None, uv=None):
        a = AcousticInputs()
        a.B = pb.B
        self._keep = [np.ascontiguousarray(pb.ph_offsets, np.int32), np.ascontiguousarray(pb.ref_offsets, np.int32)]
        a.ph_offsets = self._keep[0].ctypes.data
        a.ref_offsets = self._keep[1].ctypes.data
        if pb.frame_offsets is not None:
            fo = np.ascontiguousarray(pb.frame_offsets, np.int32)
            self._keep.append(fo)
            a.frame_offsets = fo.ctypes.data
        t = pb.t
        for k in ("txt_tokens", "note", "note_type", "note_dur", "spk_embed", "emo_embed", "ref_mels", "ref_f0"):
            assert t[k].is_cuda and t[k].is_contiguous(), k
            setattr(a, k, t[k].data_ptr())
        if "mel2ph" in t and dur is None:
            a.mel2ph = t["mel2ph"].data_ptr()
        if dur is not None:
            a.dur = dur.data_ptr()
        f0 = f0 if f0 is not None else t.get("f0")
        uv = uv if uv is not None else t.get("uv")
        if f0 is not None:
            a.f0 = f0.data_ptr()
            if uv is not None:
                a.uv = uv.data_ptr()
        if noise:
            for i in range(2):
                if noise.get("f0_gauss") is not None:
                    a.f0_gauss_noise[i] = noise["f0_gauss"][i].data_ptr()
                    a.f0_unif_noise[i] = noise["f0_unif"][i].data_ptr()
            if noise.get("mel") is not None:
                a.mel_noise = noise["mel"].data_ptr()
        a.seed = int(seed)
        a.skip_mel_diffusion = 1 if skip_mel else 0
        a.pndm_speedup = int(self.hp.get("pndm_speedup") or 0)  # reference hparam: 0 / absent = DDPM (the StyleSinger default)
        return a

    # -- entry points ------------------------------------------------------------------------------
    def predict_durations(self, pb: PackedBatch):
        """DurationPredictor.inference; returns (dur int32 [sumP], log-dur f32 [sumP]) on the device."""
        a = self._inputs(pb)
        n = lib.ssb_durations_workspace_bytes(self._h, C.byref(a))
        if n == 0:
            check(-1, "ssb_durations_workspace_bytes")
        ws = self._ws.get(n)
        P = int(pb.ph_offsets[-1])
        dur = torch.empty(P, dtype=torch.int32, device=self.device)
        logdur = torch.empty(P, dtype=torch.float32, device=self.device)
        check(lib.ssb_predict_durations(self._h, C.byref(a), _ptr(dur), _ptr(logdur), _ptr(ws), ws.numel(), self._stream()),
              "ssb_predict_durations")
        return dur, logdur

    def forward(self, pb: PackedBatch, noise=None, seed=0, skip_mel_diffusion=False, dur=None,
                want=("mel_out", "f0_denorm")):
        """StyleSinger.forward(infer=True). `pb` must carry frame_offsets (+ mel2ph, or pass `dur`).
        Returns a dict of tight device tensors for the keys in `want`."""
        assert pb.frame_offsets is not None, "frame_offsets required (run predict_durations first)"
        a = self._inputs(pb, noise, seed, skip_mel_diffusion, dur)
        Fs, Ps, Rs = int(pb.frame_offsets[-1]), int(pb.ph_offsets[-1]), int(pb.ref_offsets[-1])
        shapes = {"mel_out": (Fs, 80), "f0_denorm": (Fs,), "encoder_out": (Ps, 256), "style": (Fs, 256),
                  "rq_codes": (Rs, self.hp["rq_depth"]), "pitch_pred": (Fs, 2), "decoder_inp": (Fs, 256),
                  "coarse_mel": (Fs, 80), "diff_cond": (Fs, 256), "mel2ph": (Fs,), "spk_proj": (pb.B, 256),
                  "emo_proj": (pb.B, 256)}
        o = AcousticOutputs()
        out = {}
        want = set(want)
        if skip_mel_diffusion:
            want.discard("mel_out")  # never written in that mode: do not hand back an uninitialised buffer
        else:
            want.add("mel_out")
        for k in want:
            dt = torch.int32 if k in ("rq_codes", "mel2ph") else torch.float32
            out[k] = torch.empty(shapes[k], dtype=dt, device=self.device)
            setattr(o, k, out[k].data_ptr())
        n = lib.ssb_acoustic_workspace_bytes(self._h, C.byref(a))
        if n == 0:
            check(-1, "ssb_acoustic_workspace_bytes")
        ws = self._ws.get(n)
        check(lib.ssb_acoustic_forward(self._h, C.byref(a), C.byref(o), _ptr(ws), ws.numel(), self._stream()),
              "ssb_acoustic_forward")
        return out

    def mel_diffusion(self, cond, coarse, frame_offsets, noise=None, seed=0):
        fo = np.ascontiguousarray(frame_offsets, np.int32)
        B = len(fo) - 1
        n = lib.ssb_mel_diffusion_workspace_bytes(self._h, fo.ctypes.data, B)
        if n == 0:
            check(-1, "ssb_mel_diffusion_workspace_bytes")
        ws = self._ws.get(n)
        mel = torch.empty((int(fo[-1]), 80), dtype=torch.float32, device=self.device)
        check(lib.ssb_mel_diffusion_sample(self._h, _ptr(cond), _ptr(coarse), fo.ctypes.data, B, _ptr(noise), int(seed),
                                           _ptr(mel), _ptr(ws), ws.numel(), self._stream()), "ssb_mel_diffusion_sample")
        return mel

    def mel_diffusion_plms(self, cond, coarse, frame_offsets, interval, q_noise=None, seed=0):
        """PLMS sampler (hparams['pndm_speedup'] = interval) over the mel denoiser: T / interval (+1) evaluations."""
        fo = np.ascontiguousarray(frame_offsets, np.int32)
        B = len(fo) - 1
        n = lib.ssb_mel_diffusion_plms_workspace_bytes(self._h, fo.ctypes.data, B)
        if n == 0:
            check(-1, "ssb_mel_diffusion_plms_workspace_bytes")
        ws = self._ws.get(n)
        mel = torch.empty((int(fo[-1]), 80), dtype=torch.float32, device=self.device)
        check(lib.ssb_mel_diffusion_sample_plms(self._h, _ptr(cond), _ptr(coarse), fo.ctypes.data, B, _ptr(q_noise), int(seed),
                                                int(interval), _ptr(mel), _ptr(ws), ws.numel(), self._stream()),
              "ssb_mel_diffusion_sample_plms")
        return mel

    def denoiser_eval(self, which, x, uv, t, cond, frame_offsets):
        fo = np.ascontiguousarray(frame_offsets, np.int32)
        B, Fs = len(fo) - 1, int(fo[-1])
        C_ = self.hp["residual_channels"] if which == 0 else self.hp["f0_residual_channels"]
        L_ = self.hp["residual_layers"] if which == 0 else self.hp["f0_residual_layers"]
        rows = Fs + 16 * (B + 1) + 512
        ws = self._ws.get(rows * (12 * C_ + 2 * C_ * L_ + 1024) * 4 + (1 << 20))
        out = torch.empty((Fs, 80 if which == 0 else 3), dtype=torch.float32, device=self.device)
        check(lib.ssb_denoiser_eval(self._h, which, _ptr(x), _ptr(uv), int(t), _ptr(cond), fo.ctypes.data, B, _ptr(out),
                                    _ptr(ws), ws.numel(), self._stream()), "ssb_denoiser_eval")
        return out

    def f0_diffusion(self, which, cond, lo, hi, frame_offsets, gauss_noise=None, unif_noise=None, seed=0):
        fo = np.ascontiguousarray(frame_offsets, np.int32)
        B, Fs = len(fo) - 1, int(fo[-1])
        C_, L_ = self.hp["f0_residual_channels"], self.hp["f0_residual_layers"]
        rows = Fs + 16 * (B + 1) + 512
        ws = self._ws.get(rows * (12 * C_ + 2 * C_ * L_ + 1024) * 4 + (1 << 20))
        z = torch.empty(Fs, dtype=torch.float32, device=self.device)
        uv = torch.empty(Fs, dtype=torch.int32, device=self.device)
        check(lib.ssb_f0_diffusion_sample(self._h, which, _ptr(cond), _ptr(lo), _ptr(hi), fo.ctypes.data, B,
                                          _ptr(gauss_noise), _ptr(unif_noise), int(seed), _ptr(z), _ptr(uv), _ptr(ws),
                                          ws.numel(), self._stream()), "ssb_f0_diffusion_sample")
        return z, uv

    def fft_encoder(self, txt_tokens, ph_offsets):
        """FastspeechEncoder.forward: int32 tokens [sumP] (device) -> [sumP,256]."""
        po = np.ascontiguousarray(ph_offsets, np.int32)
        B = len(po) - 1
        n = lib.ssb_fft_workspace_bytes(self._h, 0, po.ctypes.data, B)
        if n == 0:
            check(-1, "ssb_fft_workspace_bytes")
        ws = self._ws.get(n)
        out = torch.empty((int(po[-1]), 256), dtype=torch.float32, device=self.device)
        check(lib.ssb_fft_encoder(self._h, _ptr(txt_tokens), po.ctypes.data, B, _ptr(out), _ptr(ws), ws.numel(), self._stream()),
              "ssb_fft_encoder")
        return out

    def fft_decoder(self, x, frame_offsets):
        """FastspeechDecoder.forward: x [sumF,256] (device) -> [sumF,256]."""
        fo = np.ascontiguousarray(frame_offsets, np.int32)
        B = len(fo) - 1
        n = lib.ssb_fft_workspace_bytes(self._h, 1, fo.ctypes.data, B)
        if n == 0:
            check(-1, "ssb_fft_workspace_bytes")
        ws = self._ws.get(n)
        out = torch.empty((int(fo[-1]), 256), dtype=torch.float32, device=self.device)
        check(lib.ssb_fft_decoder(self._h, _ptr(x), fo.ctypes.data, B, _ptr(out), _ptr(ws), ws.numel(), self._stream()),
              "ssb_fft_decoder")
        return out

    def get_style(self, decoder_inp, frame_offsets, ref_mels, ref_f0, ref_offsets):
        """StyleSinger.get_style: (decoder_inp [sumF,256], ref_mels [sumR,80], ref_f0 [sumR]) -> (style [sumF,256], codes)."""
        fo = np.ascontiguousarray(frame_offsets, np.int32)
        ro = np.ascontiguousarray(ref_offsets, np.int32)
        B = len(fo) - 1
        n = lib.ssb_get_style_workspace_bytes(self._h, fo.ctypes.data, ro.ctypes.data, B)
        if n == 0:
            check(-1, "ssb_get_style_workspace_bytes")
        ws = self._ws.get(n)
        style = torch.empty((int(fo[-1]), 256), dtype=torch.float32, device=self.device)
        codes = torch.empty((int(ro[-1]), self.hp["rq_depth"]), dtype=torch.int32, device=self.device)
        check(lib.ssb_get_style(self._h, _ptr(decoder_inp), fo.ctypes.data, _ptr(ref_mels), _ptr(ref_f0), ro.ctypes.data, B,
                                _ptr(style), _ptr(codes), _ptr(ws), ws.numel(), self._stream()), "ssb_get_style")
        return style, codes

    def rvq(self, x, ref_offsets):
        ro = np.ascontiguousarray(ref_offsets, np.int32)
        B, Rs = len(ro) - 1, int(ro[-1])
        D = self.hp["rq_depth"]
        ws = self._ws.get((Rs + 16 * (B + 1) + 512) * (512 + D + 8) * 4 + (1 << 20))
        q = torch.empty((Rs, 256), dtype=torch.float32, device=self.device)
        codes = torch.empty((Rs, D), dtype=torch.int32, device=self.device)
        check(lib.ssb_rvq_lookup(self._h, _ptr(x), ro.ctypes.data, B, _ptr(q), _ptr(codes), _ptr(ws), ws.numel(),
                                 self._stream()), "ssb_rvq_lookup")
        return q, codes


class Vocoder:
    """Packed HiFi-GAN(-NSF) generator on one GPU (ssb_vocoder_t)."""

    def __init__(self, state_dict, config=None, device=None):
        _require_cuda()
        self.cfg = dict(DEFAULT_VOCODER_CONFIG, **(config or {}))
        self.device = torch.device(device if device is not None else "cuda:0")
        torch.cuda.set_device(self.device)
        h = self.cfg
        if str(h.get("resblock", "1")) != "1":
            raise NotImplementedError("only ResBlock1 generators are supported")
        vc = VocoderConfig()
        vc.n_up = len(h["upsample_rates"])
        for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
            vc.up_rates[i], vc.up_kernels[i] = u, k
        vc.initial_channel = h["upsample_initial_channel"]
        vc.n_res = len(h["resblock_kernel_sizes"])
        for j, (k, d) in enumerate(zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"])):
            vc.res_kernels[j] = k
            for m in range(3):
                vc.res_dilations[j][m] = d[m]
        vc.use_pitch_embed = 1 if h.get("use_pitch_embed") else 0
        vc.sample_rate = h.get("audio_sample_rate", 48000)
        self.hop = int(np.prod(h["upsample_rates"]))
        sd = {k: v for k, v in state_dict.items() if isinstance(v, torch.Tensor)}
        arr, keep = _descs(sd)
        handle = C.c_void_p()
        check(lib.ssb_vocoder_create(C.byref(handle), arr, len(sd), C.byref(vc)), "ssb_vocoder_create")
        self._h = handle
        self._ws = _Workspace(self.device)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib.ssb_vocoder_free(h)
            self._h = None

    def set_tensor_cores(self, enable=True):
        return bool(lib.ssb_vocoder_set_tensor_cores(self._h, 1 if enable else 0))

    max_frames_per_call = 24000  # ~0.9 MB of stage buffers per frame: bounds the workspace to ~20 GB

    def generate(self, mel, f0, frame_offsets, rand_ini=None, src_noise=None, seed=0):
        """mel [sumF,80], f0 [sumF] or None (device, tight) -> wav [sumF*hop] (device).
        Large batches are processed in groups of utterances (results are per-utterance, so grouping is exact)."""
        fo = np.ascontiguousarray(frame_offsets, np.int32)
        B = len(fo) - 1
        if B > 1 and int(fo[-1]) > self.max_frames_per_call:
            wav = torch.empty(int(fo[-1]) * self.hop, dtype=torch.float32, device=self.device)
            b0 = 0
            while b0 < B:
                b1 = b0 + 1
                while b1 < B and int(fo[b1 + 1] - fo[b0]) <= self.max_frames_per_call:
                    b1 += 1
                sub_fo = (fo[b0:b1 + 1] - fo[b0]).astype(np.int32)
                a, e = int(fo[b0]), int(fo[b1])
                w = self.generate(mel[a:e], None if f0 is None else f0[a:e], sub_fo,
                                  None if rand_ini is None else rand_ini[b0:b1].contiguous(),
                                  None if src_noise is None else src_noise[a * self.hop:e * self.hop], seed + b0)
                wav[a * self.hop:e * self.hop] = w
                b0 = b1
            return wav
        n = lib.ssb_vocoder_workspace_bytes(self._h, fo.ctypes.data, B)
        if n == 0:
            check(-1, "ssb_vocoder_workspace_bytes")
        ws = self._ws.get(n)
        wav = torch.empty(int(fo[-1]) * self.hop, dtype=torch.float32, device=self.device)
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        check(lib.ssb_hifigan_generate(self._h, _ptr(mel), _ptr(f0), fo.ctypes.data, B, _ptr(rand_ini), _ptr(src_noise),
                                       int(seed), _ptr(wav), _ptr(ws), ws.numel(), stream), "ssb_hifigan_generate")
        return wav


# unit-test granularity ops -------------------------------------------------------------------------
class MelSpectrogram:
    """log10-mel front-end of the reference audio (ssb_melspec_t): librosa_wav2spec of the reference
    (utils/audios/__init__.py:36-84) with the hparams of egs/stylesinger.yaml by default."""

    def __init__(self, hp=None, device=None, eps=1e-6, pad_reflect=False, power=False, log=True):
        """pad_reflect / power / log=False select the defaults of librosa.feature.melspectrogram instead of those of
        librosa_wav2spec (ssb_melspec_create_ex; the emotion encoder's features, data_gen/tts/emotion/audio.py:43-55)."""
        _require_cuda()
        h = dict(audio_sample_rate=48000, fft_size=1024, hop_size=256, win_size=1024, audio_num_mel_bins=80, fmin=20, fmax=24000)
        h.update({k: v for k, v in (hp or {}).items() if k in h})
        self.hp = h
        self.device = torch.device(device if device is not None else "cuda:0")
        torch.cuda.set_device(self.device)
        handle = C.c_void_p()
        check(lib.ssb_melspec_create_ex(C.byref(handle), h["audio_sample_rate"], h["fft_size"], h["hop_size"], h["win_size"],
                                        h["audio_num_mel_bins"], float(h["fmin"]), float(h["fmax"]), float(eps),
                                        int(bool(pad_reflect)), int(bool(power)), int(bool(log))), "ssb_melspec_create_ex")
        self._h = handle
        self._ws = _Workspace(self.device)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib.ssb_melspec_free(h)
            self._h = None

    def num_frames(self, n_samples):
        return int(lib.ssb_melspec_num_frames(self._h, int(n_samples)))

    def __call__(self, wavs):
        """wavs: list of 1-D float32 arrays / tensors (or one).  Returns a list of device tensors [frames, n_mels]."""
        single = not isinstance(wavs, (list, tuple))
        wavs = [wavs] if single else list(wavs)
        ts = [torch.as_tensor(np.asarray(w, dtype=np.float32) if not isinstance(w, torch.Tensor) else w, dtype=torch.float32).reshape(-1)
              for w in wavs]
        offs = np.concatenate([[0], np.cumsum([t.numel() for t in ts])]).astype(np.int32)
        wav = torch.cat(ts).to(self.device).contiguous() if ts else torch.zeros(0, device=self.device)
        frames = [self.num_frames(t.numel()) for t in ts]
        out = torch.empty(sum(frames), self.hp["audio_num_mel_bins"], dtype=torch.float32, device=self.device)
        n = lib.ssb_melspec_workspace_bytes(self._h, offs.ctypes.data, len(ts))
        if n == 0:
            check(-1, "ssb_melspec_workspace_bytes")
        ws = self._ws.get(n)
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        check(lib.ssb_melspec_forward(self._h, _ptr(wav), offs.ctypes.data, len(ts), _ptr(out), _ptr(ws), ws.numel(), stream),
              "ssb_melspec_forward")
        mels = list(torch.split(out, frames))
        return mels[0] if single else mels


class LstmEncoder:
    """LSTM utterance encoder of the reference audio (ssb_lstm_encoder_t): the reference's EmotionEncoder
    (data_gen/tts/emotion/model.py:10-77).  ``state_dict``: the encoder's own keys (``lstm.weight_ih_l0`` ..., optional
    ``linear.weight`` / ``linear.bias``), torch tensors or numpy arrays."""

    def __init__(self, state_dict, device=None):
        _require_cuda()
        self.device = torch.device(device if device is not None else "cuda:0")
        torch.cuda.set_device(self.device)
        sd = {k: np.ascontiguousarray(v.detach().cpu().float().numpy() if isinstance(v, torch.Tensor) else np.asarray(v, np.float32))
              for k, v in state_dict.items() if k.startswith(("lstm.", "linear."))}
        L = 0
        while "lstm.weight_ih_l%d" % L in sd:
            L += 1
        if L == 0:
            raise KeyError("state_dict has no lstm.weight_ih_l0")
        self.layers, self.hidden = L, sd["lstm.weight_hh_l0"].shape[1]
        self.n_in = sd["lstm.weight_ih_l0"].shape[1]
        for l in range(L):
            for k, shape in (("weight_ih", (4 * self.hidden, self.n_in if l == 0 else self.hidden)), ("weight_hh", (4 * self.hidden, self.hidden)),
                             ("bias_ih", (4 * self.hidden,)), ("bias_hh", (4 * self.hidden,))):
                name = "lstm.%s_l%d" % (k, l)
                if name not in sd or sd[name].shape != shape:
                    raise ValueError(f"{name}: expected shape {shape}, got {sd[name].shape if name in sd else None}")
        table = lambda k: (C.c_void_p * L)(*[sd["lstm.%s_l%d" % (k, l)].ctypes.data for l in range(L)])
        lw, lb = sd.get("linear.weight"), sd.get("linear.bias")
        self.embed = 0 if lw is None else lw.shape[0]
        if lw is not None and (lb is None or lw.shape != (self.embed, self.hidden) or lb.shape != (self.embed,)):
            raise ValueError("linear.weight / linear.bias shapes do not match the LSTM")
        handle = C.c_void_p()
        check(lib.ssb_lstm_encoder_create(C.byref(handle), self.n_in, self.hidden, L, table("weight_ih"), table("weight_hh"),
                                          table("bias_ih"), table("bias_hh"), self.embed,
                                          None if lw is None else lw.ctypes.data, None if lw is None else lb.ctypes.data),
              "ssb_lstm_encoder_create")
        self._h = handle
        self._ws = _Workspace(self.device)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib.ssb_lstm_encoder_free(h)
            self._h = None

    def __call__(self, frames, utt_offsets=None, want_embeds=False):
        """frames: [P, T, n_in] (device or host).  Returns a dict of device tensors: ``hidden`` [P, H] (EmotionEncoder.inference),
        ``embeds`` [P, E] (EmotionEncoder.forward, if want_embeds) and ``utt_embed`` [U, H] (normalised mean of ``hidden`` over
        partials utt_offsets[u] .. utt_offsets[u+1], if utt_offsets is given)."""
        x = torch.as_tensor(frames, dtype=torch.float32).to(self.device).contiguous()
        if x.dim() != 3 or x.shape[2] != self.n_in:
            raise ValueError(f"frames must be [partials, frames, {self.n_in}]")
        Pn, T = int(x.shape[0]), int(x.shape[1])
        out = {"hidden": torch.empty(Pn, self.hidden, dtype=torch.float32, device=self.device)}
        if want_embeds:
            if self.embed == 0:
                raise _lib.SsbError("encoder was created without the linear head (linear.weight / linear.bias)")
            out["embeds"] = torch.empty(Pn, self.embed, dtype=torch.float32, device=self.device)
        offs, U = None, 0
        if utt_offsets is not None:
            offs = np.ascontiguousarray(utt_offsets, np.int32)
            U = len(offs) - 1
            out["utt_embed"] = torch.empty(U, self.hidden, dtype=torch.float32, device=self.device)
        n = lib.ssb_lstm_encoder_workspace_bytes(self._h, Pn, T, U)
        if n == 0:
            check(-1, "ssb_lstm_encoder_workspace_bytes")
        ws = self._ws.get(n)
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        check(lib.ssb_lstm_encoder_forward(self._h, _ptr(x), Pn, T, None if offs is None else offs.ctypes.data, U, _ptr(out["hidden"]),
                                           _ptr(out.get("embeds")), _ptr(out.get("utt_embed")), _ptr(ws), ws.numel(), stream),
              "ssb_lstm_encoder_forward")
        return out


def op_conv1d(x, offsets, w, b, dilation=1, act=0):
    _require_cuda()
    off = np.ascontiguousarray(offsets, np.int32)
    N, Cin, k = w.shape
    wc = w.detach().cpu().float().contiguous()
    bc = None if b is None else b.detach().cpu().float().contiguous()
    out = torch.empty((x.shape[0], N), dtype=torch.float32, device=x.device)
    stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
    check(lib.ssb_op_conv1d(_ptr(x), off.ctypes.data, len(off) - 1, Cin, C.c_void_p(wc.data_ptr()),
                            None if bc is None else C.c_void_p(bc.data_ptr()), N, k, dilation, act, _ptr(out), stream),
          "ssb_op_conv1d")
    return out


def op_conv1d_tc(x, offsets, w, b, dilation=1):
    _require_cuda()
    off = np.ascontiguousarray(offsets, np.int32)
    N, Cin, k = w.shape
    wc = w.detach().cpu().float().contiguous()
    bc = None if b is None else b.detach().cpu().float().contiguous()
    out = torch.empty((x.shape[0], N), dtype=torch.float32, device=x.device)
    stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
    check(lib.ssb_op_conv1d_tc(_ptr(x), off.ctypes.data, len(off) - 1, Cin, C.c_void_p(wc.data_ptr()),
                               None if bc is None else C.c_void_p(bc.data_ptr()), N, k, dilation, _ptr(out), stream),
          "ssb_op_conv1d_tc")
    return out


def op_attention(q, k, v, q_offsets, k_offsets, scale, tc=False):
    """tc=True: the tcgen05 / TMA kernel (ssb_op_attention_tc) instead of the fp32 one."""
    _require_cuda()
    qo = np.ascontiguousarray(q_offsets, np.int32)
    ko = np.ascontiguousarray(k_offsets, np.int32)
    out = torch.empty_like(q)
    stream = C.c_void_p(torch.cuda.current_stream(q.device).cuda_stream)
    fn = lib.ssb_op_attention_tc if tc else lib.ssb_op_attention
    check(fn(_ptr(q), _ptr(k), _ptr(v), qo.ctypes.data, ko.ctypes.data, len(qo) - 1, float(scale),
             _ptr(out), stream), "ssb_op_attention_tc" if tc else "ssb_op_attention")
    return out
