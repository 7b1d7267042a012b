"""This repo's drop-ins running INSIDE the unmodified reference (VERDICT r1 "missing" 5): the reference's own
`inference.StyleSinger.StyleSingerInfer` drives (1) its own sampler loops with the denoisers, the FFT encoder / decoder and
the style path replaced through the extension points SURVEY.md section 8b lists, (2) the whole acoustic model replaced
by `stylesinger_b200.modules.StyleSinger`, and (3) a vocoder registered under 'HifiGAN_NSF' in its vocoder registry.
Results are compared with the reference's own outputs for the same random draws (torch.randn* patched to one seeded
stream, as tools/make_golden.py does).  Needs the staged reference (baseline/_ref, written by build()) or /root/reference.
"""
import contextlib
import os
import sys

import numpy as np
import pytest
import torch

from oracle.stylesinger_oracle import NoiseSource

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "baseline"))
sys.path.insert(0, os.path.join(REPO, "tools"))
T = 8
SEED = 2024


@contextlib.contextmanager
def patched_rng(ns):
    o = (torch.randn, torch.randn_like, torch.rand, torch.rand_like)

    def shp(a):
        return tuple(a[0]) if len(a) == 1 and isinstance(a[0], (tuple, list, torch.Size)) else tuple(a)

    torch.randn = lambda *a, **k: ns.randn(shp(a)).to(k.get("device") or "cpu")
    torch.randn_like = lambda x, **k: ns.randn(tuple(x.shape)).to(x.device)
    torch.rand = lambda *a, **k: ns.rand(shp(a)).to(k.get("device") or "cpu")
    torch.rand_like = lambda x, **k: ns.rand(tuple(x.shape)).to(x.device)
    try:
        yield
    finally:
        torch.randn, torch.randn_like, torch.rand, torch.rand_like = o


@pytest.fixture(scope="module")
def ref():
    import ref_harness
    if not ref_harness.available():
        pytest.skip("reference not staged (run __graft_entry__.build() where /root/reference exists)")
    cwd = os.getcwd()
    # fp32 reference for parity: the reference's DEFAULT GPU flags run every conv in TF32 (cudnn.allow_tf32 = True), which by
    # itself moves its mel by ~1.6e-3 from its own fp32 result (measured in round 2) - BASELINE.md section 3 names the
    # allow_tf32 = False figure as the one used for parity
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    r = ref_harness.ReferenceRunner(T=T, device="cuda")
    r.hp["use_nsf"] = False  # NSF draws its noise inside the vocoder with its own RNG use: keep the vocoder deterministic here
    from stylesinger_b200 import synth
    u = synth.make_utterance(0.6, utt_idx=11, ref_frames=60, phones=9)
    r.item = r.item_from_utterance(u)
    # the reference's own outputs for seed SEED (predicted durations: the stock forward_model)
    cap = {}
    orig = r.infer.model.forward

    def spy(*a, **k):
        out = orig(*a, **k)
        cap["ret"] = out
        return out

    r.infer.model.forward = spy
    with torch.no_grad(), patched_rng(NoiseSource(SEED)):
        r.wav_ref = r.infer.forward_model(r.item)
    r.infer.model.forward = orig
    r.ret_ref = {k: v.detach().clone() for k, v in cap["ret"].items() if isinstance(v, torch.Tensor)}
    yield r
    r.close()
    os.chdir(cwd)
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32


def _maxabs(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max())


def _engine(ref):
    from stylesinger_b200.engine import AcousticModel
    from stylesinger_b200.hparams import resolve
    if not hasattr(ref, "engine"):
        sd = {k: v.detach().cpu() for k, v in ref.infer.model.state_dict().items()}  # what load_ckpt put into the reference model
        ref.engine = AcousticModel(sd, resolve(timesteps=T, K_step=T, f0_timesteps=T), "cuda:0")
    return ref.engine


def test_registry_level_dropins_inside_the_reference_model(ref):
    """DIFF_DECODERS['wavenet'] callable, both DDiffNets, FS_ENCODERS / FS_DECODERS 'fft' and get_style replaced one by one
    inside the reference's StyleSinger; the reference's own python sampler loops, duration path and glue stay."""
    from stylesinger_b200 import modules as M
    eng = _engine(ref)
    m = ref.infer.model
    saved = {"dn": m.postdiff.denoise_fn, "g1": m.f0_gen._denoise_fn, "g2": m.f0_gen_inpainte._denoise_fn, "enc": m.encoder,
             "dec": m.decoder, "gs": m.get_style}
    facade = M.StyleSinger(engine=eng, hparams=eng.hp)
    steps = [("denoisers", lambda: (setattr(m.postdiff, "denoise_fn", M.DiffNet(eng)),
                                    setattr(m.f0_gen, "_denoise_fn", M.DDiffNet(eng, 1)),
                                    setattr(m.f0_gen_inpainte, "_denoise_fn", M.DDiffNet(eng, 2)))),
             ("+ fft encoder/decoder", lambda: (setattr(m, "encoder", M.FastspeechEncoder(eng)), setattr(m, "decoder", M.FastspeechDecoder(eng)))),
             ("+ get_style", lambda: setattr(m, "get_style", facade.get_style))]
    try:
        for name, apply in steps:
            apply()
            cap = {}
            orig = m.forward

            def spy(*a, **k):
                out = orig(*a, **k)
                cap["ret"] = out
                return out

            m.forward = spy
            with torch.no_grad(), patched_rng(NoiseSource(SEED)):
                wav = ref.infer.forward_model(ref.item)
            del m.forward
            ret = cap["ret"]
            assert torch.equal(ret["mel2ph"].cpu(), ref.ret_ref["mel2ph"].cpu())
            e_mel = _maxabs(ret["mel_out"], ref.ret_ref["mel_out"])
            e_f0 = _maxabs(ret["f0_denorm"], ref.ret_ref["f0_denorm"])
            e_wav = _maxabs(wav, ref.wav_ref)
            print(f"reference with drop-ins [{name}]: mel L-inf {e_mel:.3e}, f0 {e_f0:.3e} Hz, wav {e_wav:.3e}")
            assert e_mel < 1e-3 and e_f0 < 0.5 and e_wav < 2e-3
    finally:
        m.postdiff.denoise_fn, m.f0_gen._denoise_fn, m.f0_gen_inpainte._denoise_fn = saved["dn"], saved["g1"], saved["g2"]
        m.encoder, m.decoder = saved["enc"], saved["dec"]
        if "get_style" in m.__dict__:
            del m.__dict__["get_style"]


def test_whole_model_dropin_driven_by_the_reference_inference_class(ref):
    """the reference's StyleSingerInfer.forward_model with self.model = modules.StyleSinger."""
    from stylesinger_b200 import modules as M
    from tests.common import engine_noise_from_stream
    eng = _engine(ref)

    class Injected(M.StyleSinger):  # same draws as the reference run: SURVEY A.10 order, sized by the predicted frame count
        def forward(self, *a, **k):
            k["noise"] = lambda fo: engine_noise_from_stream(SEED, T, T, int(fo[-1]), "cuda:0")[0]
            out = super().forward(*a, **k)
            self.last = out
            return out

    model = Injected(engine=eng, hparams=eng.hp)
    saved = ref.infer.model
    ref.infer.model = model
    try:
        with torch.no_grad():
            wav = ref.infer.forward_model(ref.item)
    finally:
        ref.infer.model = saved
    ret = model.last
    assert torch.equal(ret["mel2ph"].cpu(), ref.ret_ref["mel2ph"].cpu())
    e_mel = _maxabs(ret["mel_out"], ref.ret_ref["mel_out"])
    e_wav = _maxabs(wav, ref.wav_ref)
    print(f"reference driver + whole-model drop-in: mel L-inf {e_mel:.3e}, wav {e_wav:.3e}")
    assert e_mel < 1e-3 and e_wav < 2e-3
    for k in ("style", "decoder_inp", "pitch_pred"):
        assert _maxabs(ret[k], ref.ret_ref[k]) < 1e-3, k


def test_vocoder_registry_dropin(ref):
    """a class registered under 'HifiGAN_NSF' in the reference's vocoder registry, built from the
    reference's own checkpoint directory, serving the reference's forward_model."""
    from tasks.tts.vocoder_infer import base_vocoder as BV

    from stylesinger_b200 import formats
    from stylesinger_b200.modules import HifiGAN as HifiGANB200
    saved_cls = BV.REGISTERED_VOCODERS["HifiGAN_NSF"]

    @BV.register_vocoder("HifiGAN_NSF")
    class HifiGAN(BV.BaseVocoder):
        def __init__(self):
            sd, cfg, _ = formats.load_vocoder_checkpoint(ref.hp["vocoder_ckpt"])
            self.v = HifiGANB200(sd, cfg, "cuda:0", use_nsf=ref.hp.get("use_nsf"))

        def spec2wav(self, mel, **kwargs):
            return self.v.spec2wav(mel, **kwargs)

    saved_voc = ref.infer.vocoder
    try:
        ref.infer.vocoder = BV.get_vocoder_cls(ref.hp)()
        assert type(ref.infer.vocoder) is HifiGAN
        with torch.no_grad(), patched_rng(NoiseSource(SEED)):
            wav = ref.infer.forward_model(ref.item)
    finally:
        ref.infer.vocoder = saved_voc
        BV.REGISTERED_VOCODERS["HifiGAN_NSF"] = saved_cls
    e = _maxabs(wav, ref.wav_ref)
    print(f"reference model + registered B200 vocoder: wav L-inf {e:.3e}")
    assert wav.shape == ref.wav_ref.shape and e < 1e-3
