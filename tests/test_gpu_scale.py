"""Parity at the sizes bench.py runs ("next round" item 2): BASELINE.json configs[1] end to end, the CTA-pair
tcgen05 kernels through chained T=100 samplers on a >= 20 k-frame ragged batch, the pair-kernel variants by name, the
mel post-process glue, RVQ at configs[2] scale, and the reference-named module facades on the real engine.

Tolerances: mel L-inf < 1e-3 (north_star); waveform from the ORACLE's mel < 1e-3; RVQ codes bit-exact.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import stylesinger_oracle as O
from stylesinger_b200.hparams import DEFAULT_VOCODER_CONFIG
from tests.common import (acoustic_engine, acoustic_sd, engine_noise_from_stream, hp_for, oracle_forward, vocoder_engine,
                          vocoder_sd)

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
BATCH_LENS = [2900, 1700, 2999, 800, 2300, 1950, 2450, 3000, 1300, 1111]  # 20 510 frames, 165 row tiles -> pair kernels


def _maxabs(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max())


def _variants():
    from stylesinger_b200._lib import variant_launches
    return variant_launches()


def _delta(before, after):
    return {k: v - before.get(k, 0) for k, v in after.items() if v - before.get(k, 0) > 0}


class ListNoise:
    """Noise source that replays prepared tensors in call order (oracle side of the batch-scale sampler tests)."""

    def __init__(self, tensors):
        self.t, self.i, self.log, self.record = list(tensors), 0, [], None

    def _next(self, shape):
        x = self.t[self.i]
        self.i += 1
        assert tuple(x.shape) == tuple(shape), (tuple(x.shape), tuple(shape))
        return x

    def randn(self, shape):
        return self._next(shape)

    def rand(self, shape):
        return self._next(shape)


# ---------------------------------------------------------------------------------------------------
# (a) BASELINE.json configs[1]: one 10 s utterance, T=100, host in -> wav out, against the oracle
def test_config1_utt10s_T100_forward_model_vs_oracle():
    from stylesinger_b200 import formats, synth
    from stylesinger_b200.infer import StyleSingerInfer
    T = 100
    hp = hp_for(T)
    u = synth.make_utterance(10.0, utt_idx=0)
    Fr = int(u["mel2ph"].shape[0])
    assert Fr == 1875
    # reference-format item (raw Hz f0, as preprocess_input produces); both sides see the norm_interp_f0 of it
    item = {"ph_token": u["txt_tokens"].numpy(), "note": u["note"].numpy(), "note_dur": u["note_dur"].numpy(),
            "note_type": u["note_type"].numpy(), "spk_embed": u["spk_embed"].numpy(), "emo_embed": u["emo_embed"].numpy(),
            "mel": u["ref_mels"].numpy(), "f0": np.exp2(u["ref_f0"].numpy().astype(np.float64)).astype(np.float32),
            "mel2ph": u["mel2ph"].numpy()}
    f0n, _ = formats.norm_interp_f0(item["f0"])
    uo = dict(u, ref_f0=torch.from_numpy(f0n))
    seed = 4242
    r, _ = oracle_forward(uo, hp, seed)
    mel_o, f0_o = O.postprocess_mel(r["mel_out"][0].numpy(), r["f0_denorm"][0].numpy(), hp)
    ns2 = O.NoiseSource(seed + 1)
    ns2.record = []
    with torch.no_grad():
        wav_o = O.spec2wav(mel_o, f0_o, vocoder_sd(), DEFAULT_VOCODER_CONFIG, ns2)
    ini = ns2.record[0].clone()
    ini[:, 0] = 0
    voc_noise = {"rand_ini": ini.to(DEV).contiguous(), "src_noise": ns2.record[1][0].contiguous().to(DEV)}

    eng = StyleSingerInfer(hp, DEV, acoustic_sd(), vocoder_sd(), DEFAULT_VOCODER_CONFIG)
    noise, _ = engine_noise_from_stream(seed, T, T, Fr, DEV)
    before = _variants()
    wav, mel = eng.forward_model(item, noise=noise, voc_noise=voc_noise, return_mel=True)
    ran = _delta(before, _variants())
    e_mel = _maxabs(mel, r["mel_out"][0])
    e_wav_chain = _maxabs(wav, wav_o)
    # vocoder alone on the oracle's post-processed mel / f0 (isolates HiFi-GAN from the NSF phase drift that a 1e-4 Hz f0
    # difference accumulates over 480 000 samples)
    wav2 = eng.vocoder.generate(torch.from_numpy(mel_o).to(DEV), torch.from_numpy(f0_o).to(DEV), np.array([0, Fr], np.int32),
                                rand_ini=voc_noise["rand_ini"], src_noise=voc_noise["src_noise"])
    e_wav = _maxabs(wav2, wav_o)
    print(f"configs[1] 10 s / T=100: mel L-inf {e_mel:.3e}, wav (oracle mel) {e_wav:.3e}, wav (own mel, chained) {e_wav_chain:.3e}; kernels {ran}")
    assert len(wav) == Fr * 256 and wav.dtype == np.float32
    assert e_mel < 1e-3
    assert e_wav < 1e-3
    assert e_wav_chain < 2e-2


# ---------------------------------------------------------------------------------------------------
# (b) chained T=100 samplers on a 20 k-frame ragged batch: CTA-pair tcgen05 kernels vs the fp32 FFMA path and the oracle
def _batch_inputs(seed):
    gen = torch.Generator().manual_seed(seed)
    offs = np.concatenate([[0], np.cumsum(BATCH_LENS)]).astype(np.int32)
    n = int(offs[-1])
    cond = torch.randn(n, 256, generator=gen)
    coarse = (-3 + 0.8 * torch.randn(n, 80, generator=gen)).clamp(-6, 0.5)
    return offs, n, cond, coarse


def _interleaved(on):
    from stylesinger_b200._lib import lib
    return lib.ssb_set_interleaved_layers(1 if on else 0)


def test_mel_sampler_T100_pair_kernels_vs_simt_philox_20k_frames():
    """Three runs of the same T=100 Philox sampler: interleaved dual kernel (opt-in: the gate conv of one utterance group and
    the 1x1 residual conv of the other share a launch), one launch per GEMM (default), and the fp32 FFMA path."""
    T = 100
    m = acoustic_engine(T, 4)
    offs, n, cond, coarse = _batch_inputs(31)
    cond, coarse = cond.to(DEV), coarse.to(DEV)
    out, ran = {}, {}
    try:
        m.set_persistent(False)
        for mode in ("dual", "tc", "simt"):
            m.set_tensor_cores(mode != "simt")
            _interleaved(mode == "dual")
            before = _variants()
            out[mode] = m.mel_diffusion(cond, coarse, offs, None, seed=17).clone()
            ran[mode] = _delta(before, _variants())
    finally:
        m.set_tensor_cores(True)
        m.set_persistent(True)
        _interleaved(False)  # library default
    err = _maxabs(out["tc"], out["simt"])
    err_d = _maxabs(out["dual"], out["tc"])
    print(f"mel sampler T=100, {n} frames, philox: pair-tc vs simt L-inf {err:.3e}, interleaved vs per-GEMM {err_d:.3e}; "
          f"kernels {ran['dual']} | {ran['tc']}")
    assert ran["tc"].get("tc2r<128,GATE>", 0) == T * 20 and ran["tc"].get("tc2<128,RES_SKIP>", 0) == T * 20
    # two lanes, software-pipelined by half a layer: first gate and last 1x1 alone, 2L - 1 interleaved launches in between
    assert ran["dual"].get("tc2d<128,GATE+RES_SKIP>", 0) == T * 39, ran["dual"]
    # (the lone first gate / last 1x1 of a step cover half the batch: here small enough for the single-CTA kernel)
    assert sum(v for k, v in ran["dual"].items() if k.endswith(",GATE>")) == T, ran["dual"]
    assert sum(v for k, v in ran["dual"].items() if k.endswith(",RES_SKIP>")) == T, ran["dual"]
    assert not ran["simt"], ran["simt"]  # the fp32 FFMA path launches no tcgen05 kernel
    assert torch.isfinite(out["dual"]).all() and err < 1e-3 and err_d < 1e-4


def test_mel_sampler_T100_pair_kernels_vs_oracle_two_utterances_of_the_batch():
    T = 100
    hp = hp_for(T)
    m = acoustic_engine(T, 4)
    offs, n, cond, coarse = _batch_inputs(32)
    gen = torch.Generator().manual_seed(99)
    noise = torch.randn(T + 1, n, 80, generator=gen)
    try:
        m.set_persistent(False)
        before = _variants()
        mel = m.mel_diffusion(cond.to(DEV), coarse.to(DEV), offs, noise.to(DEV))
        ran = _delta(before, _variants())
    finally:
        m.set_persistent(True)
    assert ran.get("tc2r<128,GATE>", 0) == T * 20 and ran.get("tc2<128,RES_SKIP>", 0) == T * 20, ran  # default: one launch per GEMM
    worst = 0.0
    for b in (3, 9):  # 800 and 1111 frames
        a, e = int(offs[b]), int(offs[b + 1])
        ln = ListNoise([noise[i, a:e].t().contiguous()[None, None] for i in range(T + 1)])
        with torch.no_grad():
            ref = O.mel_diffusion_sample(cond[None, a:e], coarse[None, a:e], acoustic_sd(), hp, ln)
        worst = max(worst, _maxabs(mel[a:e], ref[0]))
    print(f"mel sampler T=100 inside a {n}-frame batch (pair kernels) vs oracle: L-inf {worst:.3e}")
    assert worst < 1e-3


def test_f0_sampler_T100_pair_kernels_vs_oracle_two_utterances_of_the_batch():
    """The UV half is an argmax over Gumbel-perturbed logits: a logit difference of 1e-5 flips a decision only where the
    margin is below it, which is expected for a handful of the 2 M frame-steps of this batch and is not an error.  Bar:
    >= 99 % of the frames agree in UV and within 1e-3 in normalised f0."""
    T = 100
    hp = hp_for(4, T)
    m = acoustic_engine(4, T)
    offs, n, cond, _ = _batch_inputs(33)
    gen = torch.Generator().manual_seed(7)
    midi = torch.randint(50, 70, (1, 1, n), generator=gen).float()
    lo, hi = O.midi_clip_band(midi)
    g = torch.randn(T + 1, n, generator=gen)
    u = torch.rand(T, n, 2, generator=gen)
    before = _variants()
    z, uv = m.f0_diffusion(1, cond.to(DEV), lo.reshape(n).to(DEV), hi.reshape(n).to(DEV), offs, g.to(DEV), u.to(DEV))
    ran = _delta(before, _variants())
    assert ran.get("tc2r<96,GATE>", 0) == T * 10 and ran.get("tc2<96,RES_SKIP>", 0) == T * 10, ran
    agree, total = 0, 0
    for b in (3, 9):
        a, e = int(offs[b]), int(offs[b + 1])
        Fr = e - a
        draws = [torch.zeros(1, 1, Fr), g[0, a:e].reshape(1, 1, Fr)]  # UV-init draw (unused), z_T
        for i in range(T):
            draws += [g[1 + i, a:e].reshape(1, 1, Fr), u[i, a:e].t().contiguous()[None]]
        with torch.no_grad():
            ref = O.f0_diffusion_sample(cond[a:e].t().contiguous()[None], (lo[:, :, a:e], hi[:, :, a:e]), acoustic_sd(), hp,
                                        "gm_diffnet_inpainte.", ListNoise(draws))
        ok = (uv[a:e].cpu().numpy() == ref[0, :, 1].numpy().astype(np.int32)) & \
             (np.abs(z[a:e].cpu().numpy() - ref[0, :, 0].numpy()) < 1e-3)
        agree += int(ok.sum())
        total += Fr
    print(f"f0 sampler T=100 inside a {n}-frame batch (pair kernels {ran}) vs oracle: {agree}/{total} frames agree")
    assert agree >= 0.99 * total


def test_forward_f0_nets_interleaved_vs_per_gemm_batch():
    """Full forward on a 20 k-frame batch (mel diffusion skipped): the two F0/UV samplers run in lock step with their residual
    layers interleaved on the dual kernel (tc2d<96,...>); same Philox streams as the per-GEMM schedule, so pitch must agree
    except where a Gumbel-argmax UV decision sits within rounding distance of a tie."""
    from stylesinger_b200 import synth
    from stylesinger_b200.engine import pack_batch
    T = 25
    m = acoustic_engine(4, T)
    utts = [synth.make_utterance(L / 187.5, utt_idx=40 + i, frames=L) for i, L in enumerate(BATCH_LENS)]
    pb = pack_batch(utts).to(DEV)
    out, ran = {}, {}
    try:
        for on in (True, False):
            _interleaved(on)
            before = _variants()
            r = m.forward(pb, seed=5, skip_mel_diffusion=True, want=("pitch_pred", "f0_denorm"))
            out[on] = {k: v.clone() for k, v in r.items()}
            ran[on] = _delta(before, _variants())
    finally:
        _interleaved(False)  # library default
    assert ran[True].get("tc2d<96,GATE+RES_SKIP>", 0) == T * 19, ran[True]  # L = 10: 2L - 1 interleaved launches per step
    assert "tc2d<96,GATE+RES_SKIP>" not in ran[False]
    pp_a, pp_b = out[True]["pitch_pred"].cpu().numpy(), out[False]["pitch_pred"].cpu().numpy()
    same_uv = (pp_a[:, 1] > 0) == (pp_b[:, 1] > 0)
    close = np.abs(pp_a[:, 0] - pp_b[:, 0]) < 1e-3
    frac = float((same_uv & close).mean())
    print(f"forward, {pb.total_frames} frames, T_f0={T}: interleaved vs per-GEMM F0 nets agree on {frac * 100:.3f} % of the frames; {ran[True]}")
    assert frac >= 0.995


# ---------------------------------------------------------------------------------------------------
# (c) every CTA-pair variant by name, incl. tc2<64,GENERIC> (295 launches / 4.2 % of the batch64 step, untested in round 1)
@pytest.mark.parametrize("cin,n_out,k,dil,reps,variant", [(256, 512, 3, 4, 1, "tc2r<128,GENERIC>"), (256, 384, 3, 2, 1, "tc2r<96,GENERIC>"),
                                                         (256, 512, 3, 8, 1, "tc2r<128,GENERIC>"), (256, 512, 3, 1, 1, "tc2r<128,GENERIC>"),
                                                         (256, 512, 1, 1, 1, "tc2<128,GENERIC>"), (192, 384, 5, 1, 1, "tc2<96,GENERIC>"),
                                                         (128, 128, 7, 1, 2, "tc2<64,GENERIC>"), (64, 64, 11, 1, 2, "tc2<32,GENERIC>"),
                                                         (128, 128, 3, 1, 1, "tc<64,GENERIC>")])
def test_conv1d_tc_variant_by_name(cin, n_out, k, dil, reps, variant):
    from stylesinger_b200.engine import op_conv1d_tc
    g = torch.Generator().manual_seed(5 + n_out + k)
    lens = [2800, 1500, 2999, 700, 2100, 1900, 2500, 3000, 1234, 2222] * reps + [77]
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    x = torch.randn(int(offs[-1]), cin, generator=g)
    w = torch.randn(n_out, cin, k, generator=g) / (cin * k) ** 0.5
    b = torch.randn(n_out, generator=g)
    before = _variants()
    y = op_conv1d_tc(x.to(DEV), offs, w, b, dilation=dil).cpu()
    ran = _delta(before, _variants())
    assert ran == {variant: 1}, ran
    worst = 0.0
    for i in (0, 3, len(lens) // 2, len(lens) - 1):
        xi = x[offs[i]:offs[i + 1]].t()[None]
        ref = F.conv1d(xi, w, b, padding=dil * (k - 1) // 2, dilation=dil)[0].t()
        worst = max(worst, _maxabs(y[offs[i]:offs[i + 1]], ref))
    print(f"{variant}: {cin}->{n_out} k{k} on {int(offs[-1])} rows, max err {worst:.3e}")
    assert worst < 1e-4


# ---------------------------------------------------------------------------------------------------
# (d) ssb_mel_postprocess against the reference glue (inference/StyleSinger.py:54-58), all-zero frames included
def test_mel_postprocess_matches_oracle_including_zero_frames():
    import ctypes as C

    from stylesinger_b200._lib import check, lib
    hp = hp_for(4)
    gen = torch.Generator().manual_seed(3)
    mel = torch.randn(5000, 80, generator=gen) * 4.0  # plenty of values beyond [-6, 1.5]
    zero_rows = [0, 17, 18, 19, 2500, 4999]
    mel[zero_rows] = 0.0
    mel[100, :] = 0.0
    mel[100, 7] = 1e-30  # not a zero frame
    f0 = torch.rand(5000, generator=gen) * 400
    mel_o, f0_o = O.postprocess_mel(mel.numpy(), f0.numpy(), hp)
    d = mel.clone().to(DEV)
    cnt = torch.zeros(1, dtype=torch.int32, device=DEV)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    check(lib.ssb_mel_postprocess(C.c_void_p(d.data_ptr()), 5000, float(hp["mel_vmin"]), float(hp["mel_vmax"]),
                                  C.c_void_p(cnt.data_ptr()), stream), "ssb_mel_postprocess")
    assert int(cnt.item()) == mel_o.shape[0] == 5000 - len(zero_rows)
    keep = np.abs(mel.numpy()).sum(-1) > 0
    assert np.array_equal(d.cpu().numpy()[keep], mel_o)          # clip is exact
    assert np.array_equal(d.cpu().numpy()[~keep], np.zeros((len(zero_rows), 80), np.float32))
    assert np.array_equal(f0.numpy()[keep], f0_o)


def test_infer_drops_padding_frames_like_the_reference():
    """An explicit mel2ph with trailing zeros (padding frames): the reference drops the frames whose mel is all zero
    before the vocoder (inference/StyleSinger.py:56-62); the engine path must hand back the same number of samples."""
    from stylesinger_b200 import synth
    from stylesinger_b200.engine import pack_batch
    from stylesinger_b200.infer import StyleSingerInfer
    hp = hp_for(4)
    u = synth.make_utterance(0.5, utt_idx=3, ref_frames=40, frames=90, phones=8)
    u["mel2ph"] = torch.cat([u["mel2ph"], torch.zeros(6, dtype=u["mel2ph"].dtype)])
    pb = pack_batch([u], use_mel2ph=True)
    assert pb.may_have_pad_frames
    eng = StyleSingerInfer(hp, DEV, acoustic_sd(), vocoder_sd(), DEFAULT_VOCODER_CONFIG)
    wavs, mels = eng.infer_packed(pb, seed=1, return_mel=True)
    nz = int((np.abs(mels[0]).sum(-1) > 0).sum())
    print("frames with non-zero mel:", nz, "of", mels[0].shape[0], "-> wav samples", len(wavs[0]))
    assert len(wavs[0]) == nz * 256


# ---------------------------------------------------------------------------------------------------
# (e) RVQ at the scale of configs[2]: 64 references x 1125 frames x depth 4 = 288 000 lookups, bit-exact
def test_rvq_codes_bit_exact_at_config2_scale():
    m = acoustic_engine(4)
    gen = torch.Generator().manual_seed(64)
    x = torch.randn(64 * 1125, 256, generator=gen)
    offs = (np.arange(65) * 1125).astype(np.int32)
    q, codes = m.rvq(x.to(DEV), offs)
    with torch.no_grad():
        qo, co = O.rq_quantize(x[None], acoustic_sd())
    mism = int((codes.cpu().numpy().astype(np.int64) != co[0].numpy()).sum())
    print("RVQ mismatching codes of 288000:", mism)
    assert mism == 0
    assert _maxabs(q, qo[0]) < 1e-6


# ---------------------------------------------------------------------------------------------------
# (f) the reference-named facades on the real engine
def test_module_facades_on_the_real_engine_match_the_oracle():
    from stylesinger_b200 import synth
    from stylesinger_b200.modules import HifiGAN, StyleSinger
    T = 4
    hp = hp_for(T)
    specs = [(96, 12, 64, 100), (61, 7, 40, 104)]
    utts = [synth.make_utterance(f / 187.5, utt_idx=i, ref_frames=r, frames=f, phones=p) for f, p, r, i in specs]

    def pad(xs, v=0):
        L = max(x.shape[0] for x in xs)
        return torch.stack([torch.cat([x, x.new_full((L - x.shape[0],) + tuple(x.shape[1:]), v)]) for x in xs])

    model = StyleSinger(engine=acoustic_engine(T), hparams=hp)
    from tests.common import batch_noise
    per = [engine_noise_from_stream(700 + i, T, T, u["mel2ph"].shape[0], DEV)[0] for i, u in enumerate(utts)]
    ret = model(pad([u["txt_tokens"] for u in utts]), mel2ph=pad([u["mel2ph"] for u in utts]),
                spk_embed=torch.stack([u["spk_embed"] for u in utts]), emo_embed=torch.stack([u["emo_embed"] for u in utts]),
                ref_mels=pad([u["ref_mels"] for u in utts]), ref_f0=pad([u["ref_f0"] for u in utts]), global_steps=320000,
                infer=True, note=pad([u["note"] for u in utts]), note_dur=pad([u["note_dur"] for u in utts]),
                note_type=pad([u["note_type"] for u in utts]), noise=batch_noise(per))
    assert tuple(ret["mel_out"].shape) == (2, 96, 80) and tuple(ret["f0_denorm"].shape) == (2, 96)
    for i, u in enumerate(utts):
        r, _ = oracle_forward(u, hp, 700 + i)
        n = u["mel2ph"].shape[0]
        assert _maxabs(ret["mel_out"][i, :n], r["mel_out"][0]) < 1e-3
        assert _maxabs(ret["style"][i, :n], r["style"][0]) < 1e-4
        assert _maxabs(ret["decoder_inp"][i, :n], r["decoder_inp"][0]) < 1e-4
        assert float(ret["mel_out"][i, n:].abs().max()) == 0.0 if n < 96 else True
    # HifiGAN.spec2wav: numpy in / numpy out, no f0 (deterministic) against the oracle
    voc = HifiGAN(engine=vocoder_engine())
    mel = (-3.0 + 0.8 * torch.randn(40, 80, generator=torch.Generator().manual_seed(1))).clamp(-6, 1.5).numpy()
    wav = voc.spec2wav(mel)
    with torch.no_grad():
        ref = O.spec2wav(mel, None, vocoder_sd(), DEFAULT_VOCODER_CONFIG, O.NoiseSource(0))
    assert wav.dtype == np.float32 and wav.shape == (40 * 256,)
    assert _maxabs(wav, ref) < 1e-3


def test_encoder_out_matches_oracle():
    """a1: FastspeechEncoder + NoteEncoder output (requested by round-1 tests but never asserted)."""
    from stylesinger_b200 import synth
    from stylesinger_b200.engine import pack_batch
    T = 4
    hp = hp_for(T)
    u = synth.make_utterance(0.5, utt_idx=9, ref_frames=40, frames=90, phones=13)
    r, _ = oracle_forward(u, hp, 1)
    m = acoustic_engine(T)
    out = m.forward(pack_batch([u]).to(DEV), seed=0, skip_mel_diffusion=True, want=("encoder_out",))
    err = _maxabs(out["encoder_out"], r["encoder_out"][0])
    print("encoder_out L-inf", err)
    assert err < 1e-4


# ---------------------------------------------------------------------------------------------------
# f2: PLMS sampler (pndm_speedup) on the same kernels
def test_plms_sampler_matches_reference_golden_and_oracle():
    from tests.common import golden
    g, meta = golden("ref_plms_T100_i10")
    T, k = meta["T"], meta["interval"]
    m = acoustic_engine(T, 4)
    Fr = g["cond"].shape[0]
    ns = O.NoiseSource(meta["seed"] + 1)
    q = ns.randn((1, 1, 80, Fr))[0, 0].t().contiguous()
    offs = np.array([0, Fr], np.int32)
    scale = float(np.abs(g["mel"]).max())
    errs = {}
    try:
        for tc in (True, False):
            m.set_tensor_cores(tc)
            mel = m.mel_diffusion_plms(torch.from_numpy(g["cond"]).to(DEV), torch.from_numpy(g["coarse"]).to(DEV), offs, k, q.to(DEV))
            errs[tc] = _maxabs(mel, g["mel"])
    finally:
        m.set_tensor_cores(True)
    print(f"PLMS T={T} interval={k}: L-inf vs reference golden tc {errs[True]:.3e}, fp32 FFMA {errs[False]:.3e} (max |mel| {scale:.1f})")
    assert errs[False] < 1e-4 * max(1.0, scale) and errs[True] < 1e-3 * max(1.0, scale)


def test_plms_on_a_ragged_batch_vs_oracle_and_through_forward():
    """Batch of 3 through ssb_mel_diffusion_sample_plms against the B=1 oracle, and hparams['pndm_speedup'] through the whole
    acoustic forward (Philox mode: runs, finite, differs from the DDPM result)."""
    from stylesinger_b200 import synth
    from stylesinger_b200.engine import AcousticModel, pack_batch
    T, k = 20, 5
    hp = hp_for(T)
    m = acoustic_engine(T, 4)
    gen = torch.Generator().manual_seed(8)
    lens = [130, 70, 257]
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    n = int(offs[-1])
    cond = torch.randn(n, 256, generator=gen)
    coarse = (-3 + 0.8 * torch.randn(n, 80, generator=gen)).clamp(-6, 0.5)
    q = torch.randn(n, 80, generator=gen)
    mel = m.mel_diffusion_plms(cond.to(DEV), coarse.to(DEV), offs, k, q.to(DEV))
    for b in range(3):
        a, e = int(offs[b]), int(offs[b + 1])
        with torch.no_grad():
            ref = O.mel_diffusion_sample_plms(cond[None, a:e], coarse[None, a:e], acoustic_sd(), hp,
                                              ListNoise([q[a:e].t().contiguous()[None, None]]), k)
        sc = max(1.0, float(ref.abs().max()))
        assert _maxabs(mel[a:e], ref[0]) < 1e-3 * sc
    hp2 = dict(hp, pndm_speedup=k)
    m2 = AcousticModel(acoustic_sd(), hp2, DEV)
    u = synth.make_utterance(0.5, utt_idx=5, ref_frames=40, frames=90, phones=8)
    pb = pack_batch([u]).to(DEV)
    a = m2.forward(pb, seed=3)["mel_out"].clone()
    b_ = m.forward(pb, seed=3)["mel_out"].clone()
    assert torch.isfinite(a).all() and not torch.equal(a, b_)
