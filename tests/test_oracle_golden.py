"""Pin oracle/stylesinger_oracle.py against fixtures produced by the UNMODIFIED reference
(tools/make_golden.py; the reference itself has no tests or golden vectors — SURVEY.md §4)."""
import os

import numpy as np
import torch

from oracle import stylesinger_oracle as O
from stylesinger_b200.hparams import DEFAULT_VOCODER_CONFIG
from tests.common import GOLDEN, acoustic_sd, golden, hp_for, oracle_forward, utt_from_meta, vocoder_sd

TOL = 2e-5  # fp32 CPU, same op order up to BLAS blocking


def _maxabs(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max())


def test_full_forward_T4_matches_reference():
    g, meta = golden("ref_small_T4")
    hp = hp_for(meta["T"])
    r, ns = oracle_forward(utt_from_meta(meta), hp, meta["seed"])
    assert [tuple(x[1]) for x in meta["noise_log"]] == [x[1] for x in ns.log]  # same RNG draw sequence (A.10)
    assert [x[0] for x in meta["noise_log"]] == [x[0] for x in ns.log]
    assert np.array_equal(r["rq_codes"][0].numpy(), g["rq_codes"])  # RVQ indices: bit-exact
    for k in ["style", "pitch_pred", "decoder_inp", "coarse_mel", "mel_out"]:
        assert _maxabs(r[k][0].numpy(), g[k]) < TOL, k
    assert _maxabs(r["f0_denorm"][0].numpy(), g["f0_denorm"]) < 1e-3  # Hz


def test_duration_path_matches_reference():
    g, meta = golden("ref_small_T4")
    hp = hp_for(meta["T"])
    r, _ = oracle_forward(utt_from_meta(meta), hp, meta["seed"] + 1, use_mel2ph=False)
    assert np.array_equal(r["mel2ph"][0].numpy(), g["dur_mel2ph"])  # integer path: bit-exact
    assert _maxabs(r["dur"][0].numpy(), g["dur_logdur"]) < TOL
    assert _maxabs(r["mel_out"][0].numpy(), g["dur_mel_out"]) < TOL


def test_denoisers_match_reference():
    g, meta = golden("ref_small_T4")
    hp = hp_for(meta["T"])
    sd = acoustic_sd()
    cond = torch.from_numpy(g["dn_cond"])[None]
    with torch.no_grad():
        e1 = O.diffnet(torch.from_numpy(g["dn_spec"])[None, None], torch.tensor([meta["T"] - 1]), cond, sd, hp)
        f0 = torch.from_numpy(g["dd_f0"])[None, None]
        uv = torch.from_numpy(g["dd_uv"])[None]
        e2 = O.ddiffnet(f0, uv, torch.tensor([1]), cond, sd, hp, "gm_diffnet.")
        e3 = O.ddiffnet(f0, uv, torch.tensor([0]), cond, sd, hp, "gm_diffnet_inpainte.")
    assert _maxabs(e1[0, 0].numpy(), g["dn_out"]) < TOL
    assert _maxabs(e2[0].numpy(), g["dd_out"]) < TOL
    assert _maxabs(e3[0].numpy(), g["dd_out_inp"]) < TOL


def test_T25_sampler_matches_reference():
    g, meta = golden("ref_f64_T25")
    hp = hp_for(meta["T"])
    r, _ = oracle_forward(utt_from_meta(meta), hp, meta["seed"])
    assert np.array_equal(r["rq_codes"][0].numpy(), g["rq_codes"])
    assert _maxabs(r["mel_out"][0].numpy(), g["mel_out"]) < 5e-5
    assert _maxabs(r["pitch_pred"][0].numpy(), g["pitch_pred"]) < TOL


def test_T100_sampler_matches_reference():
    """The bench's step count (T=100 mel steps, 2 x 100 F0 steps): 300 chained sampler steps of the oracle against
    the unmodified reference on the same injected draws."""
    g, meta = golden("ref_f32_T100")
    hp = hp_for(meta["T"])
    r, _ = oracle_forward(utt_from_meta(meta), hp, meta["seed"])
    assert np.array_equal(r["rq_codes"][0].numpy(), g["rq_codes"])
    assert _maxabs(r["pitch_pred"][0].numpy(), g["pitch_pred"]) < TOL
    assert _maxabs(r["f0_denorm"][0].numpy(), g["f0_denorm"]) < 1e-2  # Hz
    assert _maxabs(r["mel_out"][0].numpy(), g["mel_out"]) < 1e-4


def test_vocoder_matches_reference():
    g, meta = golden("ref_vocoder_f24")
    ns = O.NoiseSource(meta["seed"] + 5)
    with torch.no_grad():
        w = O.spec2wav(g["mel"], g["f0"], vocoder_sd(), DEFAULT_VOCODER_CONFIG, ns)
        w2 = O.spec2wav(g["mel"], None, vocoder_sd(), DEFAULT_VOCODER_CONFIG, ns)
    assert [tuple(x[1]) for x in meta["noise_log"]] == [x[1] for x in ns.log][:3]
    assert _maxabs(w, g["wav"]) < TOL
    assert _maxabs(w2, g["wav_nof0"]) < TOL


def test_schedules_match_reference_buffers_at_sweep_T():
    """The oracle's and the product's schedule tables are written independently of each other; both must reproduce the
    reference's registered buffers bit for bit at every T of BASELINE.json configs[4] (tools/make_golden.py sched)."""
    from stylesinger_b200 import schedules as S
    g = np.load(os.path.join(GOLDEN, "ref_schedules.npz"))
    hp = hp_for(4)
    for T in [int(t) for t in g["Ts"]]:
        for net, mb in (("mel", hp["max_beta"]), ("f0", hp["f0_max_beta"])):
            o = O._gauss_tables(T, mb)
            p = S.gaussian_schedule(T, mb)
            for k in p:
                ref = g[f"{net}_T{T}_{k}"]
                assert np.array_equal(o[k].numpy(), ref), (net, T, k, "oracle")
                assert np.array_equal(p[k], ref), (net, T, k, "product")
        om, pm = O._multi_tables(T, hp["f0_max_beta"]), S.multinomial_schedule(T, hp["f0_max_beta"])
        for k in pm:
            ref = g[f"f0_T{T}_{k}"]
            assert np.array_equal(om[k].numpy(), ref), (T, k, "oracle")
            assert np.array_equal(pm[k], ref), (T, k, "product")


def test_plms_sampler_matches_reference():
    """f2: oracle restatement of p_sample_plms + the pndm_speedup loop vs the reference run (tools/make_golden.py plms)."""
    g, meta = golden("ref_plms_T100_i10")
    hp = hp_for(meta["T"])
    ns = O.NoiseSource(meta["seed"] + 1)
    with torch.no_grad():
        mel = O.mel_diffusion_sample_plms(torch.from_numpy(g["cond"])[None], torch.from_numpy(g["coarse"])[None], acoustic_sd(), hp,
                                          ns, meta["interval"])
    assert [[k, list(sh)] for k, sh in ns.log] == meta["noise_log"]
    err = _maxabs(mel[0].numpy(), g["mel"])
    scale = float(np.abs(g["mel"]).max())  # the extrapolating multistep update amplifies the synthetic denoiser's output
    assert err < TOL * max(1.0, scale), (err, scale)
