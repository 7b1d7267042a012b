"""Import shim for the UNMODIFIED reference at /root/reference (build container only).

Used exclusively by tools/make_golden.py to generate the fixtures under tests/golden/.
Nothing in tests/, bench.py or the package imports this at run time on the GPU box
(the reference does not travel there).  Shims follow SURVEY.md §8(c): they only satisfy
import-time dependencies that are unused on the hot path; no hot-path arithmetic is touched.
"""
import os
import sys
import types

REF = os.environ.get("STYLESINGER_REF", "/root/reference")


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install(T=100, f0_T=None):
    """chdir to the reference, stub unused imports, load hparams. Returns the hparams dict."""
    if not os.path.isdir(REF):
        raise RuntimeError(f"reference not present at {REF}")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    os.chdir(REF)
    for n in ["librosa", "librosa.filters", "pycwt", "pycwt.wavelet", "chardet", "pyloudnorm",
              "matplotlib", "matplotlib.pyplot", "resemblyzer", "parselmouth"]:
        if n not in sys.modules:
            _stub(n)
    sys.modules["librosa"].filters = sys.modules["librosa.filters"]
    sys.modules["pycwt"].wavelet = sys.modules["pycwt.wavelet"]
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].use = lambda *a, **k: None
    import scipy.signal
    import scipy.signal.windows
    if not hasattr(scipy.signal, "kaiser"):
        scipy.signal.kaiser = scipy.signal.windows.kaiser
    from utils.hparams import set_hparams, hparams
    saved = sys.argv
    sys.argv = [saved[0]]
    try:
        set_hparams(config="egs/stylesinger.yaml", exp_name="", print_hparams=False)
    finally:
        sys.argv = saved
    hparams["timesteps"] = hparams["K_step"] = T
    hparams["f0_timesteps"] = f0_T if f0_T is not None else T
    return hparams
