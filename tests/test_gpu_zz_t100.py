"""Full ph -> mel forward at the bench's step count (T=100 mel steps + 2 x 100 F0/UV steps) against the golden produced
by the unmodified reference (tools/make_golden.py t100), same injected draws.  Runs after the other GPU files."""
import numpy as np
import pytest

from tests.common import golden
from tests.test_gpu_parity import _maxabs, _run_engine_b1

pytestmark = pytest.mark.gpu


def test_T100_matches_reference_golden():
    g, meta = golden("ref_f32_T100")
    out = _run_engine_b1(meta, meta["T"], meta["seed"])
    assert np.array_equal(out["rq_codes"].cpu().numpy().astype(np.int64), g["rq_codes"])
    e_pitch = _maxabs(out["pitch_pred"], g["pitch_pred"])
    e_mel = _maxabs(out["mel_out"], g["mel_out"])
    print(f"T=100 full forward vs reference golden: pitch_pred {e_pitch:.3e}, mel L-inf {e_mel:.3e}")
    assert e_pitch < 1e-3
    assert e_mel < 1e-3  # BASELINE.json north_star bar
