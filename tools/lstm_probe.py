"""Drive the LSTM encoder (12 partials x 160 frames): ncu target (profiles/r02_ncu_lstm.md) and A/B timing of the kernel variants.
usage: lstm_probe.py [reps] [mode ...]   modes: async2 (default kernel), barrier2 (SSB_LSTM_CLUSTER_BARRIER=1), async4 (SSB_LSTM_KSPLIT4=1)"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import numpy as np
from stylesinger_b200.engine import LstmEncoder


def synthetic_weights(seed, hidden=256, n_in=40, layers=3):
    rs = np.random.RandomState(seed)
    u = lambda *shape: rs.uniform(-1 / 16, 1 / 16, size=shape).astype(np.float32)
    sd = {}
    for l in range(layers):
        sd["lstm.weight_ih_l%d" % l] = u(4 * hidden, n_in if l == 0 else hidden)
        sd["lstm.weight_hh_l%d" % l] = u(4 * hidden, hidden)
        sd["lstm.bias_ih_l%d" % l] = u(4 * hidden)
        sd["lstm.bias_hh_l%d" % l] = u(4 * hidden)
    return sd


ENV = {"async2": {}, "barrier2": {"SSB_LSTM_CLUSTER_BARRIER": "1"}, "async4": {"SSB_LSTM_KSPLIT4": "1"}}
enc = LstmEncoder(synthetic_weights(71), "cuda:0")
x = torch.rand(12, 160, 40, generator=torch.Generator().manual_seed(0)).to("cuda:0")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
modes = sys.argv[2:] or list(ENV)
ref = None
for mode in modes:
    for k in ("SSB_LSTM_CLUSTER_BARRIER", "SSB_LSTM_KSPLIT4"):
        os.environ.pop(k, None)
    os.environ.update(ENV[mode])
    for _ in range(3):
        out = enc(x, utt_offsets=[0, 12])
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps):
        out = enc(x, utt_offsets=[0, 12])
    ev[1].record()
    torch.cuda.synchronize()
    ref = out["hidden"].clone() if ref is None else ref
    print(f"{mode}: {ev[0].elapsed_time(ev[1]) / reps:.3f} ms per call (12 partials x 160 frames, 3 layers), "
          f"max |hidden - first mode's| = {float((out['hidden'] - ref).abs().max()):.3e}")
