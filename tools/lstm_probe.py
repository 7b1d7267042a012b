"""Drive the LSTM encoder (12 partials x 160 frames): ncu target (profiles/r02_ncu_lstm.md) and A/B timing of the h exchange
(SSB_LSTM_CLUSTER_BARRIER=1: one barrier.cluster per frame; default: mbarrier-signalled st.async)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import frontend_oracle as FO  # seeded synthetic state_dict only; a tool, not the product
from stylesinger_b200.engine import LstmEncoder

enc = LstmEncoder(FO.emotion_encoder_weights(71), "cuda:0")
x = torch.rand(12, 160, 40, device="cuda:0")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
res = {}
for mode in ("0", "1"):
    os.environ["SSB_LSTM_CLUSTER_BARRIER"] = mode
    for _ in range(3):
        out = enc(x, utt_offsets=[0, 12])
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps):
        out = enc(x, utt_offsets=[0, 12])
    ev[1].record()
    torch.cuda.synchronize()
    res[mode] = out["hidden"].clone()
    print(f"SSB_LSTM_CLUSTER_BARRIER={mode}: {ev[0].elapsed_time(ev[1]) / reps:.3f} ms per call (12 partials x 160 frames, 3 layers)")
print("max |async - barrier| =", float((res["0"] - res["1"]).abs().max()))
