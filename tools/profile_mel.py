"""Mel-diffusion stage only, for ncu (--profile-from-start off): python tools/profile_mel.py [workload] [T_profiled]
The profiled region is ONE sampler call with T_profiled steps (default 2: hoisted conditioner GEMM, in_proj, then 2 x 20
layer GEMM pairs + heads), after an unprofiled warm-up call."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from bench import make_workload  # noqa: E402
from stylesinger_b200 import synth  # noqa: E402
from stylesinger_b200.engine import pack_batch  # noqa: E402
from stylesinger_b200.hparams import DEFAULT_VOCODER_CONFIG, resolve  # noqa: E402
from stylesinger_b200.infer import StyleSingerInfer  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "batch64"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
hp = resolve(timesteps=T, K_step=T, f0_timesteps=4)
eng = StyleSingerInfer(hp, dev, synth.acoustic_state_dict(hp, seed=0), synth.vocoder_state_dict(DEFAULT_VOCODER_CONFIG, seed=0),
                       DEFAULT_VOCODER_CONFIG)
utts, _ = make_workload(wl, 0, 1)
pb = pack_batch(utts, pin=True).to(dev)
out = eng.model.forward(pb, seed=1, skip_mel_diffusion=True, want=("coarse_mel", "diff_cond"))
cond, coarse = out["diff_cond"], out["coarse_mel"]
eng.model.mel_diffusion(cond, coarse, pb.frame_offsets, seed=2)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.model.mel_diffusion(cond, coarse, pb.frame_offsets, seed=3)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled mel diffusion:", wl, "T =", T, "frames =", pb.total_frames)
