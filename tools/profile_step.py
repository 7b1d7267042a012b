"""One profiled ph->mel->wav pass for ncu (--profile-from-start off): python tools/profile_step.py [workload] [T]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_workload
from stylesinger_b200 import synth
from stylesinger_b200.engine import pack_batch
from stylesinger_b200.hparams import DEFAULT_VOCODER_CONFIG, resolve
from stylesinger_b200.infer import StyleSingerInfer

wl = sys.argv[1] if len(sys.argv) > 1 else "utt10s"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
dev = torch.device("cuda:0")
hp = resolve(timesteps=T, K_step=T, f0_timesteps=T)
eng = StyleSingerInfer(hp, dev, synth.acoustic_state_dict(hp, seed=0), synth.vocoder_state_dict(DEFAULT_VOCODER_CONFIG, seed=0), DEFAULT_VOCODER_CONFIG)
utts, _ = make_workload(wl, 0, 1)
pb = pack_batch(utts, pin=True).to(dev)
eng.run_device(pb, seed=0)  # warm-up (unprofiled)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.run_device(pb, seed=1)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step:", wl, "T =", T, "frames =", pb.total_frames)
