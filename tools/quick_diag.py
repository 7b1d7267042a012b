"""Wall-clock breadcrumbs of engine start-up and one pass (written line by line, survives a timeout)."""
import os, sys, time
t0 = time.perf_counter()
def mark(msg):
    print(f"{time.perf_counter() - t0:8.2f}s {msg}", flush=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
mark("start")
import torch
mark("torch imported")
from bench import make_workload
from stylesinger_b200 import synth
from stylesinger_b200.engine import pack_batch
from stylesinger_b200.hparams import DEFAULT_VOCODER_CONFIG, resolve
from stylesinger_b200.infer import StyleSingerInfer
mark("package imported")
wl = sys.argv[1] if len(sys.argv) > 1 else "utt10s"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
dev = torch.device("cuda:0")
torch.zeros(1, device=dev); torch.cuda.synchronize()
mark("cuda context")
hp = resolve(timesteps=T, K_step=T, f0_timesteps=T)
sd = synth.acoustic_state_dict(hp, seed=0); vsd = synth.vocoder_state_dict(DEFAULT_VOCODER_CONFIG, seed=0)
mark("synthetic weights")
eng = StyleSingerInfer(hp, dev, sd, vsd, DEFAULT_VOCODER_CONFIG)
mark("engine built")
utts, desc = make_workload(wl, 0, 1)
pb = pack_batch(utts, pin=True).to(dev)
mark(f"batch packed: {pb.total_frames} frames")
m, v = eng.model, eng.vocoder
F_ = pb.total_frames
cond = torch.randn(F_, 256, device=dev); coarse = torch.randn(F_, 80, device=dev).clamp(-6, 0.5)
lo = torch.full((F_,), -1.0, device=dev); hi = torch.full((F_,), 1.0, device=dev)
for rep in range(2):
    m.f0_diffusion(0, cond, lo, hi, pb.frame_offsets, seed=3); torch.cuda.synchronize()
    mark(f"rep{rep}: one f0 diffusion (single stream)")
    mel = m.mel_diffusion(cond, coarse, pb.frame_offsets, seed=2); torch.cuda.synchronize()
    mark(f"rep{rep}: mel diffusion")
    wav = v.generate(mel.clamp(-6, 1.5).contiguous(), hi * 200.0, pb.frame_offsets, seed=4); torch.cuda.synchronize()
    mark(f"rep{rep}: vocoder")
    out = m.forward(pb, seed=1, skip_mel_diffusion=True, want=("coarse_mel", "diff_cond", "f0_denorm")); torch.cuda.synchronize()
    mark(f"rep{rep}: acoustic forward w/o mel (two F0 nets on two streams)")
