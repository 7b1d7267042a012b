"""Per-stage device times (CUDA events) of one ph->mel->wav pass.
Usage: python tools/stage_times.py [utt10s|batch8|batch64] [T] [fast]   (fast: default engine mode only, no FFMA comparison)"""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import make_workload
from stylesinger_b200 import synth
from stylesinger_b200._lib import lib
from stylesinger_b200.engine import pack_batch
from stylesinger_b200.hparams import DEFAULT_VOCODER_CONFIG, resolve
from stylesinger_b200.infer import StyleSingerInfer

wl = sys.argv[1] if len(sys.argv) > 1 else "utt10s"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
FAST = len(sys.argv) > 3 and sys.argv[3] == "fast"
dev = torch.device("cuda:0")
hp = resolve(timesteps=T, K_step=T, f0_timesteps=T)
eng = StyleSingerInfer(hp, dev, synth.acoustic_state_dict(hp, seed=0), synth.vocoder_state_dict(DEFAULT_VOCODER_CONFIG, seed=0), DEFAULT_VOCODER_CONFIG)
utts, desc = make_workload(wl, 0, 1)
pb = pack_batch(utts, pin=True).to(dev)
F_ = pb.total_frames
m, v = eng.model, eng.vocoder


def timed(fn, n=2):
    fn(); torch.cuda.synchronize()
    l0 = lib.ssb_launch_count()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for _ in range(n): r = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n, (time.perf_counter() - t0) * 1000 / n, (lib.ssb_launch_count() - l0) // n, r

res = {}
for tag in (("persist",) if FAST else ("persist", "tc", "simt")):
    m.set_tensor_cores(tag != "simt")
    m.set_persistent(tag == "persist")
    ms, wall, nl, out = timed(lambda: m.forward(pb, seed=1, skip_mel_diffusion=True, want=("coarse_mel", "diff_cond", "f0_denorm")))
    res[f"{tag}.acoustic_wo_mel(enc+style+2xF0+dec)"] = (round(ms, 2), round(wall, 2), nl)
    cond, coarse = out["diff_cond"], out["coarse_mel"]
    ms, wall, nl, mel = timed(lambda: m.mel_diffusion(cond, coarse, pb.frame_offsets, seed=2))
    res[f"{tag}.mel_diffusion"] = (round(ms, 2), round(wall, 2), nl)
    lo = torch.full((F_,), -1.0, device=dev); hi = torch.full((F_,), 1.0, device=dev)
    ms, wall, nl, _ = timed(lambda: m.f0_diffusion(0, cond, lo, hi, pb.frame_offsets, seed=3))
    res[f"{tag}.one_f0_diffusion"] = (round(ms, 2), round(wall, 2), nl)
m.set_tensor_cores(True); m.set_persistent(True)
melc = mel.clamp(-6, 1.5).contiguous(); f0 = out["f0_denorm"]
for tcv in ((True,) if FAST else (True, False)):
    v.set_tensor_cores(tcv)
    ms, wall, nl, _ = timed(lambda: v.generate(melc, f0, pb.frame_offsets, seed=4))
    res["vocoder." + ("tc" if tcv else "simt")] = (round(ms, 2), round(wall, 2), nl)
v.set_tensor_cores(True)
print(json.dumps({"workload": wl, "frames": F_, "T": T, "stage: (gpu_ms, wall_ms, launches)": res}))
