import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import cpu_reference_pass
print("cpu_count", os.cpu_count(), "torch threads default", torch.get_num_threads(), flush=True)
for th in (8, 16, 32, 64):
    if th > (os.cpu_count() or 1): break
    cpu_reference_pass(0.3, 2, th)
    t0 = time.perf_counter()
    f, dt = cpu_reference_pass(0.5, 20, th)
    print(f"threads={th}: 0.5 s utt ({f} frames), T=20: {dt:.2f} s -> est T=100: {dt*5:.1f} s", flush=True)
