"""The reference-side measurements BASELINE.md section 3 asks for, on the box this runs on (UNMODIFIED reference through
baseline/ref_harness.py; nothing of this repo's engine is on the timed path).

    python tools/baseline_arms.py --out profiles/r02_baseline_arms.json [--skip-cpu] [--skip-batch]

* CPU reference: configs[1] (10 s utterance, T=100, full ph->wav, B=1), N host threads, median of 3, plus a 1-thread
  pass on configs[0]'s length (4 s) and the same 4 s length at N threads (so the two are comparable);
* GPU-PyTorch reference (the denominator of north_star's >= 10x target): same modules on cuda:0, eager, default
  backend flags (cuDNN TF32 convs on), B=1: configs[1] median of 3; the allow_tf32=False figure; the B=1 loop over the
  64 utterances of configs[2] (how the reference's inference path serves a batch: tasks/StyleSinger/stylesinger.py:168
  asserts B=1), and the padded-batch forward of configs[2] through StyleSinger.forward (ph->mel) + a B=1 vocoder loop.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "baseline"), os.path.join(REPO, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)


def med(xs):
    return float(np.median(xs))


def cpu_part(T, threads):
    import ref_harness
    out = {"host_logical_cores": os.cpu_count(), "threads": threads, "torch": torch.__version__}
    r = ref_harness.ReferenceRunner(T=T, device="cpu", threads=threads)
    r.timed_pass(0.5)
    ps = [r.timed_pass(10.0) for _ in range(3)]
    out["utt10s"] = {"frames": ps[0][0], "s_median": med([p[1] for p in ps]), "s_all": [p[1] for p in ps],
                     "frames_per_s": ps[0][0] / med([p[1] for p in ps]), "rtf": med([p[1] for p in ps]) / 10.0}
    f4, t4 = r.timed_pass(4.0)
    out["utt4s"] = {"frames": f4, "s": t4, "frames_per_s": f4 / t4, "threads": threads}
    torch.set_num_threads(1)
    f1, t1 = r.timed_pass(4.0)
    out["utt4s_1thread"] = {"frames": f1, "s": t1, "frames_per_s": f1 / t1, "threads": 1}
    torch.set_num_threads(threads)
    r.close()
    return out


def pad(xs, v=0):
    L = max(x.shape[0] for x in xs)
    return torch.stack([torch.cat([x, x.new_full((L - x.shape[0],) + tuple(x.shape[1:]), v)]) for x in xs])


def gpu_part(T, do_batch):
    import ref_harness
    from stylesinger_b200 import synth
    out = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__,
           "flags": {"cudnn.allow_tf32": torch.backends.cudnn.allow_tf32, "matmul.allow_tf32": torch.backends.cuda.matmul.allow_tf32,
                     "cudnn.benchmark": torch.backends.cudnn.benchmark}}
    r = ref_harness.ReferenceRunner(T=T, device="cuda")
    r.timed_pass(1.0)
    r.timed_pass(10.0)
    ps = [r.timed_pass(10.0) for _ in range(3)]
    out["utt10s"] = {"frames": ps[0][0], "s_median": med([p[1] for p in ps]), "frames_per_s": ps[0][0] / med([p[1] for p in ps])}
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    r.timed_pass(10.0)
    ps = [r.timed_pass(10.0) for _ in range(3)]
    out["utt10s_allow_tf32_false"] = {"frames": ps[0][0], "s_median": med([p[1] for p in ps]),
                                      "frames_per_s": ps[0][0] / med([p[1] for p in ps])}
    torch.backends.cudnn.allow_tf32 = True
    if do_batch:
        secs = synth.batch_seconds(64, seed=1234)
        utts = [synth.make_utterance(float(secs[i]), utt_idx=i) for i in range(64)]
        frames = int(sum(len(u["mel2ph"]) for u in utts))
        # (a) the reference's inference path: one utterance at a time
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for u in utts:
            r.forward_model(r.item_from_utterance(u), u["mel2ph"].numpy())
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out["batch64_b1_loop"] = {"frames": frames, "s": dt, "frames_per_s": frames / dt}
        # (b) padded batch through StyleSinger.forward (ph -> mel), vocoder per utterance
        try:
            dev = "cuda"
            kw = dict(mel2ph=pad([u["mel2ph"] for u in utts]).to(dev), spk_embed=torch.stack([u["spk_embed"] for u in utts]).to(dev),
                      emo_embed=torch.stack([u["emo_embed"] for u in utts]).to(dev), ref_mels=pad([u["ref_mels"] for u in utts]).to(dev),
                      ref_f0=pad([u["ref_f0"] for u in utts]).to(dev), global_steps=320000, infer=True,
                      note=pad([u["note"] for u in utts]).to(dev), note_dur=pad([u["note_dur"] for u in utts]).to(dev),
                      note_type=pad([u["note_type"] for u in utts]).to(dev))
            txt = pad([u["txt_tokens"] for u in utts]).to(dev)
            res = []
            for it in range(2):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                with torch.no_grad():
                    o = r.infer.model(txt, **kw)
                torch.cuda.synchronize()
                t_mel = time.perf_counter() - t0
                mel = o["mel_out"].cpu().numpy()
                f0 = o["f0_denorm"].cpu().numpy()
                t0 = time.perf_counter()
                for b, u in enumerate(utts):
                    n = len(u["mel2ph"])
                    r.infer.vocoder.spec2wav(np.clip(mel[b, :n], -6, 1.5), f0=f0[b, :n])
                torch.cuda.synchronize()
                res.append((t_mel, time.perf_counter() - t0))
            t_mel, t_voc = res[-1]
            out["batch64_padded"] = {"frames": frames, "padded_frames": int(64 * max(len(u["mel2ph"]) for u in utts)),
                                     "s_ph2mel": t_mel, "s_vocoder_b1_loop": t_voc, "s": t_mel + t_voc,
                                     "frames_per_s": frames / (t_mel + t_voc),
                                     "note": "padded batching is not B=1-equivalent in the reference (SURVEY section 7); timing only"}
        except Exception as e:
            out["batch64_padded"] = {"unavailable": f"{type(e).__name__}: {e}"}
    r.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--T", type=int, default=100)
    ap.add_argument("--threads", type=int, default=min(os.cpu_count() or 1, 16))
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-gpu", action="store_true")
    ap.add_argument("--skip-batch", action="store_true")
    a = ap.parse_args()
    cwd = os.getcwd()
    res = {"T": a.T, "what": "unmodified reference, synthetic seed-0 checkpoints, bench.py's synthetic utterances (explicit mel2ph)"}
    if not a.skip_gpu and torch.cuda.is_available():
        res["gpu_pytorch"] = gpu_part(a.T, not a.skip_batch)
    if not a.skip_cpu:
        # the CPU run hides CUDA from the reference's vocoder wrapper (see ReferenceRunner): do it after the GPU part
        res["cpu"] = cpu_part(a.T, a.threads)
    os.chdir(cwd)
    s = json.dumps(res, indent=1)
    print(s)
    if a.out:
        with open(a.out, "w") as f:
            f.write(s + "\n")


if __name__ == "__main__":
    main()
