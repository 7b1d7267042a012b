"""Batch test path over the reference's on-disk formats: reads a released acoustic checkpoint
directory, a HiFi-GAN vocoder directory and a binarised IndexedDataset, runs ph -> mel -> wav on the B200 engine in
ragged batches and writes one wav per item.  Equivalent of `python tasks/run.py --config ... --infer` of the reference
(tasks/StyleSinger/stylesinger.py:168-275, which asserts B=1).

    python tools/infer_dataset.py --ckpt checkpoints/StyleSinger --vocoder checkpoints/hifigan --data data/binary/x/test \
        --out infer_out [--batch 64] [--T 100] [--limit N] [--use-gt-dur]

Like the reference's test_step (tasks/StyleSinger/stylesinger.py:177-180 with `use_gt_dur: false` in egs/stylesinger.yaml) the
durations come from the duration predictor unless --use-gt-dur is given (then the items' ground-truth mel2ph is fed).
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ckpt", required=True, help="work dir with model_ckpt_steps_*.ckpt, or one checkpoint file")
    ap.add_argument("--vocoder", required=True, help="dir with config.yaml + model_ckpt_steps_*.ckpt (or config.json + generator_v1)")
    ap.add_argument("--data", required=True, help="IndexedDataset prefix (<prefix>.idx / <prefix>.data)")
    ap.add_argument("--out", required=True)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--T", type=int, default=100)
    ap.add_argument("--limit", type=int, default=0)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--use-gt-dur", action="store_true", help="feed the items' ground-truth mel2ph (reference hparam use_gt_dur)")
    args = ap.parse_args()

    from scipy.io import wavfile
    from stylesinger_b200 import formats
    from stylesinger_b200.hparams import resolve
    from stylesinger_b200.infer import StyleSingerInfer

    hp = resolve(timesteps=args.T, K_step=args.T, f0_timesteps=args.T)
    sd, path = formats.load_state_dict(args.ckpt, "model")
    vsd, vcfg, vpath = formats.load_vocoder_checkpoint(args.vocoder)
    print(f"| acoustic checkpoint {path} ({len(sd)} tensors); vocoder {vpath}")
    eng = StyleSingerInfer(hp, None, sd, vsd, vcfg)
    os.makedirs(args.out, exist_ok=True)
    sr = int(hp.get("audio_sample_rate", 48000))
    with formats.IndexedDatasetReader(args.data) as ds:
        n = len(ds) if args.limit <= 0 else min(args.limit, len(ds))
        # length-sorted batches: similar lengths share a batch (the engine is ragged, this only balances tile counts)
        order = sorted(range(n), key=lambda i: len(ds[i]["mel"]))
        for b0 in range(0, n, args.batch):
            idx = order[b0:b0 + args.batch]
            utts = [formats.item_to_utterance(ds[i], hp, with_mel2ph=args.use_gt_dur) for i in idx]
            wavs = eng.infer_batch(utts, seed=args.seed + b0, use_mel2ph=args.use_gt_dur)
            for i, u, w in zip(idx, utts, wavs):
                name = str(u.get("item_name") or f"item{i}")
                wavfile.write(os.path.join(args.out, name + ".wav"), sr, np.asarray(w, np.float32))
            print(f"| {min(b0 + args.batch, n)}/{n} items", flush=True)


if __name__ == "__main__":
    main()
