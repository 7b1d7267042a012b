"""Multi-GPU plumbing: the path shards over independent utterances, one process per GPU.

No collective sits on the data path.  `torch.distributed` (NCCL over NVLink on the GPU box, gloo in the
CPU tests) is used only for the trivial batch scatter (inputs, ~0.4 MB / utterance) and gather (waveforms)
when a single rank owns the request batch, and for the max-over-ranks timing reduction in bench.py.
"""
from typing import List, Optional

import numpy as np
import torch
import torch.distributed as dist

from .engine import PackedBatch, pack_batch


def lpt_assign(lengths, world: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of utterances to ranks, balancing the total frame count
    (cost is ~linear in frames x diffusion steps).  Returns, per rank, the utterance indices it owns."""
    order = np.argsort(-np.asarray(lengths, dtype=np.float64), kind="stable")
    loads, bins = [0.0] * world, [[] for _ in range(world)]
    for i in order:
        r = int(np.argmin(loads))
        bins[r].append(int(i))
        loads[r] += float(lengths[i])
    return bins


def empty_batch() -> PackedBatch:
    """The shard of a rank that received no utterance (fewer utterances than ranks)."""
    z = np.zeros(1, np.int32)
    return PackedBatch(0, z, z.copy(), z.copy(), {}, False)


def _p2p_device(device):
    return torch.device(device) if device is not None else torch.device("cpu")


def scatter_utterances(utts: Optional[List[dict]], src: int = 0, device=None, pin: bool = False):
    """Rank `src` owns `utts` (list of per-utterance CPU tensors, see synth.make_utterance); every rank returns
    (its PackedBatch on the host, the global indices of its utterances).  Metadata goes through
    scatter_object_list, tensors through point-to-point send/recv (NCCL when `device` is a CUDA device)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = _p2p_device(device)
    if rank == src:
        lens = [int(u["mel2ph"].shape[0]) if "mel2ph" in u else int(len(u["txt_tokens"])) for u in utts]
        bins = lpt_assign(lens, world)
        # fewer utterances than ranks leaves some bins empty: those ranks get an empty batch (B = 0) and skip the compute
        packed = [pack_batch([utts[i] for i in b], use_mel2ph=all("mel2ph" in utts[i] for i in b)) if b else empty_batch()
                  for b in bins]
        meta = [{"B": p.B, "ph": p.ph_offsets, "ref": p.ref_offsets, "fr": p.frame_offsets, "idx": b, "pad": p.may_have_pad_frames,
                 "shapes": {k: (tuple(v.shape), str(v.dtype)) for k, v in p.t.items()}} for p, b in zip(packed, bins)]
    else:
        packed, meta = None, [None] * world
    mine = [None]
    dist.scatter_object_list(mine, meta if rank == src else None, src=src)
    m = mine[0]
    if rank == src:
        for r in range(world):
            if r == src:
                continue
            for k in sorted(packed[r].t.keys()):
                dist.send(packed[r].t[k].to(dev), dst=r)
        pb = packed[src]
    else:
        t = {}
        for k in sorted(m["shapes"].keys()):
            shape, dt = m["shapes"][k]
            buf = torch.empty(shape, dtype=getattr(torch, dt.replace("torch.", "")), device=dev)
            dist.recv(buf, src=src)
            t[k] = buf.cpu()
        pb = PackedBatch(m["B"], m["ph"], m["ref"], m["fr"], t, m["pad"])
    if pin and torch.cuda.is_available():
        pb = PackedBatch(pb.B, pb.ph_offsets, pb.ref_offsets, pb.frame_offsets, {k: v.pin_memory() for k, v in pb.t.items()},
                         pb.may_have_pad_frames)
    return pb, m["idx"]


def gather_waveforms(wavs: List[np.ndarray], idx: List[int], n_total: int, dst: int = 0, device=None):
    """Inverse of scatter_utterances for the outputs: rank `dst` returns the list of all waveforms in the
    original utterance order, other ranks return None."""
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = _p2p_device(device)
    metas = [None] * world
    dist.all_gather_object(metas, {"idx": list(idx), "lens": [int(len(w)) for w in wavs]})
    if rank == dst:
        out = [None] * n_total
        for i, w in zip(idx, wavs):
            out[i] = np.asarray(w)
        for r in range(world):
            if r == dst:
                continue
            tot = int(sum(metas[r]["lens"]))
            buf = torch.empty(tot, dtype=torch.float32, device=dev)
            if tot:
                dist.recv(buf, src=r)
            h = buf.cpu().numpy()
            o = 0
            for i, n in zip(metas[r]["idx"], metas[r]["lens"]):
                out[i] = h[o:o + n]
                o += n
        return out
    flat = np.concatenate(wavs) if wavs else np.zeros(0, np.float32)
    if flat.size:
        dist.send(torch.from_numpy(np.ascontiguousarray(flat, dtype=np.float32)).to(dev), dst=dst)
    return None
