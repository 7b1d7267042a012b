"""Utterance-to-rank assignment.  Plain Python / numpy: importable without the CUDA library."""
from typing import List

import numpy as np


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
