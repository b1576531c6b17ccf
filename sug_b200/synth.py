"""Synthetic PointDA-10-shaped inputs: clouds uniform in a cube, centred and scaled
to the unit sphere like ``normal_pc`` (reference data/data_utils.py:5-15), layout [B,3,N,1] fp32 as
emitted by the reference's UnifiedPointDG loader (data/dataloader.py:244-330), labels in [0,10)."""
from __future__ import annotations

import numpy as np
import torch


def synth_clouds(B: int, N: int, seed: int):
    rng = np.random.Generator(np.random.PCG64(seed))
    p = rng.random((B, N, 3), dtype=np.float32) * 2 - 1
    p = p - p.mean(1, keepdims=True)
    p = p / np.sqrt((p ** 2).sum(2)).max(1)[:, None, None]
    lab = rng.integers(0, 10, size=(B,))
    x = torch.from_numpy(np.ascontiguousarray(p.transpose(0, 2, 1)[..., None]).astype(np.float32))
    return x, torch.from_numpy(lab.astype(np.int64))
