"""Synthetic KuroSiwo batches with the reference's collated tuple layout (dataset/Dataset.py:826-839).

Inputs mimic `Dataset.__getitem__`: (clamp(raw, 0, 0.15) - mean_c) / std_c with raw ~ Exp(mean_c)
(SAR-speckle-like; configs/train/data_config.json:13-17), masks in {0,1,2,3} with the class
frequencies implied by the reference's class weights (utilities/utilities.py:393-397).
"""
from __future__ import annotations

from typing import List

import torch

DATA_MEAN = (0.0953, 0.0264)
DATA_STD = (0.0427, 0.0215)
CLASS_P = (0.897, 0.024, 0.041, 0.038)
TRAIN_ACTS = (130, 470, 555, 118, 174, 324, 421, 554, 427, 518, 502, 498, 497, 496, 492, 147, 267, 273, 275, 417, 567,
              1111011, 1111004, 1111009, 1111010, 1111006, 1111005)


def make_image(gen: torch.Generator, N: int, H: int, W: int, pin: bool = False, raw_tiles: bool = False) -> torch.Tensor:
    """raw_tiles: return the RAW backscatter planes (incl. a sprinkle of NaN no-data pixels, as in the GeoTIFFs) instead of the
    clamped + normalised tensor - the input of the device-side pipeline (`raw_input: true`, ks_sar_preprocess)."""
    mean = torch.tensor(DATA_MEAN).view(1, 2, 1, 1)
    std = torch.tensor(DATA_STD).view(1, 2, 1, 1)
    raw = torch.empty(N, 2, H, W).exponential_(1.0, generator=gen) * mean
    if raw_tiles:
        raw[torch.rand(raw.shape, generator=gen) < 1e-3] = float("nan")
        return raw.pin_memory() if pin else raw
    x = (raw.clamp_(0, 0.15) - mean) / std
    return x.pin_memory() if pin else x


def make_mask(gen: torch.Generator, N: int, H: int, W: int, pin: bool = False) -> torch.Tensor:
    m = torch.multinomial(torch.tensor(CLASS_P), N * H * W, replacement=True, generator=gen).view(N, H, W)
    return m.pin_memory() if pin else m


def make_batch(seed: int, N: int, H: int = 224, W: int = 224, pin: bool = False, raw_tiles: bool = False) -> List:
    """The 12-tuple the reference trainers unpack (scale_input set, no DEM; change_detection_trainer.py:95-106)."""
    gen = torch.Generator().manual_seed(seed)
    post, pre1, pre2 = (make_image(gen, N, H, W, pin, raw_tiles) for _ in range(3))
    mask = make_mask(gen, N, H, W, pin)
    scale = [torch.full((N,), m) for m in DATA_MEAN], [torch.full((N,), s) for s in DATA_STD]
    clz = torch.randint(1, 4, (N,), generator=gen)
    activ = torch.tensor(TRAIN_ACTS)[torch.randint(0, len(TRAIN_ACTS), (N,), generator=gen)]
    return [scale[0], scale[1], post, mask, scale[0], scale[1], pre1, scale[0], scale[1], pre2, clz, activ]


class SyntheticLoader:
    """Iterable with `.dataset.activations`, standing in for the DataLoader of utilities.prepare_loaders."""

    class _DS:
        activations = list(TRAIN_ACTS)

        def __init__(self, n):
            self.n = n

        def __len__(self):
            return self.n

    def __init__(self, batch_size: int, n_batches: int, seed: int = 999, H: int = 224, W: int = 224, pin: bool = True, distinct: int = 2):
        self.batch_size, self.n_batches = batch_size, n_batches
        self.dataset = SyntheticLoader._DS(batch_size * n_batches)
        self._batches = [make_batch(seed + i, batch_size, H, W, pin) for i in range(min(distinct, n_batches))]

    def __len__(self):
        return self.n_batches

    def __iter__(self):
        for i in range(self.n_batches):
            yield self._batches[i % len(self._batches)]
