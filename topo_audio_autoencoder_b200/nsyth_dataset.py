"""Drop-in for the reference's ``nsyth_dataset.py``: the consumer of the distance sweep's on-disk neighbour lists.

Reference: nsyth_dataset.py:7-72.  Same constructor, ``set_epoch`` schedule and ``__getitem__`` contract:
  * positive = a uniformly random one of the ``num_positive_neighbors`` NEAREST neighbours (:52-55);
  * negatives = the ``num_negative_samples`` entries of the ascending neighbour list that end at
    ``current_negative_offset`` (:57-60); the offset starts at ``len(data)`` -- the FAR end of the ordering -- and decays
    by 0.90 per epoch down to 100 (:27-29, :36-39);
  * a training item is ``cat([anchor, positive, negatives])`` (:66-70), otherwise the anchor alone.
The reference opens ``neighbors.pkl`` from the current directory (:18), not from ``./precomputed`` where
``compute_distances`` writes it; ``neighbors_path`` makes that explicit (default: the reference's behaviour).

Neighbour files written by the streaming top-k sweep (``compute_distances(..., top_k=k)``) hold only the k nearest
neighbours.  Positives are unaffected (k >= num_positive_neighbors).  For negatives the window is taken from the far end
of what the file holds: the complete ordering of a 100k-clip collection (10^10 entries) is not materialised, and the
reference's schedule only ever needs "far" negatives until the offset has decayed to 100.
"""
from __future__ import annotations

import pickle
import random
from typing import List

import torch
from torch.utils.data import Dataset


class NSynthDataset(Dataset):
    def __init__(self, data, root_dir, num_positive_neighbors=10, train=False, num_negative_samples=10,
                 neighbors_path: str = "neighbors.pkl"):
        self.data = data
        self.root_dir = root_dir
        self.epoch = 0
        with open(neighbors_path, "rb") as f:
            self.neighbors = pickle.load(f)
        self.num_positive_neighbors = num_positive_neighbors
        self.num_negative_samples = num_negative_samples
        self.initial_negative_offset = len(self.data)
        self.current_negative_offset = self.initial_negative_offset
        self.offset_decay_rate = 0.90
        self.min_negative_offset = 100
        self.train = train
        self._keys = list(self.data.keys())

    def set_epoch(self, epoch):
        self.epoch = epoch
        self.current_negative_offset = max(self.min_negative_offset,
                                           int(self.initial_negative_offset * (self.offset_decay_rate ** epoch)))

    def __len__(self):
        return len(self.data)

    def negative_window(self, item_key) -> List[str]:
        """The neighbour names the reference's index arithmetic selects (:57-60), on whatever list the file holds."""
        order = self.neighbors[item_key]["sorted_neighbors"]
        stop = min(self.current_negative_offset, len(order))           # a truncated (top-k) list: its own far end
        start = stop - self.num_negative_samples
        return [order[i] for i in range(start, stop)]                  # negative start indexes from the end, as in Python

    def _load(self, key):
        return torch.load(f"{self.root_dir}/{key}.pt")

    def __getitem__(self, idx):
        item_key = self._keys[idx]
        waveform = self._load(item_key)
        if not self.train:
            return waveform
        positive_idx = random.randrange(self.num_positive_neighbors)
        positive = self._load(self.neighbors[item_key]["sorted_neighbors"][positive_idx])
        negatives = torch.stack([self._load(neg) for neg in self.negative_window(item_key)])
        return torch.cat([waveform.unsqueeze(0), positive.unsqueeze(0), negatives])
