"""Rank-aware `RatioSampler`: the reference's mixed-source batch sampler (`vae/data_utils.py:13-87`, used by
`run.py:520-535` when `combined_dataset` is true) for one-process-per-GPU data-parallel training.

Semantics kept from the reference, bit for bit under the same torch seed:
  * the dataset is split on `split_key`; every epoch each subset is shuffled with `torch.randperm` (global generator, in
    the key order of `ratios`), subsets shorter than the longest one are tiled up to its length, and each is cut into
    groups of `round(batch_size * ratio)` indices;
  * a batch is the concatenation of one group per subset, in first-appearance order of the subsets: laid out
    `[source 0 rows..., source 1 rows...]`;
  * `len()` = ceil(longest subset / its group size) (`data_utils.py:51-59`).

What is new: `rank` / `world_size`.  Every rank runs the SAME shuffle (same seed => same global batch) and keeps rows
`rank::world_size` of it -- a strided split preserves the per-source ratio inside every shard, and the union of the
shards is exactly the batch the reference would have drawn, so `world_size` ranks at per-rank batch B/world reproduce
the reference's optimisation trajectory at batch B (the gradients are averaged over ranks, `engine.TrainEngine`).
The global batch is cut to a multiple of `world_size` (drop-last inside the batch) so that every rank steps in
lock-step; `drop_uneven=False` keeps ragged shards instead.
"""
from collections import defaultdict

import torch


class RatioSampler(torch.utils.data.sampler.Sampler):
    def __init__(self, dataset, split_key, ratios=None, batch_size=16, rank=0, world_size=1, generator=None,
                 drop_uneven=True):
        if not (0 <= rank < world_size):
            raise ValueError(f"rank {rank} outside world of size {world_size}")
        self.dataset = dataset
        self.split_key = split_key
        self.batch_size = batch_size          # GLOBAL batch size (the reference's `batch_size`)
        self.rank, self.world_size = rank, world_size
        self.generator = generator            # None = torch's global generator, as the reference
        self.drop_uneven = drop_uneven
        self.split_idxs = self._get_split_idxs()
        self.max_dataset_len = max(len(idxs) for idxs in self.split_idxs.values())
        if ratios is None:
            self.ratios = {k: 1 / len(self.split_idxs) for k in self.split_idxs.keys()}
        else:
            self.ratios = ratios

    def _group_size(self, key):
        return int(torch.round(torch.tensor(self.batch_size * self.ratios[key])))

    def _epoch_plan(self):
        """One shuffled, tiled index vector per subset.  The draws happen in the key order of `ratios` (one
        `torch.randperm` each), which is what fixes the stream under a seed (data_utils.py:72-76)."""
        plan = {}
        for key in self.ratios:
            own = self.split_idxs[key]
            idxs = own[torch.randperm(len(own), generator=self.generator)]
            if len(idxs) < self.max_dataset_len:
                # shorter subsets are tiled up to the longest one: whole copies, then a prefix (data_utils.py:78-81)
                idxs = idxs.repeat(self.max_dataset_len // len(idxs))
                idxs = torch.cat([idxs, idxs[:self.max_dataset_len % len(idxs)]])
            plan[key] = idxs
        return plan

    def global_batches(self):
        """The reference's batches (vae/data_utils.py:33-49), before sharding: batch j takes the j-th group of every
        subset, subsets in first-appearance order; the epoch ends with the subset that runs out of groups first; a
        ragged last group is kept (the reference drops its zip_longest fill values)."""
        plan = self._epoch_plan()
        sizes = {k: self._group_size(k) for k in plan}
        n_batches = min(-(-len(plan[k]) // sizes[k]) for k in plan)
        for j in range(n_batches):
            yield torch.cat([plan[k][j * sizes[k]:(j + 1) * sizes[k]] for k in self.split_idxs])

    def __iter__(self):
        w, r = self.world_size, self.rank
        for batch in self.global_batches():
            if w > 1:
                if self.drop_uneven:
                    batch = batch[:len(batch) // w * w]
                    if len(batch) == 0:
                        continue
                batch = batch[r::w]
            yield batch

    def __len__(self):
        """data_utils.py:51-59: ceil(longest subset / its group size)."""
        longest = max(self.split_idxs, key=lambda k: (len(self.split_idxs[k]), -list(self.split_idxs).index(k)))
        return int(torch.ceil(torch.tensor(len(self.split_idxs[longest]) / self._group_size(longest))))

    def _get_split_idxs(self):
        by_val = defaultdict(list)
        for i, datum in enumerate(self.dataset):
            by_val[datum[self.split_key]].append(i)
        return {val: torch.tensor(idxs) for val, idxs in by_val.items()}
