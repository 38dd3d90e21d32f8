"""Drop-in for the reference's ``data/negative_sampler.py`` (NegativeSampler), B200-native.

Same constructor and method signatures as reference data/negative_sampler.py:5-123.  The part that costs
time in the reference -- 100 python walks per query plus a dict / sort over the visits (:60-74) -- is ONE
launch of the walk / count / top-T kernel for the whole batch of queries: its top-T list ordered by
(count desc, first visit asc) IS ``sorted(visited_counts.items(), key=count, reverse=True)`` (a stable
sort of a dict in first-visit order), so the rank window ``ranked_items[min_rank:max_rank]`` is a slice of
the kernel's output.  The random picks stay what they are in the reference: ``np.random.choice`` calls on
the GLOBAL numpy RNG, issued in the same order with the same arguments, so under the same ``np.random.seed``
and the same walks the returned negatives are identical (tests/golden/hard_negatives.npz).
Note (as in the reference): with walk_length = 2 at most 200 nodes are visited, so the default window
[2000, 5000) is empty and every query falls back to uniform random negatives (:79-81).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import kernels as K


class NegativeSampler:
    def __init__(self, dataset, random_walk_sampler=None, num_negative_samples=500):
        self.dataset = dataset
        self.random_walk_sampler = random_walk_sampler
        self.num_negative_samples = num_negative_samples
        self.all_movie_indices = list(range(len(dataset.movie_id_to_idx)))          # reference :23

    def sample_random_negatives(self, batch_size, device):
        """reference :25-42 (batch_size is unused there too)."""
        return torch.tensor(np.random.choice(self.all_movie_indices, size=self.num_negative_samples, replace=False),
                            device=device)

    def ranked_visits(self, query_indices, num_walks=100):
        """Per query: every node visited by `num_walks` walks, ordered by (visit count desc, first visit asc)
        -- the reference's ``ranked_items`` (:60-74) for the whole batch in one kernel launch.  Returns
        (ids int32 [n, V] with -1 padding, nvalid int32 [n]) on the host, V = num_walks * walk_length."""
        s = self.random_walk_sampler
        q = torch.as_tensor(np.asarray(query_indices.detach().cpu() if isinstance(query_indices, torch.Tensor)
                                       else query_indices)).reshape(-1).to(torch.int64)
        s._check_nodes(q)
        V = num_walks * s.walk_length
        ids, _counts, _w, nvalid = K.walk_topt(s.csr, q, num_walks, s.walk_length, V, s.seed, s._next_epoch())
        return ids.cpu().numpy(), nvalid.cpu().numpy()

    def sample_hard_negatives(self, query_indices, num_hard_samples=5, max_rank=5000, min_rank=2000):
        """reference :44-99."""
        if self.random_walk_sampler is None:
            raise ValueError("RandomWalkSampler is required for hard negative sampling")
        ids, nvalid = self.ranked_visits(query_indices)                              # 100 walks, as :63
        n_movies = len(self.all_movie_indices)
        hard_negatives = []
        for row, nv in zip(ids, nvalid):
            ranked = row[:nv]
            window = ranked[min_rank:max_rank]
            candidates = [int(v) for v in window if 0 <= v < n_movies]               # `item in all_movie_indices`
            if not candidates:
                sampled = np.random.choice(self.all_movie_indices, size=num_hard_samples, replace=False)
            else:
                sampled = np.random.choice(candidates, size=min(num_hard_samples, len(candidates)), replace=False)
                if len(sampled) < num_hard_samples:
                    taken = set(int(v) for v in sampled)
                    additional = np.random.choice([i for i in self.all_movie_indices if i not in taken],
                                                  size=num_hard_samples - len(sampled), replace=False)
                    sampled = np.concatenate([sampled, additional])
            hard_negatives.append(sampled)
        dev = query_indices.device if isinstance(query_indices, torch.Tensor) else "cpu"
        return torch.tensor(np.asarray(hard_negatives), device=dev)

    def sample_batch_negatives(self, query_indices, device, epoch=0):
        """reference :101-123."""
        random_negatives = self.sample_random_negatives(len(query_indices), device)
        if epoch >= 1 and self.random_walk_sampler is not None:
            num_hard = min(epoch, 6)
            return random_negatives, self.sample_hard_negatives(query_indices, num_hard_samples=num_hard)
        return random_negatives, None
