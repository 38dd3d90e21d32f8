"""The on-disk outputs of the reference's ``inference.py``: ``save_embeddings`` writes
``movie_embeddings.pt`` and ``movie_mapping.csv`` exactly as reference inference.py:146-170 does, so that
``demo.py`` / ``run.py`` of the reference find what they expect (demo.py:30-34 prefers precomputed
embeddings).  The rest of that script is CLI glue around the hot path (and binds the neighbour lists to
the wrong ``forward`` parameter); the coherent entry point is ``PinSage.get_embeddings``.
"""
from __future__ import annotations

import os

import torch


def save_embeddings(embeddings, output_dir, dataset):
    """reference inference.py:146-170: embeddings -> <output_dir>/movie_embeddings.pt (torch.save of the tensor
    as given) and the movieId -> row index table -> <output_dir>/movie_mapping.csv (header ``movieId,index``,
    rows in ``dataset.movie_id_to_idx`` insertion order, as pandas ``to_csv(index=False)`` writes them)."""
    os.makedirs(output_dir, exist_ok=True)
    torch.save(embeddings, os.path.join(output_dir, "movie_embeddings.pt"))
    with open(os.path.join(output_dir, "movie_mapping.csv"), "w", newline="") as f:
        f.write("movieId,index\n")
        for movie_id, idx in dataset.movie_id_to_idx.items():
            f.write(f"{movie_id},{idx}\n")
    print(f"Saved embeddings and mapping to {output_dir}")


def generate_all_embeddings(model, dataset_features, random_walk_sampler, num_neighbors=10, out=None):
    """What reference inference.py:13-57 is for -- embeddings of every movie -- through the coherent entry
    point (``PinSage.get_embeddings``: per-layer resampling + forward over the full feature matrix; the
    reference's own loop slices features to 1,024-row batches while neighbour ids stay global and passes the
    lists as ``edge_index``)."""
    return model.get_embeddings(dataset_features, random_walk_sampler, num_neighbors, out=out)
