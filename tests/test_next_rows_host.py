"""SURVEY 8(f) "next" rows, host-side pieces against fixtures from the unmodified reference (CPU only):
N1 save_embeddings / build_bipartite_graph / get_adjacency_list."""
import json
import os
import types

import numpy as np
import torch

from tests import helpers as Hh


def _dataset(g):
    import pandas as pd
    df = pd.DataFrame({"userId": g["userId"], "movieId": g["movieId"], "rating": g["rating"]})
    return types.SimpleNamespace(ratings_df=df, movie_id_to_idx={int(m): i for i, m in enumerate(g["movie_ids_sorted"])},
                                 user_id_to_idx={int(u): i for i, u in enumerate(g["user_ids_sorted"])})


def test_save_embeddings_files_equal_the_reference(tmp_path):
    import mre_b200  # noqa: F401
    from mre_b200.inference import save_embeddings
    g = Hh.load("save_embeddings.npz")
    ds = types.SimpleNamespace(movie_id_to_idx={int(m): i for i, m in enumerate(g["movie_ids"])})
    out = tmp_path / "nested" / "out"
    save_embeddings(torch.from_numpy(g["emb"]), str(out), ds)
    assert sorted(os.listdir(out)) == json.loads(str(g["files"]))
    assert torch.equal(torch.load(out / "movie_embeddings.pt"), torch.from_numpy(g["saved"]))
    assert (out / "movie_mapping.csv").read_text() == str(g["csv"])


def test_bipartite_graph_and_adjacency_list_equal_the_reference():
    import mre_b200  # noqa: F401
    from mre_b200.data.graph_builder import GraphBuilder
    g = Hh.load("item_graph.npz")
    b = GraphBuilder(_dataset(g))
    ei, ew = b.build_bipartite_graph()
    assert ei.dtype == torch.int64 and ew.dtype == torch.float32
    np.testing.assert_array_equal(ei.numpy(), g["bip_ei"])
    np.testing.assert_array_equal(ew.numpy(), g["bip_w"])
    adj = b.get_adjacency_list(ei, ew)
    want = [[tuple(p) for p in row] for row in json.loads(str(g["adj"]))]
    assert adj == want and all(isinstance(w, float) for row in adj for _d, w in row)
    adj_none = b.get_adjacency_list(ei[:, :50])
    assert adj_none == [[tuple(p) for p in row] for row in json.loads(str(g["adj_none"]))]
