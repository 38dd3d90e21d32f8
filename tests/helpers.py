"""Shared helpers for the test-suite (fixtures loading, list <-> padded conversions)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def walk_case(name):
    g = load(name)
    w = g["edge_weights"] if bool(g["has_weights"]) else None
    return dict(ei=g["edge_index"], w=w, starts=g["starts"], W=int(g["W"]), L=int(g["L"]),
                T=int(g["T"]), seed=int(g["seed"]), epoch=int(g["epoch"]), ids=g["ids"],
                weights=g["weights"], nvalid=g["nvalid"])


def quant_shift_for(w):
    """Smallest s in [0, 10] with w * 2**s integral for all weights, else -1 (float path)."""
    if w is None:
        return 0
    for s in range(11):
        q = w.astype(np.float64) * (1 << s)
        if np.all(q == np.floor(q)):
            return s
    return -1


def lists_from_json(g, *keys):
    meta = json.loads(str(g["lists"]))
    return [meta[k] for k in keys]


def rel_row_err(a, b):
    """max over rows of ||a-b|| / max(||b||, tiny): the metric behind the 1e-3 / 1e-2 bars."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    num = np.linalg.norm(a - b, axis=1)
    den = np.maximum(np.linalg.norm(b, axis=1), 1e-30)
    zero = np.linalg.norm(b, axis=1) == 0
    return float(np.max(np.where(zero, np.linalg.norm(a, axis=1), num / den))) if len(a) else 0.0
